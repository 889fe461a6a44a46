"""Turn the round's profiler artefacts (gpurun_out/) into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/r02_launches.csv      # per-step share table + DRAM traffic
    python profiles/summarize.py ncu gpurun_out/r02_linear320.ncu-rep ...  # key `ncu --set full` metrics per kernel

Per-launch times in the launch list are cold-cache and serialised (profiler replay): compare SHARES, not absolutes.
"""
import collections
import csv
import json
import re
import subprocess
import sys


def read_rows(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value")}
    out = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        e = out.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]]})
        e[r[col["Metric Name"]]] = v
    return list(out.values())


def short(name):
    m = re.search(r"pp::(\w+(?:<[^>]*>)?)", name)
    return m.group(1) if m else re.sub(r"\(.*", "", name)[:50]


def launches(path, out_json=None):
    ev = read_rows(path)
    # one step = from one time_embed_kernel to the next (the first kernel of the recorded step program)
    starts = [i for i, e in enumerate(ev) if "time_embed_kernel" in e["name"]]
    if len(starts) < 2:
        print("no whole step captured")
        return
    a, b = starts[-2], starts[-1]  # the last whole step in the window (warmest)
    step = ev[a:b]
    tot = sum(e.get("gpu__time_duration.sum", 0.0) for e in step)
    rd = sum(e.get("dram__bytes_read.sum", 0.0) for e in step)
    wr = sum(e.get("dram__bytes_write.sum", 0.0) for e in step)
    agg = collections.OrderedDict()
    for e in step:
        k = short(e["name"])
        g = agg.setdefault(k, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += e.get("gpu__time_duration.sum", 0.0)
        g[2] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    print(f"one step: {len(step)} launches, {tot / 1e6:.2f} ms serialised, DRAM read {rd / 1e9:.2f} GB write {wr / 1e9:.2f} GB")
    print(f"{'kernel':44s} {'n':>4s} {'us':>9s} {'share':>6s} {'us/launch':>9s} {'DRAM MB':>9s}")
    for k, (n, t, by) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} {n:4d} {t / 1e3:9.1f} {100 * t / tot:5.1f}% {t / n / 1e3:9.1f} {by / 1e6:9.1f}")
    if out_json:
        json.dump({"config": "C2", "dram_bytes_per_step": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "launches": len(step), "serialised_ms": tot / 1e6,
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one un-graphed "
                             "C2 step (profiles/r02_profile.sh), summed"}, open(out_json, "w"), indent=1)


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput", "sm__throughput.avg.pct_of_peak",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_xu", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "smsp__pcsamp_warps_issue_stalled"]


def ncu(paths):
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            print(p, "empty")
            continue
        hdr, units = rows[0], rows[1]
        print(f"== {p}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"-- {short(d.get('Kernel Name', ''))} grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for h, u in zip(hdr, units):
                if any(h.startswith(k) for k in KEYS) and d.get(h) not in (None, "", "0"):
                    print(f"   {h} [{u}] = {d[h]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        ncu(sys.argv[2:])
