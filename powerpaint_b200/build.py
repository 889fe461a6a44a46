"""In-tree build of the CUDA/C-ABI library (sm_100a only).

`python -m powerpaint_b200.build` compiles every `csrc/*.cu` with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` and links
`powerpaint_b200/libpowerpaint_b200.so`. nvcc cross-compiles without a GPU; the
resulting .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libpowerpaint_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    *os.environ.get("PP_EXTRA_NVCC_FLAGS", "").split(),  # developer switches, e.g. -DATT2_PROFILE
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "powerpaint_b200.h"]
    BUILD_DIR.mkdir(exist_ok=True)
    hdr_digest = _digest(headers)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = BUILD_DIR / (src.stem + ".o")
        stamp = BUILD_DIR / (src.stem + ".stamp")
        want = hdr_digest + _digest([src])
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
            return obj, False, ""
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(PKG_DIR.parent / "include"), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(want)
        (BUILD_DIR / (src.stem + ".ptxas.log")).write_text(r.stderr)
        return obj, True, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    rebuilt = any(r[1] for r in results)
    if verbose:
        for _, did, log in results:
            if did:
                print(log)
    if rebuilt or force or not LIB_PATH.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *[str(r[0]) for r in results], "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
