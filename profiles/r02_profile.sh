#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; numbers under a profiler are never bench values):
#   1. isolated kernel timings (CUDA events, no profiler)          -> gpurun_out/r02_prof_kernels.txt
#   2. launch list + DRAM bytes of un-graphed C2 steps            -> gpurun_out/r02_launches.csv
#   3. `ncu --set full` of the dominant kernels in isolation       -> gpurun_out/r02_*.ncu-rep
#   4. role hand-over trace of the GEMM kernel (-DGEMM_TRACE build) -> gpurun_out/r02_gemm_trace.txt
set -x
mkdir -p gpurun_out
python profiles/prof_kernels.py > gpurun_out/r02_prof_kernels.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 380 -c 1150 \
    --csv --log-file gpurun_out/r02_launches.csv python bench_probe.py --no-graph --steps 4 --reps 0 > gpurun_out/r02_probe.log 2>&1
for spec in "linear320:linear 320->320 M:gemm_conv_kernel:3" "conv320:conv 320->320 @64:gemm_conv_kernel:3" \
            "attn40:self-attn d40:attn2_kernel:1" "gnfused:conv320+stats:gn_|gemm_conv:6" "cfgddim:cfg_ddim C2:cfg_ddim_kernel:2" \
            "geglu320:geglu 320->2560 M=65536 tile:gemm_conv_kernel:2" "qkv320:one launch:gemm_conv_kernel:4"; do
  IFS=: read name case regex cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 2 -c $cnt -o gpurun_out/r02_$name \
      python profiles/prof_kernels.py "$case" > gpurun_out/r02_ncu_$name.log 2>&1
  # the reports are 10-20 MB each and gpurun brings back at most 64 MiB: keep the raw-metric page as CSV, drop the
  # big reports (the two small ones travel as they are)
  ncu -i gpurun_out/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02_$name.raw.csv 2>/dev/null
  case $name in cfgddim|attn40) ;; *) rm -f gpurun_out/r02_$name.ncu-rep ;; esac
done
if [ -f powerpaint_b200/_variants/libtrace.so ]; then
  for w in linear qk geglu conv; do PP_B200_LIB=powerpaint_b200/_variants/libtrace.so python profiles/gemm_trace.py $w; done > gpurun_out/r02_gemm_trace.txt 2>&1
fi
ls -la gpurun_out/ | head -40; du -sh gpurun_out
