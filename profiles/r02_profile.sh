#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; numbers under a profiler are never bench values):
#   1. launch list + DRAM bytes of un-graphed C2 steps            -> gpurun_out/r02_launches.csv
#   2. `ncu --set full` of the dominant kernels in isolation       -> gpurun_out/r02_*.ncu-rep
#   3. isolated kernel timings (CUDA events, no profiler)          -> gpurun_out/r02_prof_kernels.txt
set -x
mkdir -p gpurun_out
python profiles/prof_kernels.py > gpurun_out/r02_prof_kernels.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 450 -c 1300 \
    --csv --log-file gpurun_out/r02_launches.csv python bench_probe.py --no-graph --steps 4 --reps 0 > gpurun_out/r02_probe.log 2>&1
for spec in "linear320:linear 320->320:gemm_conv_kernel:3" "conv320:conv 320->320 @64:gemm_conv_kernel:3" \
            "attn40:self-attn d40:attn2_kernel:1" "gnfused:conv320+stats:gn_|gemm_conv:6" "cfgddim:cfg_ddim C2:cfg_ddim_kernel:2" \
            "geglu320:geglu 320->2560:gemm_conv_kernel:2"; do
  IFS=: read name case regex cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 2 -c $cnt -o gpurun_out/r02_$name \
      python profiles/prof_kernels.py "$case" > gpurun_out/r02_ncu_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
