#!/bin/bash
# Run on the GPU box: launch list of the step + one full ncu capture per hot kernel -> gpurun_out/ (round 2 names).
set -u
mkdir -p gpurun_out
B="python bench.py --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-extras --kernel-iters 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv 70 > gpurun_out/r02_launches_summary.txt
for k in bk_fused_kernel fwd_single_kernel bk_place_kernel gemm_tf32x3_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 \
      -o gpurun_out/r02_$k $B > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/r02_*.ncu-rep
# kernels added after round 2's lease: event timings (no graph, L2 flushed between launches)
timeout 300 python tools/bench_small_kernels.py > gpurun_out/r02_small_kernels.txt 2>&1
