#!/bin/bash
# The ncu evidence under profiles/ (one gpurun call): launch list of the forward, full captures of the body kernel and of
# the batched weight gradient.  Every command runs once without ncu first (B200_PROFILING.md).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-eager --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
python tools/fwd_time.py 64 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:body2_umma -s 4 -c 1 -o gpurun_out/r02_body2_kernel_full python tools/fwd_time.py 64 > gpurun_out/ncu2.log 2>&1
python tools/step_once.py 32 2 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wgrad_batch -s 3 -c 1 -o gpurun_out/r02_wgrad_batch_kernel_full python tools/step_once.py 32 2 > gpurun_out/ncu3.log 2>&1
