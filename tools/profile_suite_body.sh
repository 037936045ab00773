#!/bin/bash
# The part of tools/profile_suite.sh that depends on the body kernel and the per-layer convolution kernel (one gpurun
# call): forward launch list, ncu capture of the body kernel, training step launch list and timing.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-train --no-eager --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
python tools/fwd_time.py 64 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:body2_umma -s 4 -c 1 -o gpurun_out/r02_body2_kernel_full python tools/fwd_time.py 64 > gpurun_out/ncu2.log 2>&1
python tools/step_once.py 32 2 > gpurun_out/plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_step_launches.csv python tools/step_once.py 32 2 > gpurun_out/ncu3a.log 2>&1
python tools/launch_summary.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launches_summary.txt 2>&1
FEN_BODY_MS=1 python tools/step_once.py 32 3 > gpurun_out/r02_train_body.txt 2>&1
python tools/train_bench.py > gpurun_out/r02_train_step.txt 2>&1; cat gpurun_out/r02_train_body.txt >> gpurun_out/r02_train_step.txt
python tools/fwd_time.py 16 32 64 128 > gpurun_out/r02_fwd_vs_batch.txt 2>&1
cat gpurun_out/r02_launches_summary.txt; tail -3 gpurun_out/r02_train_step.txt; cat gpurun_out/r02_fwd_vs_batch.txt
