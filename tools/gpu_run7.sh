#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
FEN_BODY_MS=1 timeout 200 python tools/step_once.py 32 3 > gpurun_out/r02_body_train_ms.txt 2>&1; cat gpurun_out/r02_body_train_ms.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_step_launches.csv python tools/step_once.py 32 2 > gpurun_out/ncu_step.log 2>&1
python tools/launch_summary.py gpurun_out/r02_step_launches.csv > gpurun_out/r02_step_launches_summary.txt 2>&1; head -40 gpurun_out/r02_step_launches_summary.txt
