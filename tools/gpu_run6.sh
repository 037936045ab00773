#!/bin/bash
# round 2, GPU run 6: deferred batched weight gradients
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x --timeout 100 2>&1 | tail -30 > gpurun_out/r02_pytest6.txt; tail -5 gpurun_out/r02_pytest6.txt
timeout 200 python tools/train_bench.py > gpurun_out/r02_train_breakdown.txt 2>&1; cat gpurun_out/r02_train_breakdown.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 100 2>&1 | tail -8 > gpurun_out/r02_pytest6b.txt; tail -3 gpurun_out/r02_pytest6b.txt
