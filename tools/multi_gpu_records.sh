#!/bin/bash
# multi-GPU records: bench.py (inference + train sub-record) and BASELINE config 4 at the stated 65 536 images
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "1" ]; then
  timeout 300 python tools/configs_bench.py 65536 > gpurun_out/r02_configs_n1.txt 2>&1; cat gpurun_out/r02_configs_n1.txt
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'e2e_fp32', d['e2e_fp32']['value']); t=d['train']; print('train', t['value'], t['ms_per_step'], 'exposed allreduce us', t['allreduce_exposed_us'])"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/configs_bench.py 65536 > gpurun_out/r02_configs_n$N.txt 2>&1; grep -E "^C[34]" gpurun_out/r02_configs_n$N.txt
fi
