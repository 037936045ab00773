#!/bin/bash
# round 2, GPU run 1: parity suite, bench line, body-kernel variants, external ceiling check, LR kernel bandwidth
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -rP 2>&1 | tail -400 > gpurun_out/r02_pytest1.txt
tail -5 gpurun_out/r02_pytest1.txt
: > gpurun_out/r02_variants.txt
timeout 300 python tools/variant_time.py base >> gpurun_out/r02_variants.txt 2>> gpurun_out/r02_variants.err
for v in turn turn_nostore turn_nores turn_nose; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 300 python tools/variant_time.py $v >> gpurun_out/r02_variants.txt 2>> gpurun_out/r02_variants.err
done
cat gpurun_out/r02_variants.txt
FEN_B200_LIB=$PWD/$V/libfen_b200_turn_trace.so timeout 300 python tools/body2_trace.py > gpurun_out/r02_trace_turn.txt 2>&1
FEN_B200_LIB=$PWD/$V/libfen_b200_base_trace.so timeout 300 python tools/body2_trace.py > gpurun_out/r02_trace_base.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -c 3000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
timeout 300 python tools/ceiling_check.py > gpurun_out/r02_ceiling.txt 2>&1
cat gpurun_out/r02_ceiling.txt
timeout 200 python tools/lr_bandwidth.py > gpurun_out/r02_lr_bw.txt 2>&1
cat gpurun_out/r02_lr_bw.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref1.json 2> gpurun_out/r02_bench_ref1.err
cat gpurun_out/r02_bench_ref1.json
