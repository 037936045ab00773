#!/bin/bash
# round 2, GPU run 2: turn-token variants (short timeouts), trace, full parity suite, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
: > gpurun_out/r02_variants2.txt
timeout 200 python tools/variant_time.py base >> gpurun_out/r02_variants2.txt 2>> gpurun_out/r02_variants2.err
for v in turn turn_nostore turn_nores turn_nose; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 100 python tools/variant_time.py $v >> gpurun_out/r02_variants2.txt 2>> gpurun_out/r02_variants2.err || echo "{\"variant\": \"$v\", \"failed\": $?}" >> gpurun_out/r02_variants2.txt
done
cat gpurun_out/r02_variants2.txt
FEN_B200_LIB=$PWD/$V/libfen_b200_turn_trace.so timeout 100 python tools/body2_trace.py > gpurun_out/r02_trace_turn.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 150 -rP 2>&1 | tail -300 > gpurun_out/r02_pytest2.txt
tail -4 gpurun_out/r02_pytest2.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
tail -c 1500 gpurun_out/r02_bench2.json; tail -3 gpurun_out/r02_bench2.err
timeout 200 python tools/soak.py 200 64 > gpurun_out/r02_soak.txt 2>&1; tail -2 gpurun_out/r02_soak.txt
