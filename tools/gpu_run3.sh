#!/bin/bash
# round 2, GPU run 3: 3 / 4 issuer warps, epilogue constants from shared memory; training step breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
: > gpurun_out/r02_variants3.txt
timeout 200 python tools/variant_time.py base >> gpurun_out/r02_variants3.txt 2>> gpurun_out/r02_variants3.err
for v in ni4 ni6; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 100 python tools/variant_time.py $v >> gpurun_out/r02_variants3.txt 2>> gpurun_out/r02_variants3.err || echo "{\"variant\": \"$v\", \"failed\": $?}" >> gpurun_out/r02_variants3.txt
done
cat gpurun_out/r02_variants3.txt
for v in ni4 ni6; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 100 -k "geometries or batch64 or deterministic or golden or streams" 2>&1 | tail -15 > gpurun_out/r02_pytest3_$v.txt
  tail -3 gpurun_out/r02_pytest3_$v.txt
done
