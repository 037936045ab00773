#!/bin/bash
# round 2, GPU run 4: tile-completion ring protocol with 2 / 3 / 4 / 6 issuer warps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
OUT=gpurun_out/r02_variants4.txt
: > $OUT
timeout 150 python tools/variant_time.py base >> $OUT 2>> gpurun_out/r02_variants4.err || echo '{"variant": "base", "failed": 1}' >> $OUT
for v in ni2s ni3 ni4 ni6; do
  if FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 90 python tools/variant_time.py $v >> $OUT 2>> gpurun_out/r02_variants4.err; then
    FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -k "geometries or batch64 or deterministic or golden or streams" 2>&1 | tail -4 > gpurun_out/r02_pytest4_$v.txt
    tail -2 gpurun_out/r02_pytest4_$v.txt
  else
    echo "{\"variant\": \"$v\", \"failed\": 1}" >> $OUT
  fi
done
cat $OUT
timeout 300 python -m pytest tests -m gpu -q -x --timeout 100 2>&1 | tail -5 > gpurun_out/r02_pytest4.txt; tail -3 gpurun_out/r02_pytest4.txt
