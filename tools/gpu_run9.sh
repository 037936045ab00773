#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
OUT=gpurun_out/r02_variants9.txt
: > $OUT
if timeout 150 python tools/variant_time.py pair >> $OUT 2>> gpurun_out/r02_variants9.err; then
  FEN_B200_LIB=$PWD/$V/libfen_b200_nopair.so timeout 90 python tools/variant_time.py nopair >> $OUT 2>> gpurun_out/r02_variants9.err
  timeout 150 python tools/variant_time.py pair >> $OUT 2>> gpurun_out/r02_variants9.err
  FEN_B200_LIB=$PWD/$V/libfen_b200_nopair.so timeout 90 python tools/variant_time.py nopair >> $OUT 2>> gpurun_out/r02_variants9.err
  cat $OUT
  timeout 400 python -m pytest tests -m gpu -q --timeout 100 2>&1 | tail -8 > gpurun_out/r02_pytest9.txt; tail -3 gpurun_out/r02_pytest9.txt
  timeout 200 python tools/train_bench.py 2>&1 | head -3
else
  echo '{"variant": "pair", "failed": 1}' >> $OUT; cat $OUT; tail -5 gpurun_out/r02_variants9.err
fi
