#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
OUT=gpurun_out/r02_variants8.txt
: > $OUT
timeout 150 python tools/variant_time.py base >> $OUT 2>> gpurun_out/r02_variants8.err || echo '{"variant": "base", "failed": 1}' >> $OUT
for v in noepi noepi_nose noldtm_keepstore; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 90 python tools/variant_time.py $v >> $OUT 2>> gpurun_out/r02_variants8.err || echo "{\"variant\": \"$v\", \"failed\": 1}" >> $OUT
done
cat $OUT
