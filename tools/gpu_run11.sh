#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
V=face-super-resolution_b200/variants
OUT=gpurun_out/r02_variants11.txt
: > $OUT
timeout 150 python tools/variant_time.py base >> $OUT 2>> gpurun_out/r02_variants11.err || echo '{"variant": "base", "failed": 1}' >> $OUT
for v in resahead resahead_sc; do
  FEN_B200_LIB=$PWD/$V/libfen_b200_$v.so timeout 90 python tools/variant_time.py $v >> $OUT 2>> gpurun_out/r02_variants11.err || echo "{\"variant\": \"$v\", \"failed\": 1}" >> $OUT
done
timeout 150 python tools/variant_time.py base >> $OUT 2>> gpurun_out/r02_variants11.err
cat $OUT
timeout 300 python bench.py --steps 20 --warmup 3 --no-train --no-eager --no-cpu-baseline > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench11.json')); print(d['value'], d['e2e'], d['e2e_u8'])"
