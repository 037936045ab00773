# Developer script (one gpurun call): A/B timing of variants/libfen_b200_<name>.so against the product library.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; V=$PWD/face-super-resolution_b200/variants
OUT=gpurun_out/${1:-variants}.txt; : > $OUT; shift
python tools/variant_time.py product 64 >> $OUT 2>&1
for v in "$@" ; do
  FEN_B200_LIB=$V/libfen_b200_$v.so python tools/variant_time.py $v 64 >> $OUT 2>&1
done
python tools/variant_time.py product 64 >> $OUT 2>&1
cut -c1-170 $OUT
