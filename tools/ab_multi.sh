# Same-box comparison of several library builds (first argument: tag; then .so names under b200sr3/, "default" = libb200sr3.so):
# isolated layers, the head / tail / total of the per-launch step profile, and the full bench.
#   gpurun -- 'bash tools/ab_multi.sh r03p "default libb200sr3_f2.so libb200sr3_f3.so" "c1 64->64 @128" ...'
TAG=${1:-r03x}; LIBS=$2; shift 2
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
DIR=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3
for rep in 1 2; do
for lib in $LIBS; do
  L=$DIR/$lib; [ "$lib" = default ] && L=$DIR/libb200sr3.so
  echo "== $lib (pass $rep)" >> $OUT
  for name in "$@"; do B200SR3_LIB=$L python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1; done
  B200SR3_LIB=$L python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_${lib}.txt 2>&1
  grep -E "^downs\.0 |^downs\.1\.conv1|^ups\.17\.conv1|^ups\.8\.conv1|tail|total" gpurun_out/${TAG}_step_${lib}.txt >> $OUT
done
done
for lib in $LIBS; do
  L=$DIR/$lib; [ "$lib" = default ] && L=$DIR/libb200sr3.so
  B200SR3_LIB=$L python bench.py --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_bench_${lib}.json 2> /dev/null
  python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_${lib}.json'));print('$lib full bench faces/s', a['value'], 'psnr', a.get('psnr_vs_ref_db'))" >> $OUT
done
cat $OUT
