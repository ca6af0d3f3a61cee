# Same-box A/B of two library builds over BASELINE configs: bash tools/ab_configs.sh <tag> <variant .so> <config> ...
TAG=${1:-r03x}; VAR=$2; shift 2
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab_configs.txt
: > $OUT
LIBV=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/$VAR
for rep in 1 2; do
for CFG in "$@"; do
  for arm in variant default; do
    if [ $arm = variant ]; then export B200SR3_LIB=$LIBV; else unset B200SR3_LIB; fi
    timeout 600 python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_tmp.json 2> /dev/null
    python -c "import json;a=json.load(open('gpurun_out/${TAG}_tmp.json'));print('$CFG', '$arm', round(a['value'],2), 'faces/s', a['clocks']['sm_mhz'], 'MHz')" >> $OUT
  done
done
done
cat $OUT
