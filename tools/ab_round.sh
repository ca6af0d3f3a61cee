# Same-box A/B of halo-conv variants: default library vs a variant library (B200SR3_LIB).
#   gpurun -- 'bash tools/ab_round.sh r02d <variant .so name> "<shape>" "<shape>" ...'
TAG=${1:-r02x}; VAR=${2:-libb200sr3_ast.so}; shift 2
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
LIBV=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/$VAR
for name in "$@"; do
  echo "== $name" >> $OUT
  echo -n "default : " >> $OUT; python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1
  echo -n "variant : " >> $OUT; B200SR3_LIB=$LIBV python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1
done
cat $OUT
python bench.py --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_bench_default.json 2> /dev/null
B200SR3_LIB=$LIBV python bench.py --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_bench_variant.json 2> /dev/null
python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_default.json'));b=json.load(open('gpurun_out/${TAG}_bench_variant.json'));print('full step faces/s: default',a['value'],' variant',b['value'])"
