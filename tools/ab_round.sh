# Same-box A/B of halo-conv variants: default library vs a variant library (B200SR3_LIB), optional forced tile shape.
#   gpurun -- 'bash tools/ab_round.sh r02d'
TAG=${1:-r02x}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
LIBV=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_ast.so
for name in "c2 64->64+res192 @128" "c2 64->64+res128 @128" "c2 64->64+id @128" "c1 64->64 @128" "c1 128+64->64 @128" "c2 128->128+res384 @64" "c2 128->128+res192 @64" "c2 128->128+id @64" "c1 256+128->128 @64"; do
  echo "== $name" >> $OUT
  echo -n "default           : " >> $OUT; python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1
  echo -n "default  MT=1     : " >> $OUT; B200SR3_HALO_MT=1 python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1
  echo -n "deep ring MT=1    : " >> $OUT; B200SR3_LIB=$LIBV B200SR3_HALO_MT=1 python tools/halo_bench.py 32 30 "$name" 1 >> $OUT 2>&1
done
cat $OUT
