# Where the energy of a power-capped conv launch goes: 2 s loops with one role's work removed (timing build).
TAG=${1:-r02x}
LIBT=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so
OUT=gpurun_out/${TAG}_energy_ablate.md
: > $OUT
for cls in "c1 128+64->64 @128" "c1 256+128->128 @64" "c1 512+256->256 @32"; do
  for mask in 0 2 32 34 51 59; do
    echo "ablate=$mask" >> $OUT
    B200SR3_LIB=$LIBT B200SR3_CONV_ABLATE=$mask python tools/power_by_class.py 32 "$cls" 2>&1 | tail -1 >> $OUT
  done
done
cat $OUT
