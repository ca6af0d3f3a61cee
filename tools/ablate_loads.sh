LIBT=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so
for cls in "c1 64->64 @128" "c2 64->64+res192 @128" "c1 128+64->64 @128"; do
  for mask in 0 4 8; do
    echo -n "ablate=$mask  "; B200SR3_LIB=$LIBT B200SR3_CONV_ABLATE=$mask python tools/power_by_class.py 32 "$cls" 2>&1 | tail -1
  done
done
