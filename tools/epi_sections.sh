# Epilogue section timers of isolated layers (timing build compiled with -DHALO_EPI_PROFILE=1), with roles ablated:
#   mask 0 = everything, 2 = transform arrives without touching the tile, 4 = no weight loads, 8 = no halo loads
LIBT=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so
for name in "c1 64->64 @128" "c1 128+64->64 @128"; do
  for mask in 0 2 4 12 14; do
    echo "== $name ablate=$mask"
    B200SR3_LIB=$LIBT B200SR3_CONV_TIMING=1 B200SR3_CONV_ABLATE=$mask timeout 120 python tools/halo_bench.py 32 20 "$name" 1 2>&1 | grep -E "halo timing|GroupNorm table|us " | cut -c1-330
  done
done
