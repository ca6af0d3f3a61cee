# Role counters (timing build) and one ncu --set full capture for the named halo_bench shapes.
#   gpurun -- 'bash tools/roles_round.sh r02c "c2 64->64+res192 @128"'      (first arg: tag; second: the shape ncu captures)
TAG=${1:-r02x}
NCU_SHAPE=${2:-"c2 64->64+res192 @128"}
mkdir -p gpurun_out
LIBT=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so
: > gpurun_out/${TAG}_roles.txt
for name in "c1 64->64 @128" "c2 64->64+id @128" "c2 64->64+res192 @128" "c1 128+64->64 @128" "c1 512->512 @8" "c2 512->512+res1024 @8" "c1 512+512->512 @16"; do
  B200SR3_LIB=$LIBT B200SR3_CONV_TIMING=1 python tools/halo_bench.py 32 20 "$name" 1 >> gpurun_out/${TAG}_roles.txt 2>&1
done
cat gpurun_out/${TAG}_roles.txt
python tools/halo_bench.py 32 3 "$NCU_SHAPE" 1 > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o gpurun_out/${TAG}_halo -f python tools/halo_bench.py 32 3 "$NCU_SHAPE" 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/${TAG}_ncu.log
