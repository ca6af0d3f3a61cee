# ncu --set full of the two HBM-bound ends of the step inside a real chain (run as: gpurun -- 'bash tools/profile_hbm_ends.sh r02t'):
# the head (downs.0: conv_halo_kernel<64, 2, 0, 0, 1, 1, ...>) and the tail (final_conv + update: conv_halo_kernel<16, ...>).
TAG=${1:-r02x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity"
$CMD > gpurun_out/${TAG}_plain_ends.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_kernel<\(int\)16' -s 50 -c 1 -o gpurun_out/${TAG}_tail -f $CMD > gpurun_out/${TAG}_ncu_tail.log 2>&1
echo "ncu tail rc $?"
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_kernel<\(int\)64, \(int\)2, \(bool\)0, \(int\)0, \(int\)1, \(bool\)1' -s 50 -c 1 -o gpurun_out/${TAG}_head -f $CMD > gpurun_out/${TAG}_ncu_head.log 2>&1
echo "ncu head rc $?"
for t in head tail; do
  ncu -i gpurun_out/${TAG}_$t.ncu-rep --page raw --csv > gpurun_out/${TAG}_$t.raw.csv 2>/dev/null && rm -f gpurun_out/${TAG}_$t.ncu-rep
done
ls -la gpurun_out | grep ${TAG}
