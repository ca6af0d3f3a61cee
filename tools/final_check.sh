TAG=r01h
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python tools/config_sweep.py > gpurun_out/${TAG}_config_sweep.jsonl 2> gpurun_out/${TAG}_config_sweep.err
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_B32_R128.txt 2>&1
for name in "c1 512+256->256 @32" "c1 128->128 @64" "c1 64->64 @128"; do
  B200SR3_LIB=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so B200SR3_CONV_TIMING=1 python tools/halo_bench.py 32 20 "$name" 1 >> gpurun_out/${TAG}_roles.txt 2>&1
done
cat gpurun_out/${TAG}_gpu_tests.txt gpurun_out/${TAG}_smoke.txt gpurun_out/${TAG}_roles.txt
