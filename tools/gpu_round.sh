# One GPU call: tests, smoke, headline bench line, per-launch step profile (run as: gpurun -- 'bash tools/gpu_round.sh r02a').
# Extra arguments after the tag are environment assignments for a same-box A/B bench (e.g. B200SR3_LIB=<previous build>).
TAG=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_B32_R128.txt 2>&1
if [ -n "$2" ]; then
  env "${@:2}" python bench.py --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_bench_line_AB.json 2>> gpurun_out/${TAG}_bench.err
  python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_line.json'));b=json.load(open('gpurun_out/${TAG}_bench_line_AB.json'));print('A/B faces/s: new',a['value'],' other',b['value'])"
fi
if [ -f 3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so ]; then
  B200SR3_LIB=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so B200SR3_CONV_TIMING=1 B200SR3_TIMING_OPS=downs.0,final_conv.tail,downs.3,ups.17.conv1,ups.17.conv2,downs.4.conv1,ups.0.conv1,ups.8.conv1,downs.1.conv1 python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_roles_in_situ.txt 2>&1
fi
tail -5 gpurun_out/${TAG}_gpu_tests.txt; cat gpurun_out/${TAG}_smoke.txt; tail -3 gpurun_out/${TAG}_bench.err
python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_line.json'));print({k:a.get(k) for k in ('value','ms_per_diffusion_step','psnr_vs_ref_db','final_max_abs','gpu_launches')}, a['roofline']['frac'], a['e2e'])"
grep -E "downs\.(0|3|6|9|12) |tail|total" gpurun_out/${TAG}_step_profile_B32_R128.txt
