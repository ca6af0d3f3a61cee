# One GPU call: tests, smoke, headline bench line, per-launch step profile (run as: gpurun -- 'bash tools/gpu_round.sh r02a').
# Extra arguments after the tag are environment assignments for an A/B bench of the previous kernel paths.
TAG=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python bench.py --config sr_sr3_VGGF2_32_128_model2 --no-cpu-baseline --no-torch-baseline > gpurun_out/${TAG}_bench_line_cfg5.json 2>> gpurun_out/${TAG}_bench.err; echo "bench cfg5 rc $?"; head -c 900 gpurun_out/${TAG}_bench_line_cfg5.json; echo
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_B32_R128.txt 2>&1
if [ -n "$2" ]; then
  env "${@:2}" python bench.py --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_bench_line_AB.json 2>> gpurun_out/${TAG}_bench.err
  python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_line.json'));b=json.load(open('gpurun_out/${TAG}_bench_line_AB.json'));print('A/B faces/s: new',a['value'],' old (${@:2})',b['value'])"
fi
tail -5 gpurun_out/${TAG}_gpu_tests.txt; cat gpurun_out/${TAG}_smoke.txt; tail -3 gpurun_out/${TAG}_bench.err
python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_line.json'));print({k:a.get(k) for k in ('value','ms_per_diffusion_step','psnr_vs_ref_db','final_max_abs','gpu_launches')}, a['roofline']['frac'], a['e2e'])"
grep -E "downs\.(0|3|6|9|12) |tail|total" gpurun_out/${TAG}_step_profile_B32_R128.txt
