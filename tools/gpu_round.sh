# One GPU call: tests, smoke, headline bench line, per-launch step profile (run as: gpurun -- 'bash tools/gpu_round.sh r02a').
TAG=${1:-r02x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1
python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_B32_R128.txt 2>&1
tail -5 gpurun_out/${TAG}_gpu_tests.txt; cat gpurun_out/${TAG}_smoke.txt; tail -3 gpurun_out/${TAG}_bench.err; head -c 3000 gpurun_out/${TAG}_bench_line.json
