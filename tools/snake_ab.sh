# (Applies with profiles/r04c_snake_order.patch - the order was measured and removed)
# One-call check of the snake tile order (ConvHaloParams::reverse): the GPU suite on the new default, then a same-box
# A/B of the full chain with the order on / off (B200SR3_SNAKE) and against the previous build (libb200sr3_prev.so), and
# the per-launch profile of one step both ways.
TAG=${1:-r04c}; mkdir -p gpurun_out
PKG=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3
timeout 110 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.txt 2>&1; echo "tests rc $?"; tail -1 gpurun_out/${TAG}_gpu_tests.txt
Q="--steps 2 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity"
timeout 40 python bench.py $Q > gpurun_out/${TAG}_bench_snake1.json 2> /dev/null; echo "snake1 rc $?"
B200SR3_SNAKE=0 timeout 40 python bench.py $Q > gpurun_out/${TAG}_bench_snake0.json 2> /dev/null; echo "snake0 rc $?"
timeout 30 python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_snake1.txt 2>&1
B200SR3_SNAKE=0 timeout 30 python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_snake0.txt 2>&1
[ -f $PKG/libb200sr3_prev.so ] && B200SR3_LIB=$PKG/libb200sr3_prev.so timeout 40 python bench.py $Q > gpurun_out/${TAG}_bench_prev.json 2> /dev/null
python - <<P
import json
for k in ("snake1", "snake0", "prev"):
    try:
        j = json.load(open("gpurun_out/${TAG}_bench_%s.json" % k))
        print(k, "faces/s", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "conv frac", round(j["roofline"]["frac"], 4), "MHz", j["clocks"]["sm_mhz"])
    except Exception as e:
        print(k, "missing", e)
P
