# One GPU call that produces the round's measured evidence (run as: gpurun -- 'bash tools/profile_round.sh r01g').
TAG=${1:-r01x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/${TAG}_gpu_tests.txt
python tools/config_sweep.py > gpurun_out/${TAG}_config_sweep.jsonl 2> gpurun_out/${TAG}_config_sweep.err
python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_line.json 2>> gpurun_out/${TAG}_bench.err
python tools/profile_step.py 32 128 600 > gpurun_out/${TAG}_step_profile_B32_R128.txt 2>&1
python tools/handoff_bench.py 512 128 > gpurun_out/${TAG}_handoff_bench.json 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu list rc $?"
for sh in "c1 512+256->256 @32:n256" "c1 128+64->64 @128:n64"; do
  name="${sh%%:*}"; tag="${sh##*:}"
  B200SR3_LIB=$PWD/3d-super-resolution-face-reconstruction_b200/b200sr3/libb200sr3_timing.so B200SR3_CONV_TIMING=1 \
    python tools/halo_bench.py 32 20 "$name" 1 > gpurun_out/${TAG}_roles_$tag.txt 2>&1      # needs `build.py --timing`
  python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/${TAG}_plain_$tag.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o gpurun_out/${TAG}_halo_$tag -f python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/${TAG}_ncu_$tag.log 2>&1
  echo "ncu full $tag rc $?"
done
ls gpurun_out | grep ${TAG}
