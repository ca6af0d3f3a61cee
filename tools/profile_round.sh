set -x
mkdir -p gpurun_out
python tools/config_sweep.py > gpurun_out/r01f_config_sweep.jsonl 2> gpurun_out/r01f_config_sweep.err
python bench.py > gpurun_out/r01f_bench_line.json 2> gpurun_out/r01f_bench.err; echo "bench rc $?"
python tools/profile_step.py 32 128 600 > gpurun_out/r01f_step_profile_B32_R128.txt 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r01f_ncu_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_ncu_bench.log 2>&1
echo "ncu list rc $?"
for sh in "c1 512+256->256 @32:n256" "c1 128+64->64 @128:n64"; do
  name="${sh%%:*}"; tag="${sh##*:}"
  python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/r01f_plain_$tag.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o gpurun_out/r01f_halo_$tag -f python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/r01f_ncu_$tag.log 2>&1
  echo "ncu full $tag rc $?"
done
ls -la gpurun_out | tail -15
