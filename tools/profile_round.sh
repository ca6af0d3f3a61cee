# ncu evidence of the round (run as: gpurun -- 'bash tools/profile_round.sh r02n'): the launch list of one bench run and
# `--set full` captures of the dominant conv class (N=256 @32x32) and of the Cout=64 @128x128 class.
TAG=${1:-r02x}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu list rc $?"
for sh in "c1 512+256->256 @32:n256" "c1 128+64->64 @128:n64" "c2 64->64+res128 @128:n64deep"; do
  name="${sh%%:*}"; tag="${sh##*:}"
  python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/${TAG}_plain_$tag.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o gpurun_out/${TAG}_halo_$tag -f python tools/halo_bench.py 32 3 "$name" 1 > gpurun_out/${TAG}_ncu_$tag.log 2>&1
  echo "ncu full $tag rc $?"
  # keep the raw metric table, not the 12 MB report (gpurun copies back at most 64 MiB)
  ncu -i gpurun_out/${TAG}_halo_$tag.ncu-rep --page raw --csv > gpurun_out/${TAG}_halo_$tag.raw.csv 2>/dev/null && rm -f gpurun_out/${TAG}_halo_$tag.ncu-rep
done
ls -la gpurun_out | grep ${TAG}
