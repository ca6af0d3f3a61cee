# N-GPU bench lines exactly as the driver launches them (run as: gpurun --gpus N -- 'bash tools/multi_gpu_round.sh r02f N [configs...]').
TAG=${1:-r02x}; N=${2:-2}; shift 2
CONFIGS=${@:-sr_sr3_VGGF2_16_128_model3}
mkdir -p gpurun_out
PORT=29517
for CFG in $CONFIGS; do
  OUT=gpurun_out/${TAG}_bench_${CFG}_${N}gpu.json
  if [ "$N" = "1" ]; then
    python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline > $OUT 2> gpurun_out/${TAG}_bench_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --config $CFG --steps 2 --warmup 3 > $OUT 2> gpurun_out/${TAG}_bench_${N}gpu.err
  fi
  echo "$CFG N=$N rc $?"; tail -2 gpurun_out/${TAG}_bench_${N}gpu.err
  python -c "import json;a=json.load(open('$OUT'));print({k:a.get(k) for k in ('metric','value','n_gpus','ms_per_step','per_rank_ms_per_step','e2e','psnr_vs_ref_db','outputs_finite')}, a['roofline']['frac'])"
  PORT=$((PORT+1))
done
