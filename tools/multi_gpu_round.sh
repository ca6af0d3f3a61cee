# N-GPU bench exactly as the driver launches it (run as: gpurun --gpus N -- 'bash tools/multi_gpu_round.sh r02f N').
TAG=${1:-r02x}; N=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "rc $?"; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
python -c "import json;a=json.load(open('gpurun_out/${TAG}_bench_${N}gpu.json'));print({k:a.get(k) for k in ('value','n_gpus','ms_per_step','per_rank_ms_per_step','e2e','psnr_vs_ref_db','outputs_finite')})"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_${N}gpu_reference.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err
echo "reference rc $?"; head -c 600 gpurun_out/${TAG}_bench_${N}gpu_reference.json
