#!/bin/bash
# Samples SM clock / power / throttle reasons while ONE conv shape runs back to back for a few seconds.
# usage: tools/power_probe.sh "<halo_bench name filter>" <iters>
python tools/halo_bench.py 32 "$2" "$1" 1 &
PID=$!
sleep 6   # import torch + setup
for i in $(seq 1 12); do
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
  sleep 0.25
done
wait $PID
