# Targeted-metric captures (tools/profile_metrics_fallback.sh's metric list) of the Cout=64 layers at 128x128 and of an
# 8x8-level conv on the current build: isolated launches of tools/halo_bench.py, GroupNorm fused, B=32.
TAG=${1:-r04b}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
mkdir -p gpurun_out
i=0
for L in "c1 128+64->64 @128" "c2 64->64+res192 @128" "c1 512->512 @8"; do
  i=$((i+1))
  timeout 75 ncu --metrics $M --clock-control none -k regex:conv_halo -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_layer${i}_metrics.csv python tools/halo_bench.py 32 3 "$L" 1 > /dev/null 2>&1; echo "$L rc $?"
done
