# Targeted-metric captures (no SASS instrumentation) of the two launches whose `ncu --set full` replay failed with
# LaunchFailed on the final build (kernels at 128 registers x 512 threads): N=256 conv and the head inside a chain.
TAG=${1:-r03x}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
mkdir -p gpurun_out
timeout 600 ncu --metrics $M --clock-control none -k regex:conv_halo -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_n256_metrics.csv python tools/halo_bench.py 32 3 "c1 512+256->256 @32" 1 > /dev/null 2>&1; echo "n256 rc $?"
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_kernel<\(int\)64, \(int\)2, \(bool\)0, \(int\)0, \(int\)1, \(bool\)1' -s 50 -c 1 --csv --log-file gpurun_out/${TAG}_head_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity > /dev/null 2>&1; rc=$?; echo "head rc $rc"
if [ $rc -ne 0 ]; then      # the multi-pass replay of this launch sometimes ends in LaunchFailed: DRAM bytes and duration fit one pass
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_kernel<\(int\)64, \(int\)2, \(bool\)0, \(int\)0, \(int\)1, \(bool\)1' -s 50 -c 1 --csv --log-file gpurun_out/${TAG}_head_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity > /dev/null 2>&1; echo "head (one pass) rc $?"
fi
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_kernel<\(int\)16' -s 50 -c 1 --csv --log-file gpurun_out/${TAG}_tail_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline --no-parity > /dev/null 2>&1; echo "tail rc $?"
python tools/metrics_to_traffic.py ${TAG}
cat gpurun_out/${TAG}_n256_metrics.csv | cut -d, -f5,13- | tail -12; cat gpurun_out/${TAG}_head_metrics.csv | cut -d, -f13- | tail -10
