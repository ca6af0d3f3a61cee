"""gpurun_out/<tag>_{n256,head,tail}_metrics.csv (tools/profile_metrics_fallback.sh) -> gpurun_out/<tag>_traffic.json and
<tag>_traffic_hbm.json, the files bench.py copies into roofline.traffic / roofline_hbm.traffic once they sit in profiles/."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "gpurun_out")
digest = open(os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200", "b200sr3", "libb200sr3.so.stamp")).read().strip()


def metrics(path):
    m = {}
    for r in csv.reader(open(path)):
        if len(r) >= 3:
            try:
                m[r[-3]] = float(r[-1].replace(",", ""))
            except ValueError:
                pass
    return m


n = metrics(os.path.join(out, f"{tag}_n256_metrics.csv"))
src = f"ncu --metrics <list> --clock-control none (tools/profile_metrics_fallback.sh, profiles/{tag}_*_metrics.csv)"
json.dump({
    "kernel": "conv_halo_kernel<256,1,gn,geo0> (ups.8.conv1: 512+256->256 3x3 @32x32, B=32 - the kernel class with the largest share of a step)",
    "dram_bytes_per_launch": n["dram__bytes_read.sum"] + n["dram__bytes_write.sum"],
    "algorithmic_bytes_per_launch": 70600000.0,
    "l2_to_sm_bytes_per_launch": n["l1tex__m_xbar2l1tex_read_bytes.sum"],
    "duration_us_under_ncu": n["gpu__time_duration.sum"] / 1e3,
    "tensor_pipe_pct_of_active": n["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"],
    "smem_tc_wavefronts_pct": n["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"],
    "smem_lsu_wavefronts_pct": n["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"],
    "l2_hit_rate_pct": n["lts__t_sector_hit_rate.pct"], "build_digest": digest, "source": src}, open(os.path.join(out, f"{tag}_traffic.json"), "w"), indent=1)
kernels = {}
for name, key, alg in (("downs.0", "head", 79691776.0), ("final_conv.tail", "tail", 79691776.0)):
    path = os.path.join(out, f"{tag}_{key}_metrics.csv")
    if not os.path.exists(path):
        continue
    h = metrics(path)
    if "dram__bytes_read.sum" not in h:
        continue
    kernels[name] = {"dram_bytes_per_launch": h["dram__bytes_read.sum"] + h["dram__bytes_write.sum"],
                     "dram_read": h["dram__bytes_read.sum"], "dram_write": h["dram__bytes_write.sum"],
                     "algorithmic_bytes_per_launch": alg, "duration_us_under_ncu": h["gpu__time_duration.sum"] / 1e3,
                     "l2_to_sm_bytes_per_launch": h.get("l1tex__m_xbar2l1tex_read_bytes.sum")}
json.dump({"kernels": kernels, "build_digest": digest, "source": src + "; launches inside a real chain (bench.py --steps 1)"},
          open(os.path.join(out, f"{tag}_traffic_hbm.json"), "w"), indent=1)
print(json.dumps(kernels, indent=1))
