"""profiles/<tag>_traffic.json from an `ncu --set full` report of one conv launch: DRAM bytes of the launch
(dram__bytes_read.sum + dram__bytes_write.sum) beside its algorithmic bytes, stamped with the digest of the build that
was profiled. bench.py picks the newest such file for `roofline.traffic`.

    python tools/ncu_traffic.py <report.ncu-rep> <tag> "<kernel description>" <algorithmic bytes per launch>
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, tag, desc, algo = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    if rep.endswith(".csv"):          # already exported with `ncu -i <rep> --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def to_bytes(name):
        u, v = col[name]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v.replace(",", "")) * scale

    def num(name):
        return float(col[name][1].replace(",", "")) if name in col and col[name][1] else None

    dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    stamp = os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200", "b200sr3", "libb200sr3.so.stamp")
    out = {
        "kernel": desc, "kernel_name": col.get("Kernel Name", ("", ""))[1][:120],
        "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": algo,
        "l2_to_sm_tma_bytes_per_launch": to_bytes("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum")
        if "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum" in col else None,
        "duration_us_under_ncu": num("gpu__time_duration.sum"),
        "tensor_pipe_pct_of_active": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "smem_tc_wavefronts_pct": num("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "smem_lsu_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "build_digest": open(stamp).read().strip() if os.path.exists(stamp) else None,
        "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}",
    }
    path = os.path.join(ROOT, "profiles", f"{tag}_traffic.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
