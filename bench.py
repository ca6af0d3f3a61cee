"""Benchmark of the SR3 sampling hot path (BASELINE.json metric: SR3 16->128 faces/sec, full
sampling loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full reverse-diffusion chain (all T=600 timesteps) over one batch of synthetic
16->128 conditioning images: 32 faces per GPU (config sr_sr3_VGGF2_16_128_model3 is 256 faces
over 8 GPUs), weak scaling, batch sharded with no per-step communication and one final gather.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)

CONFIG = "sr_sr3_VGGF2_16_128_model3"
R, T, PER_GPU_BATCH = 128, 600, 32
GFLOP_PER_IMG_STEP = {32: 5.5629, 64: 22.2483, 128: 88.9896}   # reference graph, 2*MAC
METRIC = "SR3 16->128 faces/sec (full sampling loop)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [v.strip() for v in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                out["sm_max_mhz"] = float(f[1])
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_reference_arm(steps, warmup, sample_B=2):
    """The reference's CPU implementation of the path (oracle port, torch fp32, all host threads)
    on a bounded sample: `steps` timed diffusion steps at B=sample_B, R=128, extrapolated
    linearly in T to faces/sec. /root/reference does not exist on the GPU box, so the port in
    oracle/ (pinned against the reference by oracle/make_golden.py) is what runs."""
    import torch
    from oracle import sr3_oracle as O
    from oracle.weights import make_inputs, make_state_dict
    import b200sr3
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mopt = b200sr3.configs.named(CONFIG)["sr"]["model"]
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    cond, noise = make_inputs(sample_B, R, 2, seed=123)
    x = noise[0]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x = O.p_sample(sd, mopt, tabs, x, T - 1 - (i % T), cond, noise[1])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    s_per_step = sum(times) / len(times)
    faces_per_s = sample_B / (s_per_step * T)
    return {"value": faces_per_s, "unit": "faces/s", "cores": cores, "kind": "port",
            "sample": f"{steps} of {T} diffusion steps at B={sample_B}, R={R} (+{warmup} warm-up), "
                      f"{s_per_step:.3f} s/step, extrapolated linearly in T",
            "s_per_diffusion_step": s_per_step}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="faces per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": CONFIG, "l_resolution": 16, "r_resolution": R, "n_timestep": T,
              "faces_per_gpu": args.batch, "global_batch": args.batch * world,
              "parallelism": f"batch-sharded x{world}, no per-step collective, one final gather",
              "l2": "per-step working set (GBs of activations) far exceeds the 126 MB L2; no flush needed",
              "weights": "synthetic (numpy PCG64 seed 0), default-init scale"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(max(args.steps, 1), max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "faces/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cb["s_per_diffusion_step"] * T,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import b200sr3
    from b200sr3.sharding import shard_bounds
    from b200sr3 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    opt = b200sr3.configs.named(CONFIG)
    mopt = opt["sr"]["model"]
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic.state_dict(net, seed=0, gain=1.0), strict=True)
    net = net.to(dev).eval()
    net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [dev])

    B = args.batch
    glob = B * world
    cond_all = synthetic.inputs(glob, R, seed=123)
    lo, hi = shard_bounds(glob, rank, world)
    cond_host = cond_all[lo:hi].contiguous().pin_memory()
    out_host = torch.empty_like(cond_host).pin_memory()
    cond = cond_host.to(dev)
    gathered = [torch.empty_like(cond) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_chain(i):
        out = net.super_resolution_batched(cond, seed=1000 + i)
        if world > 1:
            dist.all_gather(gathered, out)            # the one final gather
        return out

    for i in range(args.warmup):
        one_chain(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for i in range(args.steps):
        one_chain(100 + i)
        launches += net.launch_counts()[0]
    e1.record()
    barrier()
    elapsed = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())

    # ---- end to end through the public API with HOST buffers (H2D + chain + D2H each step)
    net.sample_host(cond_host, out_host, seed=7)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        net.sample_host(cond_host, out_host, seed=200 + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    img_bytes = cond_host.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = glob * args.steps / elapsed
    # ---- roofline of the dominant kernel (tcgen05 conv), measured live: one eager step with a
    # CUDA event between launches gives every launch's device time.
    prof = net.profile_step(B, R)
    conv = [(n, ms, fl) for n, ms, fl, by in prof if fl > 0]
    conv_ms = sum(ms for _, ms, _ in conv)
    conv_flops = sum(fl for _, _, fl in conv)
    step_ms = sum(ms for _, ms, _, _ in prof)
    gn = [(ms, by) for n, ms, fl, by in prof if by > 0]
    gn_ms, gn_bytes = sum(m for m, _ in gn), sum(b for _, b in gn)
    # `achieved`: the conv launches' FLOPs / their duration in the TIMED region. The timed region replays one CUDA graph
    # per sampling step (no per-launch events possible inside it), so the duration is the graph-replayed step scaled by
    # the convs' share of the eager per-launch profile of the same step (the eager profile itself is kept beside it:
    # its absolute times carry ~1-2 us of event/launch gap per launch).
    graph_step_ms = 1e3 * elapsed / args.steps / T
    conv_share = conv_ms / step_ms
    achieved = conv_flops / (graph_step_ms * conv_share * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    traffic, traffic_note = None, None
    try:      # DRAM bytes of the dominant kernel class from the committed `ncu --set full` capture (per launch)
        with open(os.path.join(ROOT, "profiles", "r01f_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_launch"]
        traffic_note = (f"{tj['kernel']}: {tj['dram_bytes_per_launch'] / 1e6:.1f} MB DRAM per launch against "
                        f"{tj['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic ({tj['source']}); "
                        "`achieved` above aggregates all 66 conv launches of a step")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "conv_halo_kernel + conv_umma_kernel (all tcgen05 conv launches of one sampling step)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a long step)",
                "launches_per_step": len(conv), "flops_per_step": conv_flops,
                "conv_ms_per_step": graph_step_ms * conv_share, "conv_share_of_step": conv_share,
                "eager_profile": {"conv_ms_per_step": conv_ms, "step_ms": step_ms,
                                  "achieved": conv_flops / (conv_ms * 1e-3) / 1e12},
                "whole_step_frac": value / world * T * GFLOP_PER_IMG_STEP[R] * 1e9 / (peak * 1e12)}
    roofline_hbm = {"bound": "hbm", "kernel": "gn_apply_kernel (the GroupNorm passes that are not fused into a conv)",
                    "achieved": gn_bytes / (gn_ms * 1e-3) / 1e9,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gn_bytes / (gn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "ms_per_step": gn_ms, "share_of_step": gn_ms / step_ms, "traffic": None}
    line = {"metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "ms_per_diffusion_step": 1e3 * elapsed / args.steps / T,
            "e2e": {"value": glob * args.steps / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": img_bytes,
                    "d2h_bytes_per_step": img_bytes},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm}
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_arm(10, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
