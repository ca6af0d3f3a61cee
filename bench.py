"""Benchmark of the SR3 sampling hot path (BASELINE.json metric: SR3 16->128 faces/sec, full
sampling loop; PSNR vs ref).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full reverse-diffusion chain (all T timesteps) over one batch of synthetic
conditioning images of the named BASELINE config (default: the headline sr_sr3_VGGF2_16_128_model3,
T=600, 32 faces per GPU - 256 faces over 8 GPUs), weak scaling, batch sharded with no per-step
communication and one final gather (D2H concat into a pinned host buffer all ranks map; no NCCL
anywhere: torch.distributed runs on gloo and only carries the barrier and the timing reduction).
Prints ONE JSON line on rank 0.
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)

# BASELINE.json configs -> faces per GPU, the parity fixture produced by the reference (if any)
CONFIGS = {
    "sr_sr3_VGGF2_8_32_model2": {"faces_per_gpu": 4, "golden": None},
    "sr_sr3_VGGF2_16_64_model3": {"faces_per_gpu": 64, "golden": "chain_r64_T200.npz"},
    "sr_sr3_VGGF2_16_128_model3": {"faces_per_gpu": 32, "golden": "chain_r128_T600.npz"},
    "sr_sr3_VGGF2_8_128_model3": {"faces_per_gpu": 32, "golden": None},
    "sr_sr3_VGGF2_32_128_model2": {"faces_per_gpu": 64, "golden": None, "mica_handoff": True},
}
HEADLINE = "sr_sr3_VGGF2_16_128_model3"
# algorithmic FLOPs per image per diffusion step on the UN-OPTIMISED reference graph (2*MAC of every
# Conv2d + Linear at the reference's shapes; Upsample convs at full output resolution). The contract figure.
GFLOP_PER_IMG_STEP = {32: 5.5629, 64: 22.2483, 128: 88.9896}


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
        sm, power, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [v.strip() for v in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                out["sm_max_mhz"] = float(f[1])
                try:
                    power.append(float(f[2]))
                except ValueError:
                    pass
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_min_mhz"] = sm[0]
            out["samples"] = len(sm)
        if power:
            power.sort()
            out["power_w_median"] = power[len(power) // 2]
        out["reasons"] = sorted(reasons)
        return out


def workload(name, batch=None):
    import b200sr3
    opt = b200sr3.configs.named(name)
    mopt = opt["sr"]["model"]
    spec = CONFIGS[name]
    return {"name": name, "opt": opt, "mopt": mopt, "R": opt["r_resolution"], "L": opt["l_resolution"],
            "T": mopt["beta_schedule"]["val"]["n_timestep"], "B": int(batch or spec["faces_per_gpu"]),
            "golden": spec.get("golden"), "mica": bool(spec.get("mica_handoff"))}


def metric_name(w):
    return f"SR3 {w['L']}->{w['R']} faces/sec (full sampling loop)"


def cpu_reference_arm(w, steps, warmup, sample_B=2):
    """The reference's CPU implementation of the path (oracle port, torch fp32, all host threads)
    on a bounded sample: `steps` timed diffusion steps at B=sample_B of the config's R, extrapolated
    linearly in T to faces/sec. /root/reference does not exist on the GPU box, so the port in
    oracle/ (pinned against the reference by oracle/make_golden*.py) is what runs."""
    import torch
    from oracle import sr3_oracle as O
    from oracle.weights import make_inputs, make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    R, T, mopt = w["R"], w["T"], w["mopt"]
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


def torch_gpu_baseline(w, dev, steps=5, warmup=2):
    """Informational ("the bar is the same ops under the box's torch + cuDNN/cuBLAS Blackwell kernels"):
    the oracle port moved to the GPU - exactly what the reference module does after `.cuda()` - timed on the same B200
    at the benchmark batch for a few diffusion steps (all steps cost the same), fp32 with TF32 convs and
    bf16 autocast, extrapolated linearly in T. A checker leg like cpu_baseline: never on the product path."""
    import torch
    from oracle import sr3_oracle as O
    from oracle.weights import make_inputs, make_state_dict
    R, T, mopt, B = w["R"], w["T"], w["mopt"], w["B"]
    sd = {k: v.to(dev) for k, v in make_state_dict(mopt, seed=0, gain=1.0).items()}
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    cond, noise = make_inputs(B, R, 2, seed=123)
    cond, x0, z = cond.to(dev), noise[0].to(dev), noise[1].to(dev)
    out = {"what": f"oracle port (= the reference's torch ops) on this GPU, B={B}, R={R}, {steps} diffusion steps "
                   f"(+{warmup} warm-up) extrapolated linearly in T={T}", "torch": torch.__version__}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    for mode in ("fp32_tf32", "bf16_autocast"):
        try:
            x = x0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                for i in range(warmup + steps):
                    if i == warmup:
                        e0.record()
                    x = O.p_sample(sd, mopt, tabs, x.float(), T - 1 - i, cond, z)
                e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"ms_per_diffusion_step": ms, "faces_per_s": B / (ms * 1e-3 * T)}
        except Exception as e:      # never let the informational leg break the bench line
            out[mode] = {"error": str(e)[:200]}
    del sd
    torch.cuda.empty_cache()
    return out


def headline_parity(w, net, dev):
    """"PSNR vs ref" (BASELINE.json metric): the final image of the full-T chain against the one the unmodified
    reference produced (tests/golden/<fixture>, oracle/make_golden_headline.py), the reference's noise list injected
    for the golden faces, which sit in the middle of a benchmark-sized batch so the chain is the one timed above.
    uint8 PSNR as core/metrics.py:16-42,74-81 computes it (tensor2img on the device)."""
    import numpy as np
    import torch
    from b200sr3 import mica_handoff, synthetic
    path = os.path.join(ROOT, "tests", "golden", w["golden"])
    g = np.load(path)
    T, R, Bg = int(g["T"]), int(g["R"]), int(g["B"])
    assert T == w["T"] and R == w["R"]
    cond, noise = synthetic.inputs(Bg, R, T, seed=int(g["input_seed"]))      # the PCG64 stream the fixture was made from
    assert np.array_equal(cond.numpy(), g["cond"])
    batch = max(w["B"], Bg + 2)
    row = batch // 2 + 1
    gen = torch.Generator(device=dev).manual_seed(9)
    big_noise = torch.randn((T, batch, 3, R, R), generator=gen, device=dev)
    big_noise[:, row:row + Bg] = noise.to(dev)
    big_cond = torch.rand((batch, 3, R, R), generator=gen, device=dev) * 2 - 1
    big_cond[row:row + Bg] = cond.to(dev)
    out = net.super_resolution_batched(big_cond, noise=big_noise)[row:row + Bg]
    del big_noise
    ref = torch.from_numpy(g["final"]).to(dev)
    a = mica_handoff.tensor2img(out).double()
    b = mica_handoff.tensor2img(ref).double()
    mse = ((a - b) ** 2).flatten(1).mean(1)
    psnr = float((20 * torch.log10(255.0 / mse.clamp_min(1e-12).sqrt())).min())
    return {"psnr_vs_ref_db": psnr, "final_max_abs": float((out - ref).abs().max()),
            "parity_fixture": f"tests/golden/{w['golden']} (unmodified reference, T={T}, R={R}, injected noise, "
                              f"golden faces at rows {row}..{row + Bg - 1} of a B={batch} batch)"}


def newest_traffic(build_digest):
    """DRAM bytes of the dominant kernel class from the newest committed `ncu --set full` capture
    (profiles/*_traffic.json, written by tools/ncu_traffic.py); says whether it was taken on this build."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        tj = json.load(f)
    same = tj.get("build_digest") == build_digest
    note = (f"{tj['kernel']}: {tj['dram_bytes_per_launch'] / 1e6:.1f} MB DRAM per launch against "
            f"{tj['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic ({os.path.basename(files[-1])}, "
            f"{'this build' if same else 'capture of build ' + str(tj.get('build_digest'))[:12]}); "
            "`achieved` aggregates all conv launches of a step")
    return tj["dram_bytes_per_launch"], note


def build_digest():
    try:
        from b200sr3 import _lib
        with open(_lib.LIB_PATH + ".stamp") as f:
            return f.read().strip()
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=HEADLINE, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="faces per GPU (default: the config's, SURVEY.md 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = workload(args.config, args.batch)
    R, T, B = w["R"], w["T"], w["B"]
    config = {"workload": w["name"], "l_resolution": w["L"], "r_resolution": R, "n_timestep": T,
              "faces_per_gpu": B, "global_batch": B * world,
              "parallelism": f"batch-sharded x{world}, no per-step collective, one final gather "
                             "(D2H concat into a shared pinned host buffer; no NCCL)",
              "l2": "per-step working set (GBs of activations) far exceeds the 126 MB L2; no flush needed",
              "weights": "synthetic (numpy PCG64 seed 0), default-init scale"}
    if w["mica"]:
        config["tail_stage"] = ("each chain is followed by the SR->MICA hand-off kernels (tensor2img, resize 224, ArcFace blob), "
                                "the ArcFace iResNet-100, F.normalize and the MappingNetwork regressor (identity + shape code)")

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(w, max(args.steps, 1), max(args.warmup, 1))
        line = {"impl": "reference", "metric": metric_name(w), "value": cb["value"], "unit": "faces/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * cb["s_per_diffusion_step"] * T, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "steps_meaning": "diffusion steps at B=2 (a bounded sample), extrapolated in T; the b200 arm's steps "
                                 "are full chains"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import b200sr3
    from b200sr3.sharding import HostGather, shard_bounds
    from b200sr3 import mica_handoff, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")          # control plane only: barrier + timing reduction (CPU tensors)

    net = b200sr3.define_G(w["opt"])
    net.load_state_dict(synthetic.state_dict(net, seed=0, gain=1.0), strict=True)
    net = net.to(dev).eval()
    net.set_new_noise_schedule(w["mopt"]["beta_schedule"]["val"], [dev])

    glob_b = B * world
    cond_all = synthetic.inputs(glob_b, R, seed=123)
    lo, hi = shard_bounds(glob_b, rank, world)
    cond_host = cond_all[lo:hi].contiguous().pin_memory()
    cond = cond_host.to(dev)
    gather = HostGather((glob_b, 3, R, R))       # the one final gather lands here (pinned, mapped by every rank)
    mica = None
    if w["mica"]:
        # config 5: "SR output feeding the MICA/FLAME recon stage end-to-end" - hand-off kernels, ArcFace iResNet-100,
        # F.normalize and the MappingNetwork regressor on the device (the FLAME decoder needs licensed assets)
        mica = b200sr3.MicaEncoder()
        mica.arcface.load_state_dict(synthetic.mica_state_dict(mica.arcface, seed=1), strict=True)
        mica.regressor.load_state_dict(synthetic.mica_state_dict(mica.regressor, seed=2), strict=True)
        mica = mica.to(dev).eval()
    mica_launches = [0]

    def recon_stage(sr):
        h = mica_handoff.sr_to_mica(sr)
        ident, shape = mica(h["arcface"])
        mica_launches[0] = 2 + 107
        return shape

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_chain(i):
        # one seed for the whole (global) batch; rows are told apart by the global row offset
        out = net.super_resolution_batched(cond, seed=1000 + i, row_offset=lo)
        gather.put(lo, out)                      # async D2H of this rank's slice: the final gather, no rendezvous
        if w["mica"]:
            recon_stage(out)
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
        launches += net.launch_counts()[0] + mica_launches[0]
    e1.record()
    barrier()
    mine = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop() if rank == 0 else None
    per_rank = [mine]
    if world > 1:
        buf = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(buf, torch.tensor([mine], dtype=torch.float64))
        per_rank = [float(b.item()) for b in buf]
    elapsed = max(per_rank)
    finite = bool(torch.isfinite(gather.full()).all())

    # ---- end to end through the public API with HOST buffers (H2D + chain + D2H each step); the output lands
    # directly in this rank's rows of the shared gather buffer
    out_rows = gather.full()[lo:hi]
    net.sample_host(cond_host, out_rows, seed=7, row_offset=lo)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        net.sample_host(cond_host, out_rows, seed=200 + i, row_offset=lo)
        if w["mica"]:
            recon_stage(out_rows.to(dev, non_blocking=True)).cpu()      # the shape codes are what leaves the device
    torch.cuda.synchronize()
    e2e_mine = time.perf_counter() - t0
    e2e_s = e2e_mine
    if world > 1:
        buf = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(buf, torch.tensor([e2e_mine], dtype=torch.float64))
        e2e_s = max(float(b.item()) for b in buf)
    img_bytes = cond_host.numel() * 4

    if rank != 0:
        barrier()
        gather.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = glob_b * args.steps / elapsed
    # ---- roofline of the dominant kernel class (tcgen05 convs), measured live: one eager step with a
    # CUDA event between launches gives every launch's device time.
    prof = net.profile_step(B, R)
    conv = [p for p in prof if p[2] > 0]
    conv_ms = sum(p[1] for p in conv)
    step_ms = sum(p[1] for p in prof)
    flops_contract = B * GFLOP_PER_IMG_STEP[R] * 1e9         # B x F(R), exactly
    flops_ops = sum(p[2] for p in conv)                      # the same graph counted launch by launch (convs only)
    flops_executed = sum(p[4] for p in conv)
    # `achieved`: the step's algorithmic FLOPs / the conv launches' duration in the TIMED region. The timed region replays
    # one CUDA graph per sampling step (no per-launch events possible inside it), so that duration is the graph-replayed
    # step scaled by the convs' share of the eager per-launch profile of the same step (kept beside it: its absolute
    # times carry ~1-2 us of event/launch gap per launch).
    graph_step_ms = 1e3 * elapsed / args.steps / T
    conv_share = conv_ms / step_ms
    achieved = flops_contract / (graph_step_ms * conv_share * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    traffic, traffic_note = newest_traffic(build_digest())
    roofline = {"bound": "tensor",
                "kernel": "conv_halo_kernel (+ conv_umma_kernel for 1x1 attention convs): all tcgen05 conv launches of one sampling step",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a long step)",
                "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
                "launches_per_step": len(conv), "flops_per_step": flops_contract,
                "flops_per_step_note": "B x F(R) of SURVEY.md 8(d), reference graph; launch-by-launch recount of the convs "
                                       f"{flops_ops:.6e} (differs only by the Linear layers of the noise MLP)",
                "flops_executed_per_step": flops_executed,
                "conv_ms_per_step": graph_step_ms * conv_share, "conv_share_of_step": conv_share,
                "eager_profile": {"conv_ms_per_step": conv_ms, "step_ms": step_ms,
                                  "achieved": flops_contract / (conv_ms * 1e-3) / 1e12},
                "whole_step_frac": value / world * T * GFLOP_PER_IMG_STEP[R] * 1e9 / (peak * 1e12)}
    # ---- the HBM-bound ends of the chain, one entry per kernel (achieved GB/s against the measured copy bandwidth)
    groups = {}
    for name, ms, fl, by, fx in prof:
        key = "downs.0" if name.startswith("downs.0.") else name      # the head's operand pack belongs to the head
        g = groups.setdefault(key, [0.0, 0.0])
        g[0] += ms
        g[1] += by
    hbm = []
    for name, (ms, by) in groups.items():
        if by > 0:
            gbs = by / (ms * 1e-3) / 1e9
            hbm.append({"kernel": name, "us": ms * 1e3, "algorithmic_mb": by / 1e6, "achieved": gbs,
                        "frac": gbs / peaks["hbm_gbs"], "share_of_step": ms / step_ms})
    hbm_traffic = None
    try:      # DRAM bytes of the head and tail launches from the newest committed ncu capture (tools/profile_hbm_ends.sh)
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic_hbm.json")))
        with open(files[-1]) as f:
            tj = json.load(f)
        hbm_traffic = {"file": os.path.basename(files[-1]), "this_build": tj.get("build_digest") == build_digest(),
                       "dram_bytes_per_launch": {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}}
    except Exception:
        pass
    hbm_ms = sum(h["us"] for h in hbm) / 1e3
    hbm_bytes = sum(h["algorithmic_mb"] for h in hbm) * 1e6
    roofline_hbm = {"bound": "hbm", "kernel": "head (downs.0: fp32 NCHW cond/x -> 64-ch bf16) and tail (final_conv + sampler "
                                              "update); plus any stand-alone GroupNorm pass the config still has", "unit": "GB/s", "peak": peaks["hbm_gbs"],
                    "achieved": hbm_bytes / (hbm_ms * 1e-3) / 1e9 if hbm_ms else None,
                    "frac": hbm_bytes / (hbm_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if hbm_ms else None,
                    "ms_per_step": hbm_ms, "share_of_step": hbm_ms / step_ms, "traffic": hbm_traffic, "kernels": hbm}
    ranks_ms = sorted(1e3 * s / args.steps for s in per_rank)
    line = {"metric": metric_name(w), "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "ms_per_diffusion_step": 1e3 * elapsed / args.steps / T,
            "per_rank_ms_per_step": {"min": ranks_ms[0], "median": ranks_ms[len(ranks_ms) // 2], "max": ranks_ms[-1]},
            "e2e": {"value": glob_b * args.steps / e2e_s, "unit": "faces/s", "h2d_bytes_per_step": img_bytes,
                    "d2h_bytes_per_step": img_bytes},
            "gpu_launches": launches, "outputs_finite": finite, "clocks": clocks, "roofline": roofline,
            "roofline_hbm": roofline_hbm, "build_digest": build_digest()}
    if w["golden"] and not args.no_parity:
        try:
            line.update(headline_parity(w, net, dev))
        except Exception as e:
            line["psnr_vs_ref_db"] = None
            line["parity_error"] = str(e)[:300]
    if world == 1 and not args.no_torch_baseline:
        del net
        torch.cuda.empty_cache()
        line["torch_gpu_baseline"] = torch_gpu_baseline(w, dev)
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_arm(w, 10, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    barrier()
    gather.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
