"""Probe: does running the batch as TWO independent half-batches on two CUDA streams (two engines, two host
threads) beat one launch chain? Each sampling step is a chain of ~70 kernels with a global barrier between
them; a second independent chain can fill the SMs that idle in the first one's wave tails and fill/drain phases.

    python tools/dual_stream_probe.py [B] [T]
"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
import torch
import b200sr3
from b200sr3 import synthetic


def make(T):
    opt = b200sr3.configs.named("sr_sr3_VGGF2_16_128_model3")
    mopt = opt["sr"]["model"]
    mopt["beta_schedule"]["val"]["n_timestep"] = T
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic.state_dict(net, seed=0, gain=1.0), strict=True)
    net = net.to("cuda").eval()
    net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device("cuda")])
    return net


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    ways = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    cond = synthetic.inputs(B, 128, seed=123).cuda()
    one = make(T)
    one.super_resolution_batched(cond, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one.super_resolution_batched(cond, seed=2)
    torch.cuda.synchronize()
    t1 = time.perf_counter() - t0
    print(f"one chain  B={B}: {t1 / T * 1e3:.3f} ms per sampling step, {B / t1 * T / 600:.2f} faces/s at T=600")
    nets = [one] + [make(T) for _ in range(ways - 1)]
    streams = [torch.cuda.Stream() for _ in range(ways)]
    parts = list(cond.chunk(ways))

    def run(i, seed):
        with torch.cuda.stream(streams[i]):
            nets[i].super_resolution_batched(parts[i], seed=seed)

    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(i, 10 + rep)) for i in range(ways)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        t2 = time.perf_counter() - t0
    print(f"{ways} chains x B={B // ways}: {t2 / T * 1e3:.3f} ms per sampling step, {B / t2 * T / 600:.2f} faces/s at T=600")


if __name__ == "__main__":
    main()
