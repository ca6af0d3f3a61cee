"""Drop-in for the reference GaussianDiffusion (model/sr/sr3_modules/diffusion.py:65-313).

Same constructor, attributes, buffers and method names; the reverse-diffusion path
(p_sample / p_sample_loop / sample / super_resolution, diffusion.py:164-225) runs in
libb200sr3 on a B200. There is no PyTorch or CPU fallback for it: without the library or an
sm_100 device these methods raise. The training loss (forward / p_losses, diffusion.py:275-313)
stays a differentiable torch path — it is outside the accelerated hot path.
"""
import ctypes as C
from functools import partial

import numpy as np
import torch
from torch import nn

from . import _lib


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 betas for the schedule names the reference accepts (diffusion.py:20-50)."""
    lin = partial(np.linspace, num=n_timestep, dtype=np.float64)
    if schedule == "linear":
        return lin(linear_start, linear_end)
    if schedule == "quad":
        return lin(linear_start ** 0.5, linear_end ** 0.5) ** 2
    if schedule == "const":
        return np.full(n_timestep, linear_end, dtype=np.float64)
    if schedule in ("warmup10", "warmup50"):
        betas = np.full(n_timestep, linear_end, dtype=np.float64)
        n = int(n_timestep * (0.1 if schedule == "warmup10" else 0.5))
        betas[:n] = np.linspace(linear_start, linear_end, n, dtype=np.float64)
        return betas
    if schedule == "jsd":
        return 1.0 / lin(n_timestep, 1)
    if schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        alphas = alphas / alphas[0]
        return (1 - alphas[1:] / alphas[:-1]).clamp(max=0.999).numpy()
    raise NotImplementedError(schedule)


class _Engine:
    """Owns one b200sr3_handle (one device)."""

    def __init__(self, cfg, device_index):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        _lib.check(self.lib.b200sr3_create(C.byref(cfg), int(device_index), C.byref(self.handle)))
        self.device_index = int(device_index)
        self.weights_version = None
        self.schedule_version = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.b200sr3_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


class GaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, image_size, channels=3, loss_type="l1", conditional=True, schedule_opt=None):
        super().__init__()
        self.channels = channels
        self.image_size = image_size
        self.denoise_fn = denoise_fn
        self.loss_type = loss_type
        self.conditional = conditional
        # like the reference (diffusion.py:81-83) the schedule is NOT installed here
        self._cfg = None          # _lib.Config, set by define_G
        self._engines = {}
        self._sched_host = None
        self._sched_serial = 0
        self.noise_seed = None    # fixed seed of the in-kernel Philox stream; None: drawn from torch's global generator
        self._weights_serial = 0  # bumped by invalidate_weights(): the engines repack before their next call

    def __getstate__(self):       # engines hold device handles: never pickled / deep-copied
        state = self.__dict__.copy()
        state["_engines"] = {}
        return state

    # ------------------------------------------------------------------ reference surface
    def set_loss(self, device):
        if self.loss_type == "l1":
            self.loss_func = nn.L1Loss(reduction="sum").to(device)
        elif self.loss_type == "l2":
            self.loss_func = nn.MSELoss(reduction="sum").to(device)
        else:
            raise NotImplementedError()

    def set_new_noise_schedule(self, schedule_opt, device):
        """diffusion.py:93-142, including its `device` convention (0, or a list whose [0] is used)."""
        if device != 0:
            device = device[0]
        to_torch = partial(torch.tensor, dtype=torch.float32, device=device)
        betas = make_beta_schedule(schedule_opt["schedule"], schedule_opt["n_timestep"],
                                   schedule_opt["linear_start"], schedule_opt["linear_end"])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.sqrt_alphas_cumprod_prev = np.sqrt(np.append(1.0, ac))
        self.num_timesteps = int(betas.shape[0])
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        tables = {
            "betas": betas,
            "alphas_cumprod": ac,
            "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": np.sqrt(ac),
            "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
            "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
            "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
            "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
            "posterior_variance": post_var,
            "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
            "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
            "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        }
        for name, arr in tables.items():
            self.register_buffer(name, to_torch(arr))
        f32 = lambda k: np.ascontiguousarray(tables[k], dtype=np.float32)
        self._sched_host = (f32("sqrt_recip_alphas_cumprod"), f32("sqrt_recipm1_alphas_cumprod"),
                            f32("posterior_mean_coef1"), f32("posterior_mean_coef2"),
                            f32("posterior_log_variance_clipped"),
                            np.ascontiguousarray(self.sqrt_alphas_cumprod_prev, dtype=np.float64))
        self._sched_serial += 1

    # ------------------------------------------------------------------ engine plumbing
    def invalidate_weights(self):
        """Tell the engines that parameter VALUES changed, so the bf16 operands are repacked before the next sampling
        call. In-place edits through autograd-visible ops (`p.copy_`, `p.zero_` under no_grad, optimizer steps) and
        storage changes (`.to()`, `.cuda()`, `load_state_dict`) are detected automatically; edits through `p.data`
        (the reference's finetune_norm `v.data.zero_()`, model/sr/model.py:45, and EMA-style `p.data.copy_()`) bump a
        separate version counter torch does not expose on the parameter - call this after them."""
        self._weights_serial += 1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _next_seed(self, seed):
        """Seed of the Philox stream for one sampling call. Like the reference, whose noise comes from torch's global
        generator (diffusion.py:186,205), the default is drawn from that generator: torch.manual_seed() makes runs
        reproducible, and ranks seeded alike agree on the stream (rows are told apart by `row_offset`)."""
        if seed is not None:
            return int(seed) & (2 ** 64 - 1)
        if self.noise_seed is not None:
            return int(self.noise_seed) & (2 ** 64 - 1)
        return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())

    def _sampling_device(self):
        dev = self.betas.device if hasattr(self, "betas") else next(self.denoise_fn.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError(
                f"b200sr3: sampling needs the model on a CUDA (sm_100a) device, found '{dev}'. "
                "There is no CPU fallback; use the reference module on CPU.")
        return dev

    def _engine(self, dev=None):
        dev = dev or self._sampling_device()
        if self._cfg is None:
            raise RuntimeError("b200sr3: GaussianDiffusion must be built by define_G (no engine config)")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        eng = self._engines.get(idx)
        if eng is None:
            eng = self._engines[idx] = _Engine(self._cfg, idx)
        params = self.denoise_fn.tensors()
        # staleness: storage identity and torch's version counters folded into two integers (plus the explicit serial)
        ptrs = vers = 0
        for _, p in params:
            ptrs ^= p.data_ptr()
            vers += p._version
        version = (self._weights_serial, ptrs, vers)
        if eng.weights_version != version:
            with torch.cuda.device(idx):
                staged = []
                for key, p in params:
                    t = p.detach()
                    if t.device.type != "cuda" or t.device.index != idx or t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.to(device=f"cuda:{idx}", dtype=torch.float32).contiguous()
                    staged.append((key, t))
                torch.cuda.current_stream().synchronize()      # once: the library copies on its own stream
                for key, t in staged:
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(eng.lib.b200sr3_load_tensor(eng.handle, key.encode(), _ptr(t), shape, t.dim()))
                del staged
                _lib.check(eng.lib.b200sr3_finalize_weights(eng.handle, _stream()))
            eng.weights_version = version
        if getattr(self, "num_timesteps", None) and eng.schedule_version != self._sched_serial:
            a, b, c1, c2, lv, sp = self._sched_host
            host = lambda arr: C.c_void_p(arr.ctypes.data)
            with torch.cuda.device(idx):
                _lib.check(eng.lib.b200sr3_set_schedule(eng.handle, self.num_timesteps, host(a), host(b), host(c1),
                                                        host(c2), host(lv), host(sp), _stream()))
            eng.schedule_version = self._sched_serial
        return eng

    @staticmethod
    def _as_input(t, dev):
        return t.detach().to(device=dev, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ accelerated path
    @torch.no_grad()
    def unet_eps(self, cond, x, noise_level):
        """denoise_fn(cat([cond, x], 1), noise_level) with one scalar level (diffusion.py:170)."""
        dev = self._sampling_device()
        eng = self._engine(dev)
        x = self._as_input(x, dev)
        cond = self._as_input(cond, dev) if cond is not None else None
        eps = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(eng.lib.b200sr3_unet_forward(eng.handle, _ptr(cond), _ptr(x), float(noise_level),
                                                    x.shape[0], x.shape[-1], _ptr(eps), _stream()))
        return eps

    @torch.no_grad()
    def p_sample(self, x, t, clip_denoised=True, condition_x=None, noise=None):
        """diffusion.py:182-187. `noise` injects z_t; when None it is drawn with torch.randn_like
        exactly where the reference draws it, so the global RNG stream is consumed identically."""
        dev = self._sampling_device()
        eng = self._engine(dev)
        x = self._as_input(x, dev)
        cond = self._as_input(condition_x, dev) if condition_x is not None else None
        if noise is None and t > 0:
            noise = torch.randn_like(x)
        z = self._as_input(noise, dev) if (noise is not None and t > 0) else None
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(eng.lib.b200sr3_step(eng.handle, _ptr(cond), _ptr(x), _ptr(z), int(t), 1 if clip_denoised else 0,
                                            x.shape[0], x.shape[-1], _ptr(out), _stream()))
        return out

    @torch.no_grad()
    def sample_batched(self, x_in, noise=None, seed=None, return_snapshots=False, row_offset=0, return_x_T=False):
        """All T steps for a whole batch: the entry point benchmarks and parity tests use.

        x_in: cond [B,3,R,R] (conditional) — or a shape tuple for unconditional models.
        noise: optional injected list [T,B,3,R,R] = [x_T, z_{T-1}, ..., z_1]; otherwise the
        in-kernel Philox stream with `seed` (default: `noise_seed`, else drawn from torch's global generator).
        row_offset: global index of batch row 0 in the Philox stream - shards / chunks of one logical batch pass
        their start row and the same seed, and every face gets the noise it would get in the unsplit batch.
        Returns x_0 [B,3,R,R] (then the `continous=True` snapshots [n,B,3,R,R] and / or x_T if asked).
        """
        dev = self._sampling_device()
        eng = self._engine(dev)
        if self.conditional:
            cond = self._as_input(x_in, dev)
            shape = tuple(cond.shape)
        else:
            cond, shape = None, tuple(x_in)
        B, R = shape[0], shape[-1]
        if shape[-2] != R:
            raise ValueError("b200sr3: only square images are supported")
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        snaps = None
        if return_snapshots:
            snaps = torch.empty((eng.lib.b200sr3_num_snapshots(eng.handle),) + shape, dtype=torch.float32, device=dev)
        if noise is not None:
            noise = self._as_input(noise, dev)
            if tuple(noise.shape) != (self.num_timesteps,) + shape:
                raise ValueError(f"b200sr3: injected noise must have shape {(self.num_timesteps,) + shape}")
            mode, sd = _lib.NOISE_INJECTED, 0
        else:
            mode, sd = _lib.NOISE_PHILOX, self._next_seed(seed)
        with torch.cuda.device(dev):
            _lib.check(eng.lib.b200sr3_sample(eng.handle, _ptr(cond), mode, _ptr(noise), C.c_uint64(sd),
                                              C.c_int64(int(row_offset)), B, R, _ptr(out), _ptr(snaps), _stream()))
        res = (out,) + ((snaps,) if return_snapshots else ())
        if return_x_T:
            if noise is not None:
                x_T = noise[0].clone()
            else:
                x_T = torch.empty(shape, dtype=torch.float32, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(eng.lib.b200sr3_philox_normal(eng.handle, C.c_uint64(sd), int(self.num_timesteps),
                                                             C.c_int64(int(row_offset)), B, R, _ptr(x_T), _stream()))
            res = res + (x_T,)
        return res if len(res) > 1 else out

    @torch.no_grad()
    def philox_normal(self, shape, t, seed, row_offset=0):
        """The sampler's own N(0,1) stream: the draw the Philox mode makes at key t (t = num_timesteps: x_T,
        diffusion.py:205; 0 < t < T: z_t, diffusion.py:186) for rows row_offset.. of the global batch."""
        dev = self._sampling_device()
        eng = self._engine(dev)
        out = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(eng.lib.b200sr3_philox_normal(eng.handle, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(t),
                                                     C.c_int64(int(row_offset)), out.shape[0], out.shape[-1], _ptr(out),
                                                     _stream()))
        return out

    @torch.no_grad()
    def p_sample_loop(self, x_in, continous=False, noise=None, seed=None):
        """diffusion.py:189-215, including its return convention: `continous=True` returns
        cat([x_in, snapshots...], 0); otherwise ONLY the last batch element, shape [3,R,R]."""
        if continous:
            # the list starts with x_in (conditional, diffusion.py:203-204) or with x_T (unconditional, :195-196)
            out, snaps, x_T = self.sample_batched(x_in, noise=noise, seed=seed, return_snapshots=True, return_x_T=True)
            head = self._as_input(x_in, out.device) if self.conditional else x_T
            flat = snaps.reshape((-1,) + tuple(snaps.shape[2:]))
            return torch.cat([head, flat], dim=0)
        return self.sample_batched(x_in, noise=noise, seed=seed)[-1]

    @torch.no_grad()
    def sample(self, batch_size=1, continous=False):
        return self.p_sample_loop((batch_size, self.channels, self.image_size, self.image_size), continous)

    @torch.no_grad()
    def super_resolution(self, x_in, continous=False, noise=None, seed=None):
        return self.p_sample_loop(x_in, continous, noise=noise, seed=seed)

    @torch.no_grad()
    def super_resolution_batched(self, x_in, noise=None, seed=None, row_offset=0):
        """[B,3,R,R] in -> [B,3,R,R] out (the reference returns only the last element)."""
        return self.sample_batched(x_in, noise=noise, seed=seed, row_offset=row_offset)

    @torch.no_grad()
    def super_resolution_samples(self, x_in, n_samples, noise=None, seed=None, max_batch=None):
        """`n_samples` independent chains per conditioning image in ONE batched call.

        The reference's evaluation loop draws `cfg.sample` (the -s flag, 15 in the paper's runs) chains for the same
        LR input one after the other at B=1 (lib/trainer_temp.py:441-444 calling model/sr3d/model.py:366 test_val,
        each a full super_resolution call). Chains are independent, so they stack along the batch: image i, sample k
        sits at row i*n_samples + k. x_in [B,3,R,R] -> [B, n_samples, 3, R, R]. `noise` (optional, parity runs):
        [T, B*n_samples, 3, R, R] in that row order; otherwise Philox with ONE `seed` for the whole call, keyed by the
        global row i*n_samples + k. `max_batch` bounds the rows per launch chain (the workspace is sized per batch);
        in both noise modes every row's result is independent of the chunking (chunks pass their start row as the
        stream's row offset)."""
        n = int(n_samples)
        if n < 1:
            raise ValueError("b200sr3: n_samples must be >= 1")
        dev = self._sampling_device()
        cond = self._as_input(x_in, dev)
        rows = cond.repeat_interleave(n, dim=0)
        total = rows.shape[0]
        step = total if not max_batch else max(1, int(max_batch))
        if noise is not None and step < total:
            noise = self._as_input(noise, dev)
        outs = []
        sd = None if noise is not None else self._next_seed(seed)
        for lo in range(0, total, step):
            hi = min(total, lo + step)
            nz = None if noise is None else noise[:, lo:hi].contiguous()
            outs.append(self.sample_batched(rows[lo:hi].contiguous(), noise=nz, seed=sd, row_offset=lo))
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        return out.reshape((cond.shape[0], n) + tuple(out.shape[1:]))

    @torch.no_grad()
    def sample_host(self, cond_host, out_host=None, seed=None, row_offset=0):
        """End-to-end call on HOST tensors (pinned for full PCIe speed): cond is copied to the
        device, the full chain runs with Philox noise (`seed` as in sample_batched: None draws it from
        torch's global generator), the result is copied back; returns when the copy has landed. This is
        the path bench.py reports as `e2e`."""
        dev = self._sampling_device()
        eng = self._engine(dev)
        if cond_host.device.type != "cpu" or cond_host.dtype != torch.float32 or not cond_host.is_contiguous():
            raise ValueError("b200sr3: sample_host expects a contiguous fp32 CPU tensor")
        if out_host is None:
            out_host = torch.empty_like(cond_host, pin_memory=True)
        B, R = cond_host.shape[0], cond_host.shape[-1]
        with torch.cuda.device(dev):
            _lib.check(eng.lib.b200sr3_sample_host(eng.handle, _ptr(cond_host), C.c_uint64(self._next_seed(seed)),
                                                   C.c_int64(int(row_offset)), B, R, _ptr(out_host), _stream()))
        return out_host

    def profile_step(self, B, R):
        """Per-launch device times of one sampling step: list of (name, ms, flops, bytes, flops_executed)."""
        eng = self._engine()
        n_max = 1024
        ms = (C.c_float * n_max)()
        fl = (C.c_double * n_max)()
        fx = (C.c_double * n_max)()
        by = (C.c_double * n_max)()
        names = C.create_string_buffer(64 * n_max)
        n = C.c_int()
        with torch.cuda.device(self._sampling_device()):
            _lib.check(eng.lib.b200sr3_profile_step(eng.handle, B, R, n_max, ms, fl, fx, by, names, len(names),
                                                    C.byref(n), _stream()))
        nm = names.value.decode().split("\n")
        return [(nm[i], ms[i], fl[i], by[i], fx[i]) for i in range(n.value)]

    def layer_output(self, name):
        """Activation of a UNet module ('downs.3', 'mid.0', ...) from the most recent forward."""
        eng = self._engine()
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        _lib.check(eng.lib.b200sr3_layer_output(eng.handle, name.encode(), None, C.byref(c), C.byref(h), C.byref(w), _stream()))
        return c.value, h.value, w.value

    def launch_counts(self):
        eng = self._engine()
        total, conv = C.c_int64(), C.c_int64()
        _lib.check(eng.lib.b200sr3_last_launch_count(eng.handle, C.byref(total), C.byref(conv)))
        return total.value, conv.value

    # ------------------------------------------------------------------ training loss (torch)
    def q_sample(self, x_start, continuous_sqrt_alpha_cumprod, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return continuous_sqrt_alpha_cumprod * x_start + (1 - continuous_sqrt_alpha_cumprod ** 2).sqrt() * noise

    def p_losses(self, x_in, noise=None):
        """diffusion.py:284-313: noise-prediction loss at a random continuous noise level."""
        x_start = x_in["HR"]
        b = x_start.shape[0]
        t = np.random.randint(1, self.num_timesteps + 1)
        level = torch.FloatTensor(np.random.uniform(self.sqrt_alphas_cumprod_prev[t - 1],
                                                    self.sqrt_alphas_cumprod_prev[t], size=b)).to(x_start.device)
        level = level.view(b, -1)
        noise = torch.randn_like(x_start) if noise is None else noise
        x_noisy = self.q_sample(x_start, level.view(-1, 1, 1, 1), noise)
        if self.conditional:
            recon = self.denoise_fn(torch.cat([x_in["SR"], x_noisy], dim=1), level)
        else:
            recon = self.denoise_fn(x_noisy, level)
        return self.loss_func(noise, recon)

    def forward(self, x, sr_out=False, *args, **kwargs):
        if sr_out:
            # diffusion.py:243-273,308-310 back-propagates through a checkpointed sampler
            # (model3 training); training is outside the accelerated path.
            raise NotImplementedError("b200sr3: sr_out=True (differentiable sampling) is a training feature")
        if not isinstance(x, dict) and hasattr(x, "data") and isinstance(x.data, dict):
            x = x.data                  # the reference's DictTensor wrapper (diffusion.py:323-344)
        return self.p_losses(x, *args, **kwargs)
