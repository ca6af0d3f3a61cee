"""Checkpoint ingestion for the drop-in generator.

The reference stores the SR generator in two ways:
  * `<prefix>_gen.pth` - the bare `netG.state_dict()` (model/sr/model.py:146-162 save_network, :164-195
    load_network; lib/trainer_temp.py:190-216): 337 `denoise_fn.*` tensors, and the 12 schedule buffers when the
    schedule had been installed before saving;
  * a combined training checkpoint whose `['sr_model_state']` is the same dict, possibly with a `module.` prefix from
    DataParallel / DistributedDataParallel (lib/trainer_temp.py:165-188).
Both load into b200sr3's GaussianDiffusion unchanged, because its parameter names and shapes are the reference's
(tests/test_host_cpu.py::test_state_dict_contract); the engine repacks them to bf16 K-major on the next sampling call.
"""
import os

import torch

_SCHEDULE_KEYS = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2")


def strip_module_prefix(state):
    """`module.` is what nn.DataParallel / DDP prepend (lib/trainer_temp.py:176-178 adds it, we remove it)."""
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}


def _load(netG, state, strict):
    state = strip_module_prefix(state)
    # The schedule buffers are re-derived by set_new_noise_schedule (diffusion.py:93-142) for the phase in use; a
    # checkpoint saved at another T must not overwrite them (the reference loads with strict=False for that reason).
    own = netG.state_dict()
    state = {k: v for k, v in state.items()
             if not (k in _SCHEDULE_KEYS and (k not in own or tuple(own[k].shape) != tuple(v.shape)))}
    missing, unexpected = netG.load_state_dict(state, strict=False)
    missing = [k for k in missing if k not in _SCHEDULE_KEYS]
    if strict and (missing or unexpected):
        raise RuntimeError(f"b200sr3: checkpoint does not match the generator: missing {missing[:5]}..., "
                           f"unexpected {list(unexpected)[:5]}...")
    return missing, list(unexpected)


def load_network(netG, load_path, strict=True):
    """model/sr/model.py:164-195: `load_path` is the prefix, the generator lives in `<load_path>_gen.pth`."""
    gen_path = "{}_gen.pth".format(load_path)
    if not os.path.exists(gen_path):
        raise FileNotFoundError(gen_path)
    return _load(netG, torch.load(gen_path, map_location="cpu", weights_only=True), strict)


def load_combined(netG, path, strict=False, allow_pickle=False):
    """lib/trainer_temp.py:165-188: the SR generator out of a combined checkpoint (`sr_model_state`).

    Loaded with torch's restricted unpickler (tensors and plain containers only). A combined checkpoint that also
    pickles arbitrary Python objects (the reference stores optimizer state and the config node beside the weights) is
    refused unless the caller opts in with allow_pickle=True - unpickling executes code from the file."""
    try:
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:          # pickle.UnpicklingError from the allow-list
        if not allow_pickle:
            raise RuntimeError(f"b200sr3: '{path}' holds objects torch's safe loader refuses ({e}); pass "
                               "allow_pickle=True only for a checkpoint you trust") from e
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if "sr_model_state" not in ckpt:
        raise KeyError("b200sr3: not a combined checkpoint (no 'sr_model_state')")
    return _load(netG, ckpt["sr_model_state"], strict)


def save_network(netG, save_path):
    """model/sr/model.py:146-162: writes `<save_path>_gen.pth` with CPU tensors, as the reference does."""
    state = {k: v.detach().cpu() for k, v in netG.state_dict().items()}
    torch.save(state, "{}_gen.pth".format(save_path))
    return "{}_gen.pth".format(save_path)
