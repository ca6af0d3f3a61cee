"""The `sr.model` block of the reference YAMLs (config/sr_sr3_VGGF2_<L>_<R>_model{2,3}.yml).

All 22 files share one UNet and differ only in LR/HR size and n_timestep,
so the named configs are generated here instead of vendoring the YAML files; `load_yaml` reads
an unmodified reference YAML when one is available.
"""
import copy

# (l_resolution, r_resolution) -> n_timestep, from config/*.yml:52,57
TIMESTEPS = {(8, 16): 100, (8, 32): 400, (8, 64): 600, (8, 128): 1000, (16, 32): 200, (16, 64): 200,
             (16, 128): 600, (32, 64): 100, (32, 128): 100, (64, 128): 100}


def model_opt(n_timestep):
    sched = {"schedule": "linear", "n_timestep": int(n_timestep), "linear_start": 1e-6, "linear_end": 1e-2}
    return {
        "which_model_G": "sr3",
        "finetune_norm": False,
        "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64, "channel_multiplier": [1, 2, 4, 8, 8],
                 "attn_res": [16], "res_blocks": 2, "dropout": 0.2},
        "beta_schedule": {"train": copy.deepcopy(sched), "val": copy.deepcopy(sched)},
        "diffusion": {"image_size": 224, "channels": 3, "conditional": True},
    }


def named(name, phase="val"):
    """'sr_sr3_VGGF2_16_128_model3' -> the opt mapping define_G expects, plus l/r resolution."""
    parts = name.replace(".yml", "").split("_")
    lr, hr = int(parts[3]), int(parts[4])
    return {"name": name, "phase": phase, "l_resolution": lr, "r_resolution": hr,
            "sr": {"model": model_opt(TIMESTEPS[(lr, hr)])}}


def load_yaml(path, phase="val"):
    import yaml
    with open(path) as f:
        opt = yaml.safe_load(f)
    opt["phase"] = phase
    return opt
