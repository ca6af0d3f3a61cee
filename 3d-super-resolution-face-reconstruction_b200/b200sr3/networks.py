"""Drop-in for model/sr/networks.py::define_G (reference lines 83-116): same `opt` mapping in,
an nn.Module with the reference GaussianDiffusion surface out."""
from . import _lib
from .diffusion import GaussianDiffusion
from .unet import UNet


def define_G(opt):
    model_opt = opt["sr"]["model"]
    which = model_opt["which_model_G"]
    if which != "sr3":
        # every shipped YAML selects sr3 (config/*.yml:33); the ddpm branch is not accelerated
        raise NotImplementedError(f"b200sr3 implements which_model_G == 'sr3' only, got {which!r}")
    unet_opt = model_opt["unet"]
    if ("norm_groups" not in unet_opt) or unet_opt["norm_groups"] is None:
        unet_opt["norm_groups"] = 32          # the reference writes this default back into opt
    diff_opt = model_opt["diffusion"]
    model = UNet(
        in_channel=unet_opt["in_channel"],
        out_channel=unet_opt["out_channel"],
        norm_groups=unet_opt["norm_groups"],
        inner_channel=unet_opt["inner_channel"],
        channel_mults=unet_opt["channel_multiplier"],
        attn_res=unet_opt["attn_res"],
        res_blocks=unet_opt["res_blocks"],
        dropout=unet_opt["dropout"],
        image_size=diff_opt["image_size"],
    )
    netG = GaussianDiffusion(
        model,
        image_size=diff_opt["image_size"],
        channels=diff_opt["channels"],
        loss_type="l1",
        conditional=diff_opt["conditional"],
        schedule_opt=model_opt["beta_schedule"]["train"],
    )
    netG._cfg = _lib.make_config(unet_opt, diff_opt)
    if opt["phase"] == "train":
        model.init_orthogonal()
    return netG
