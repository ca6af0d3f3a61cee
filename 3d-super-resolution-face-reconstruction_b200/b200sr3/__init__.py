"""b200sr3 — B200-native drop-in for the SR3 sampling path of
zouiner/3d-super-resolution-Face-reconstruction (define_G / GaussianDiffusion)."""
from .networks import define_G                      # noqa: F401
from .diffusion import GaussianDiffusion, make_beta_schedule   # noqa: F401
from .unet import UNet                              # noqa: F401
from . import configs                               # noqa: F401
from . import mica_handoff                          # noqa: F401
from .arcface import Arcface, MappingNetwork, MicaEncoder      # noqa: F401

__all__ = ["define_G", "GaussianDiffusion", "UNet", "make_beta_schedule", "configs", "mica_handoff",
           "Arcface", "MappingNetwork", "MicaEncoder"]
