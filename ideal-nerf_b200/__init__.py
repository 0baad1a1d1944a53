"""ideal-nerf_b200: B200-native (sm_100a) implementation of IDEAL-NeRF's render_rays hot path.

Host side mirrors the reference's Python surface (same names/signatures); the work is done by the
hand-written CUDA kernels in csrc/ behind the C ABI of include/inerf_b200.h.  There is no CPU,
PyTorch-eager or Triton fallback: importing is cheap, but every op raises if libinerf_b200.so has
not been built or the tensors are not on a CUDA (sm_100) device.
"""
from . import _lib, ops                                                   # noqa: F401
from ._lib import build, lib                                              # noqa: F401
from .face_nerf import FaceNeRF                                           # noqa: F401
from .audio_net import AudioNet, AudioAttNet                              # noqa: F401
from .helper import config_parser, get_embedder, get_rays, sample_pdf, Embedder, to8b   # noqa: F401
from .render import (Network, TorsoNetwork, raw2outputs, raw2outputs_torso, render_rays, init_weights,   # noqa: F401
                     pose_to_euler_trans, default_args)

__all__ = ["FaceNeRF", "AudioNet", "AudioAttNet", "Network", "TorsoNetwork", "raw2outputs", "raw2outputs_torso", "render_rays", "sample_pdf",
           "get_embedder", "get_rays", "to8b", "config_parser", "init_weights", "pose_to_euler_trans", "build", "lib", "ops"]
