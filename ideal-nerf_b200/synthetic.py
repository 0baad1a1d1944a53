"""Seeded synthetic inputs of the shapes BASELINE.json names: camera, background,
audio / expression / latent codes, random-init FaceNeRF weights with the normalised-density preset.

Used by bench.py and __graft_entry__.smoke(); there is no dataset or checkpoint in this environment.
The tensors follow the same generator sequence as the test oracle's synthetic_frame so both arms of
the benchmark see the same data, but nothing here imports the test infrastructure.
"""
import torch

from . import ops

NEAR = 0.5772005200386048   # NeRFs/HeadNeRF/configs/audio_expr_nerf/may/paper_model/torso_bg.txt:11-12
FAR = 1.1772005200386046


def camera():
    c2w = torch.eye(4)
    c2w[2, 3] = 0.7772
    return dict(H=450, W=450, focal=1200., cx=225., cy=225., c2w=c2w)


def frame_inputs(seed=0):
    """Host tensors of one frame: pose (4,4), bc_rgb (H*W,3), aud (64), expr (76), latent (32)."""
    cam = camera()
    g = torch.Generator().manual_seed(seed)
    bc = torch.rand(cam["H"] * cam["W"], 3, generator=g)
    aud = torch.randn(64, generator=g)
    expr = torch.randn(76, generator=g)
    return dict(pose=cam["c2w"], bc_rgb=bc, aud=aud, expr=expr, latent=torch.ones(32))


def video_codes(n_frames, seed=0, sigma=0.1):
    """Seeded Gaussian random walk of audio/expression codes for an n_frames eval video (config 5)."""
    g = torch.Generator().manual_seed(seed + 17)
    aud = torch.cumsum(torch.randn(n_frames, 64, generator=g) * sigma, 0)
    expr = torch.cumsum(torch.randn(n_frames, 76, generator=g) * sigma, 0)
    return aud, expr


@torch.no_grad()
def normalise_density_(net, rays, aud, expr, latent, n_samples=64, target_std=8.0):
    """SURVEY.md 7-7 preset, in place: rescale alpha_linear so raw sigma on the coarse samples is ~N(0.01, 8^2).
    Random-init FaceNeRF gives sigma ~ -0.3 +- 0.1 => every ray would be pure background."""
    mode, net.mlp_mode = net.mlp_mode, "fp32"
    z = ops.sample_coarse(rays, n_samples)
    s = net.query(rays, z, aud, expr, latent)[..., 3]
    net.mlp_mode = mode
    k = target_std / float(s.std())
    net.alpha_linear.weight.mul_(k)
    net.alpha_linear.bias.copy_((net.alpha_linear.bias - float(s.mean())) * k + 0.01)
    return net
