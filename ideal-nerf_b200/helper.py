"""Host-side mirror of NeRFs/HeadNeRF/helper.py (and its twin NeRFs/TorsoNeRF/run_nerf_helpers.py).

Same names, argument meaning and return shapes as the reference; the arithmetic runs in the CUDA
kernels behind include/inerf_b200.h.
"""
import argparse

import numpy as np
import torch

from . import _lib, ops


def config_parser():
    """Flags of the hot path, same names and defaults as helper.py:16-138.

    configargparse is not a dependency here: ``--config file`` is read as ``key = value`` lines, and
    argparse prefix matching keeps the reference's ``N_sample`` spelling working (README.md:38).
    """
    parser = _ConfigParser()
    a = parser.add_argument
    a('--config', type=str, default=None, help='config file path (key = value lines)')
    a("--expname", type=str); a("--basedir", type=str)
    a("--datadir", type=str, default='./dataset/Obama'); a("--vis_path", type=str, default='./dataset/Obama/run')
    a("--save_path", type=str, default='output/render/Obama-Noah/'); a("--evalExpr_path", type=str)
    a("--mouth_rays", type=int, default=0); a("--torso_rays", type=int, default=0)
    a("--dim_expr", type=int, default=0); a("--dim_aud", type=int, default=0)
    a("--dim_aud_body", type=int, default=64)
    a("--lc_weight", type=float, default=0.0005); a("--gt_dirs", type=str, default='head_imgs')
    a("--gpu_num", type=int, default=0); a("--num_work", type=int, default=3); a("--batch_size", type=int, default=4)
    a("--netdepth", type=int, default=8); a("--netwidth", type=int, default=256)
    a("--netdepth_fine", type=int, default=8); a("--netwidth_fine", type=int, default=256)
    a("--N_rand", type=int, default=2048); a("--lrate", type=float, default=8e-4)
    a("--lrate_decay", type=int, default=500)
    a("--chunk", type=int, default=1024 * 8); a("--netchunk", type=int, default=1024 * 64)
    a("--use_batching", action='store_false'); a("--no_reload", action='store_true')
    a("--ft_path", type=str, default=None); a("--N_iters", type=int, default=90)
    a("--N_samples", type=int, default=64); a("--N_importance", type=int, default=128)
    a("--perturb", type=float, default=1.)
    a("--use_viewdirs", action='store_false')          # store_false => default True, as in the reference
    a("--i_embed", type=int, default=0); a("--multires", type=int, default=10)
    a("--multires_views", type=int, default=4); a("--raw_noise_std", type=float, default=0.)
    a("--render_only", action='store_true'); a("--render_test", action='store_true')
    a("--render_factor", type=int, default=0)
    a("--precrop_iters", type=int, default=0); a("--precrop_frac", type=float, default=.5)
    a("--testskip", type=int, default=8)
    a("--white_bkgd", action='store_false'); a("--half_res", action='store_true')
    a("--with_test", type=int, default=0); a("--sample_rate", type=float, default=0.95)
    a("--near", type=float, default=0.3); a("--far", type=float, default=0.9)
    a("--test_file", type=str); a("--aud_file", type=str, default='aud.npy')
    a("--win_size", type=int, default=16); a("--smo_size", type=int, default=8)
    a('--nosmo_iters', type=int, default=300000)
    a("--no_ndc", action='store_true'); a("--lindisp", action='store_true')
    a("--i_print", type=int, default=10); a("--i_img", type=int, default=500)
    a("--i_weights", type=int, default=5000); a("--i_testset", type=int, default=1000)
    a("--i_video", type=int, default=5000)
    # B200 build only: arithmetic mode of the FaceNeRF kernel ("fp32" | "bf16")
    a("--mlp_mode", type=str, default="fp32")                       # fp32 (FFMA) | fp16x2 (tensor cores, fp32 gate) | bf16 (tensor cores, PSNR gate)
    a("--check_numerics", action="store_true")                      # the NaN / Inf report of audio_exp_nerf.py:367-369 (one kernel + one host read)
    return parser


class _ConfigParser(argparse.ArgumentParser):
    def parse_args(self, args=None, namespace=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        pre, _ = argparse.ArgumentParser.parse_known_args(self, argv)
        if getattr(pre, "config", None):
            extra = []
            with open(pre.config) as fh:
                for line in fh:
                    line = line.split('#')[0].strip()
                    if not line:
                        continue
                    k, _, v = line.partition('=')
                    k, v = k.strip(), v.strip()
                    if v.lower() in ("true", ""):
                        extra.append('--' + k)
                    elif v.lower() != "false":
                        extra += ['--' + k, v]
            argv = extra + argv                      # command line wins over the file
        return argparse.ArgumentParser.parse_args(self, argv, namespace)


# ------------------------------------------------------------------------------------------------
# positional encoding (helper.py:174-224)
# ------------------------------------------------------------------------------------------------
class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if not (kwargs.get('include_input', True) and kwargs.get('log_sampling', True)):
            raise NotImplementedError("only include_input=True, log_sampling=True (what get_embedder builds)")
        self.n_freqs = kwargs['num_freqs']
        self.out_dim = kwargs['input_dims'] * (1 + 2 * self.n_freqs)

    def embed(self, inputs):
        return ops.posenc(inputs, self.n_freqs)


def get_embedder(multires, i=0, input_dims=3):
    if i == -1:
        return torch.nn.Identity(), 3
    embedder_obj = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1,
                            num_freqs=multires, log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=embedder_obj):
        return eo.embed(x)

    return embed, embedder_obj.out_dim


# ------------------------------------------------------------------------------------------------
# rays (helper.py:228-243)
# ------------------------------------------------------------------------------------------------
def get_rays(H, W, focal, c2w, cx=None, cy=None):
    """rays_o, rays_d of shape (H, W, 3).  c2w must be a CUDA tensor."""
    rays = ops.get_rays_packed(H, W, focal, c2w[:3, :4], 0., 1., cx, cy)
    return rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)


# ------------------------------------------------------------------------------------------------
# hierarchical sampling (helper.py:269-313)
# ------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, pytest=False, policy=_lib.INERF_PDF_EXACT_TORCH_CPU):
    dev = bins.device
    n = bins.shape[0]
    if pytest:                                     # helper.py:285-293
        np.random.seed(0)
        if det:
            u = torch.Tensor(np.linspace(0., 1., N_samples)).to(dev)
        else:
            u = torch.Tensor(np.random.rand(n, N_samples)).to(dev)
    elif det:
        u = ops.linspace_table(N_samples, dev)
    else:
        u = torch.rand((n, N_samples), device=dev)
    samples, _ = ops.sample_pdf_raw(bins, weights, u, policy)
    return samples


def to8b(x):
    """helper.py:154 ``(255 * np.clip(x, 0, 1)).astype(np.uint8)`` for a CUDA tensor: returns a uint8 CUDA tensor of the same shape
    (one kernel; the frame then leaves the GPU as bytes)."""
    return ops.to8b(x)
