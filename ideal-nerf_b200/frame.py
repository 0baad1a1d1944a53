"""Full-frame driver around render_rays, with the frame's rays block-partitioned across GPUs.

Reference: the per-frame loop of NeRFs/HeadNeRF/test/eval_aud_exp_nerf.py:485-495 (pose -> get_rays ->
batchify_rays -> rgb) and the only multi-GPU code the reference has, nn.DataParallel over a ray
re-shape (NeRFs/HeadNeRF/train/distribute_nerf.py:423,457-462).  Here: one process per GPU
(torch.distributed), contiguous row bands of the image per rank, a full weight replica per rank, and
ONE collective per frame -- the gather of the rendered bands (rays are independent: SURVEY.md 8e).
"""
import torch
import torch.distributed as dist

from . import ops
from .render import _render_rays_impl


def band(n_rays, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; every rank gets ceil(n/world) rays except the tail."""
    per = (n_rays + world - 1) // world
    lo = min(n_rays, rank * per)
    return lo, min(n_rays, lo + per)


class FrameRenderer:
    def __init__(self, network, rank=0, world=1, group=None):
        self.net, self.rank, self.world, self.group = network, rank, world, group

    def render_band(self, pose, aud, expr, latent, bc_rgb, perturb=0., lo=None, hi=None):
        """Render this rank's rays of one H x W frame.  bc_rgb: (H*W,3) or (H,W,3) full background."""
        n = self.net
        H, W = n.H, n.W
        rays = ops.get_rays_packed(H, W, n.focal, pose[:3, :4], n.near, n.far)
        if lo is None:
            lo, hi = band(H * W, self.rank, self.world)
        bc = bc_rgb.reshape(-1, 3)[lo:hi]
        ret = _render_rays_impl(rays[lo:hi], bc, n.face_nerf_coarse, n.face_nerf_fine, aud, expr, latent,
                                n.args.N_samples, n.args.N_importance, perturb=perturb)
        return ret, (lo, hi)

    def gather_image(self, rgb_band, n_rays, dst=0):
        """Assemble the (n_rays,3) image on rank `dst` from every rank's band (NCCL gather over NVLink)."""
        if self.world == 1:
            return rgb_band
        per = (n_rays + self.world - 1) // self.world
        if rgb_band.shape[0] < per:                                  # tail rank: pad to the common band size
            pad = rgb_band.new_zeros((per - rgb_band.shape[0], 3))
            rgb_band = torch.cat([rgb_band, pad], 0)
        rgb_band = rgb_band.contiguous()
        if self.rank == dst:
            out = torch.empty((self.world * per, 3), device=rgb_band.device, dtype=rgb_band.dtype)
            dist.gather(rgb_band, list(out.split(per, 0)), dst=dst, group=self.group)
            return out[:n_rays]
        dist.gather(rgb_band, None, dst=dst, group=self.group)
        return None

    def render_frame(self, pose, aud, expr, latent, bc_rgb, perturb=0.):
        ret, _ = self.render_band(pose, aud, expr, latent, bc_rgb, perturb)
        return self.gather_image(ret['rgb_map'], self.net.H * self.net.W)


def render_video(renderer, frames, latent, bc_rgb, perturb=0.):
    """The eval loop of eval_aud_exp_nerf.py:485-495 without its per-frame host round trip: every frame is rendered (rays sharded over
    the ranks of `renderer`), gathered on rank 0, converted to uint8 on the device (to8b) and copied asynchronously into ONE pinned
    (T, H, W, 3) uint8 host array -- 0.6 MB per frame instead of 2.4 MB of fp32, no synchronisation until the last frame.
    frames: sequence of (pose (4,4) or (3,4), aud (dim_aud,), expr (dim_expr,)) device tensors.  Returns the host array on rank 0."""
    n = renderer.net
    H, W = n.H, n.W
    out = torch.empty((len(frames), H, W, 3), dtype=torch.uint8).pin_memory() if renderer.rank == 0 else None
    for i, (pose, aud, expr) in enumerate(frames):
        rgb = renderer.render_frame(pose, aud, expr, latent, bc_rgb, perturb)
        if renderer.rank == 0:
            out[i].copy_(ops.to8b(rgb).reshape(H, W, 3), non_blocking=True)
    torch.cuda.synchronize()
    return out


def allreduce_grads(parameters, world, group=None):
    """Data-parallel training step, the reference's nn.DataParallel backward (distribute_nerf.py:423) as one process per GPU: every rank
    ran render_rays + loss.backward() on its band of the N_rand rays (band(N_rand, rank, world)); the gradients of all parameters are
    flattened into ONE buffer, summed with a single NCCL all-reduce (~6 MB over NVLink) and divided by `world`, so that mean-over-rays
    losses reproduce the single-process gradient when the bands have equal size.  Parameters without a gradient are skipped."""
    ps = [p for p in parameters if p.grad is not None]
    if world == 1 or not ps:
        return
    grads = [p.grad for p in ps]
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)          # the division happens inside the collective
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    # one multi-tensor copy back instead of a launch per parameter (the step is ~5 ms: 50 tiny launches would show)
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
