"""Full-frame driver around render_rays, with the frame's rays block-partitioned across GPUs.

Reference: the per-frame loop of NeRFs/HeadNeRF/test/eval_aud_exp_nerf.py:485-495 (pose -> get_rays ->
batchify_rays -> rgb) and the only multi-GPU code the reference has, nn.DataParallel over a ray
re-shape (NeRFs/HeadNeRF/train/distribute_nerf.py:423,457-462).  Here: one process per GPU
(torch.distributed), contiguous row bands of the image per rank, a full weight replica per rank, and
ONE collective per frame -- the all-gather of the rendered bands (rays are independent: SURVEY.md 8e).
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
    """Renders frames whose rays are block-partitioned over `world` ranks.  Each rank generates ONLY its band's rays
    (inerf_get_rays_range), renders them, and the bands meet in ONE collective per frame: an all_gather_into_tensor of the fixed-size
    (per, 3) band buffers (2.4 MB per frame over NVLink; NCCL runs it on its own stream, so with async_gather the next frame's kernels
    overlap it).  Band and image buffers are allocated once and double-buffered: no per-frame torch.cat / pad / allocation."""

    def __init__(self, network, rank=0, world=1, group=None):
        self.net, self.rank, self.world, self.group = network, rank, world, group
        self._bufs = None            # [(band (per,3), image (world*per,3))] x 2
        self._pending = [None, None]
        self._slot = 0

    def _buffers(self, n_rays, device):
        per = (n_rays + self.world - 1) // self.world
        if self._bufs is None or self._bufs[0][0].shape[0] != per or self._bufs[0][0].device != device:
            self._bufs = [(torch.zeros((per, 3), device=device), torch.empty((self.world * per, 3), device=device)) for _ in range(2)]
            self._pending = [None, None]
        return self._bufs

    def render_band(self, pose, aud, expr, latent, bc_rgb, perturb=0., lo=None, hi=None):
        """Render this rank's rays of one H x W frame.  bc_rgb: (H*W,3) or (H,W,3) full background."""
        n = self.net
        H, W = n.H, n.W
        if lo is None:
            lo, hi = band(H * W, self.rank, self.world)
        rays = ops.get_rays_range(H, W, n.focal, pose[:3, :4], n.near, n.far, lo, hi - lo)
        bc = bc_rgb.reshape(-1, 3)[lo:hi]
        ret = _render_rays_impl(rays, bc, n.face_nerf_coarse, n.face_nerf_fine, aud, expr, latent,
                                n.args.N_samples, n.args.N_importance, perturb=perturb)
        return ret, (lo, hi)

    def gather_image(self, rgb_band, n_rays, dst=0, async_op=False):
        """Assemble the (n_rays,3) image from every rank's band.  Returns the image (a view of a persistent buffer that the gather after
        next overwrites), or -- with async_op -- a handle whose .wait() returns it.  Every rank receives the image."""
        if self.world == 1:
            return _Done(rgb_band) if async_op else rgb_band
        bufs = self._buffers(n_rays, rgb_band.device)
        slot = self._slot
        self._slot ^= 1
        if self._pending[slot] is not None:                            # the gather that last used these buffers
            self._pending[slot].wait()
            self._pending[slot] = None
        band_buf, img = bufs[slot]
        band_buf[:rgb_band.shape[0]].copy_(rgb_band)                   # tail rank: the rows past its band stay zero
        work = dist.all_gather_into_tensor(img, band_buf, group=self.group, async_op=True)
        h = _Gathered(work, img[:n_rays])
        if async_op:
            self._pending[slot] = h
            return h
        return h.wait()

    def render_frame(self, pose, aud, expr, latent, bc_rgb, perturb=0., async_op=False):
        ret, _ = self.render_band(pose, aud, expr, latent, bc_rgb, perturb)
        return self.gather_image(ret['rgb_map'], self.net.H * self.net.W, async_op=async_op)


class _Done:
    def __init__(self, img):
        self.img = img

    def wait(self):
        return self.img


class _Gathered:
    """Handle of an in-flight band gather: wait() makes the current stream wait for the collective and returns the image."""

    def __init__(self, work, img):
        self.work, self.img = work, img

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self.img


def render_video(renderer, frames, latent, bc_rgb, perturb=0.):
    """The eval loop of eval_aud_exp_nerf.py:485-495 without its per-frame host round trip: every frame is rendered (rays sharded over
    the ranks of `renderer`), gathered, converted to uint8 on the device (to8b) and copied asynchronously into ONE pinned
    (T, H, W, 3) uint8 host array -- 0.6 MB per frame instead of 2.4 MB of fp32, no synchronisation until the last frame.  The gather of
    frame i overlaps the kernels of frame i+1 (it is consumed one frame late).
    frames: sequence of (pose (4,4) or (3,4), aud (dim_aud,), expr (dim_expr,)) device tensors.  Returns the host array on rank 0."""
    n = renderer.net
    H, W = n.H, n.W
    out = torch.empty((len(frames), H, W, 3), dtype=torch.uint8).pin_memory() if renderer.rank == 0 else None

    def finish(i, h):
        rgb = h.wait()
        if renderer.rank == 0:
            out[i].copy_(ops.to8b(rgb).reshape(H, W, 3), non_blocking=True)

    prev = None
    for i, (pose, aud, expr) in enumerate(frames):
        h = renderer.render_frame(pose, aud, expr, latent, bc_rgb, perturb, async_op=True)
        if prev is not None:
            finish(*prev)
        prev = (i, h)
    if prev is not None:
        finish(*prev)
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
