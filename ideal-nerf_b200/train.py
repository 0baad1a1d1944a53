"""The optimisation step around render_rays, NeRFs/HeadNeRF/train/audio_exp_nerf.py:529-558: loss, backward, Adam, the
exponential learning-rate schedule and (multi-GPU) the gradient all-reduce, with the reference's hyper-parameter names.

    loss = mse(rgb, target) + mse(rgb0, target) + 10 * lc_weight * ||latent_code||_2            (:540-548)
    new_lrate = lrate * 0.1 ** (global_step / (lrate_decay * 1500))                             (:554-558)

The two image losses and their gradients come from one kernel (ops.mse_pair).  Adam is torch's fused implementation run on ONE tensor:
FlatParams re-homes every parameter as a view of a single buffer, so the optimiser step is one bandwidth-bound launch (~20 us for the
2.9 M parameters of two FaceNeRFs) instead of a multi-tensor sweep over 53 small tensors (~150 us), and the data-parallel all-reduce
runs on the flat gradient directly.  Adam is element-wise, so the result equals the per-parameter optimiser of the reference."""
import torch
import torch.distributed as dist

from . import ops


class FlatParams:
    """Parameters (and their gradients) as views of one flat buffer each.  Every view starts on a 256-byte boundary (the kernels read
    weights with vector loads).  `flat` is the single nn.Parameter the optimiser sees; `gather_grads()` copies the .grad tensors autograd
    produced into the flat gradient with one multi-tensor launch (parameters without a gradient contribute zeros = no update)."""
    ALIGN = 64      # floats

    def __init__(self, params):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatParams: no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        offs, o = [], 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("FlatParams: parameters must share device and dtype")
            offs.append(o)
            o += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.nn.Parameter(torch.zeros(o, device=dev, dtype=dt))
        self.grad = torch.zeros(o, device=dev, dtype=dt)
        self.flat.grad = self.grad
        self.gviews = []
        with torch.no_grad():
            for p, off in zip(self.params, offs):
                v = self.flat.data[off:off + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.gviews.append(self.grad[off:off + p.numel()].view(p.shape))

    def gather_grads(self, world=1, group=None):
        have = [(gv, p.grad) for gv, p in zip(self.gviews, self.params) if p.grad is not None]
        none = [gv for gv, p in zip(self.gviews, self.params) if p.grad is None]
        if none:
            torch._foreach_zero_(none)
        if have:
            torch._foreach_copy_([h[0] for h in have], [h[1] for h in have])
        if world > 1:      # the reference's nn.DataParallel backward (distribute_nerf.py:423) as one all-reduce of the flat gradient
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.grad, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
                self.grad.div_(world)
        self.flat.grad = self.grad


def head_loss(ret, target, latent_code, lc_weight):
    """Returns (loss, img_loss, latent_code_loss) as the reference logs them."""
    l2 = ops.mse_pair(ret["rgb_map"], ret["rgb0"], target)
    latent_loss = torch.norm(latent_code) * lc_weight
    return l2[0] + l2[1] + latent_loss * 10, l2[0], latent_loss


def learning_rate(args, global_step):
    return args.lrate * (0.1 ** (global_step / (args.lrate_decay * 1500)))


class TrainStep:
    """One call = one iteration of the reference's inner loop for a ray batch already on the device."""

    def __init__(self, network, latent_codes, args, world=1, group=None):
        self.net, self.latent_codes, self.args, self.world, self.group = network, latent_codes, args, world, group
        latent_codes.requires_grad_(True)
        self.flat = FlatParams(list(network.parameters()) + [latent_codes])
        self.optimizer = torch.optim.Adam([self.flat.flat], lr=args.lrate, betas=(0.9, 0.999), fused=latent_codes.is_cuda)
        self.global_step = 0

    def __call__(self, rays, bc_rgb, target, aud_feature, expr, index, perturb=None):
        a = self.args
        latent_code = self.latent_codes[index]
        for p in self.flat.params:
            p.grad = None
        ret = self.net.render_rays(rays, bc_rgb, aud_feature, None, latent_code, expr, perturb=a.perturb if perturb is None else perturb)
        loss, img_loss, latent_loss = head_loss(ret, target, latent_code, a.lc_weight)
        loss.backward()
        self.flat.gather_grads(self.world, self.group)
        self.optimizer.step()
        lr = learning_rate(a, self.global_step)
        for g in self.optimizer.param_groups:
            g["lr"] = lr
        self.global_step += 1
        return {"loss": loss.detach(), "img_loss": img_loss.detach(), "latent_code_loss": latent_loss.detach(), "lr": lr}
