"""The optimisation step around render_rays, NeRFs/HeadNeRF/train/audio_exp_nerf.py:529-558: loss, backward, Adam, the
exponential learning-rate schedule and (multi-GPU) the gradient all-reduce, with the reference's hyper-parameter names.

    loss = mse(rgb, target) + mse(rgb0, target) + 10 * lc_weight * ||latent_code||_2            (:540-548)
    new_lrate = lrate * 0.1 ** (global_step / (lrate_decay * 1500))                             (:554-558)

The two image losses and their gradients come from one kernel (ops.mse_pair); Adam is torch's fused multi-tensor implementation."""
import torch

from . import ops
from .frame import allreduce_grads


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
        params = list(network.parameters()) + [latent_codes]
        latent_codes.requires_grad_(True)
        self.optimizer = torch.optim.Adam(params, lr=args.lrate, betas=(0.9, 0.999), fused=latent_codes.is_cuda)
        self.global_step = 0

    def __call__(self, rays, bc_rgb, target, aud_feature, expr, index, perturb=None):
        a = self.args
        latent_code = self.latent_codes[index]
        self.optimizer.zero_grad(set_to_none=True)
        ret = self.net.render_rays(rays, bc_rgb, aud_feature, None, latent_code, expr, perturb=a.perturb if perturb is None else perturb)
        loss, img_loss, latent_loss = head_loss(ret, target, latent_code, a.lc_weight)
        loss.backward()
        if self.world > 1:
            allreduce_grads(list(self.net.parameters()) + [self.latent_codes], self.world, self.group)
        self.optimizer.step()
        lr = learning_rate(a, self.global_step)
        for g in self.optimizer.param_groups:
            g["lr"] = lr
        self.global_step += 1
        return {"loss": loss.detach(), "img_loss": img_loss.detach(), "latent_code_loss": latent_loss.detach(), "lr": lr}
