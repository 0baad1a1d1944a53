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
        self.offsets = offs
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
    """One call = one iteration of the reference's inner loop for a ray batch already on the device.

    The optimiser is ONE Adam over the flat buffer; `optimizer_state_dict()` / `load_optimizer_state_dict()` convert to and from the
    reference's per-parameter Adam state (params = list(network.parameters()) + [latent_codes], audio_exp_nerf.py:487-493) so the
    'optimizer' entry of head.tar moves both ways; `save()` / `load()` write and resume the reference's checkpoint (:516-525, :584-591)."""

    def __init__(self, network, latent_codes, args, world=1, group=None, cuda_graph=False):
        self.net, self.latent_codes, self.args, self.world, self.group = network, latent_codes, args, world, group
        latent_codes.requires_grad_(True)
        self.flat = FlatParams(list(network.parameters()) + [latent_codes])
        self.cuda_graph = bool(cuda_graph)
        self._static = None
        if self.cuda_graph:
            if not latent_codes.is_cuda:
                raise RuntimeError("TrainStep(cuda_graph=True) needs CUDA tensors")
            if world > 1:
                raise NotImplementedError("cuda_graph=True is built for one process per model replica without a gradient all-reduce")
            dev = latent_codes.device
            self._lr_t = torch.tensor(float(args.lrate), device=dev)          # the learning rate and the step count live on the device
            self._gs_t = torch.zeros((), device=dev)
            self.optimizer = torch.optim.Adam([self.flat.flat], lr=self._lr_t, betas=(0.9, 0.999), fused=True, capturable=True)
        else:
            self.optimizer = torch.optim.Adam([self.flat.flat], lr=args.lrate, betas=(0.9, 0.999), fused=latent_codes.is_cuda)
        self.global_step = 0

    # -- checkpoint interchange with the reference's per-parameter Adam -----------------------------------------------------------
    def optimizer_state_dict(self):
        """Adam state in the layout torch.optim.Adam(list(network.parameters()) + [latent_codes]) writes: one entry per parameter."""
        sd = self.optimizer.state_dict()
        st = sd["state"].get(0)
        state = {}
        if st:
            for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
                n = p.numel()
                state[i] = {"step": st["step"].detach().clone().cpu(), "exp_avg": st["exp_avg"][off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": st["exp_avg_sq"][off:off + n].view(p.shape).clone()}
        group = {k: v for k, v in sd["param_groups"][0].items() if k != "params"}
        group.update(fused=None, foreach=None, params=list(range(len(self.flat.params))))
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Accepts the per-parameter layout above (the reference's head.tar) or this class's own flat one-parameter layout."""
        n_par = len(self.flat.params)
        ids = sd["param_groups"][0]["params"]
        if len(ids) == 1 and n_par != 1:
            self.optimizer.load_state_dict(sd)
            return
        if len(ids) != n_par:
            raise ValueError(f"optimizer state has {len(ids)} parameters, this TrainStep has {n_par}")
        flat = self.flat.flat
        exp_avg, exp_avg_sq = torch.zeros_like(flat.data), torch.zeros_like(flat.data)
        step = None
        for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
            st = sd["state"].get(ids[i])
            if not st:                                   # a parameter that never received a gradient has no state in the reference
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state of parameter {i} has shape {tuple(st['exp_avg'].shape)}, expected {tuple(p.shape)}")
            n = p.numel()
            exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            s = float(st["step"])
            step = s if step is None else max(step, s)
        group = {k: v for k, v in self.optimizer.state_dict()["param_groups"][0].items()}
        for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
            if k in sd["param_groups"][0]:
                group[k] = sd["param_groups"][0][k]
        state = {} if step is None else {0: {"step": torch.tensor(step, dtype=torch.float32), "exp_avg": exp_avg, "exp_avg_sq": exp_avg_sq}}
        self.optimizer.load_state_dict({"state": state, "param_groups": [group]})

    def save(self, path):
        """head.tar as the reference writes it (audio_exp_nerf.py:584-591), loadable by the reference's resume code (:516-525)."""
        torch.save({"global_step": int(self.global_step), "model_state_dict": self.net.state_dict(),
                    "optimizer": self.optimizer_state_dict(), "latent_codes": self.latent_codes.data.clone()}, path)

    def load(self, path, map_location=None):
        """Resume from a head.tar written by the reference or by save(): weights and latent codes are copied IN PLACE (they are views of
        the flat Adam buffer), the Adam moments are scattered into the flat state, global_step and the learning rate are restored."""
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        self.net.load_state_dict(ckpt["model_state_dict"])                 # nn.Module.load_state_dict copies in place
        with torch.no_grad():
            self.latent_codes.data.copy_(ckpt["latent_codes"].to(self.latent_codes.device))
        self.global_step = int(ckpt.get("global_step", 0))
        if ckpt.get("optimizer") is not None:
            self.load_optimizer_state_dict(ckpt["optimizer"])              # carries the learning rate the schedule had reached (:554-558)
        else:
            for g in self.optimizer.param_groups:                          # the lr in force at step k is the one set after step k-1
                g["lr"] = learning_rate(self.args, self.global_step - 1) if self.global_step > 0 else self.args.lrate
        if self.cuda_graph:                                                # keep the device-resident schedule state the graph reads
            if self._static is not None and self._static["graph"] is not None:
                raise RuntimeError("TrainStep.load() after the CUDA graph was captured: load before the first graphed step")
            for g in self.optimizer.param_groups:
                self._lr_t.fill_(float(g["lr"]))
                g["lr"] = self._lr_t
            self._gs_t.fill_(float(self.global_step))
        for m in self.net.modules():
            if hasattr(m, "invalidate_packed"):
                m.invalidate_packed()
        return self.global_step

    def __call__(self, rays, bc_rgb, target, aud_feature, expr, index, perturb=None, aud_window=None):
        """aud_feature: the (dim_aud,) audio code (a constant, or a tensor with a graph into the conditioning nets); or pass
        aud_window = the (smo_size, 16, 29) DeepSpeech window (Network.audio_window) and the code is computed -- and the two
        conditioning nets are trained -- inside the step, as the reference's forward does (audio_exp_nerf.py:241-266)."""
        perturb = self.args.perturb if perturb is None else perturb
        if self.cuda_graph:
            return self._graphed(rays, bc_rgb, target, aud_feature, expr, index, perturb, aud_window)
        a = self.args
        latent_code = self.latent_codes[index]
        if aud_window is not None:
            aud_feature = self.net.aud_att_net(self.net.aud_net(aud_window))
        loss, img_loss, latent_loss = self._body(rays, bc_rgb, target, aud_feature, expr, latent_code, perturb)
        lr = learning_rate(a, self.global_step)
        for g in self.optimizer.param_groups:
            g["lr"] = lr
        self.global_step += 1
        return {"loss": loss.detach(), "img_loss": img_loss.detach(), "latent_code_loss": latent_loss.detach(), "lr": lr}

    def _body(self, rays, bc_rgb, target, aud_feature, expr, latent_code, perturb):
        """Forward, loss, backward, gradient gather (+ all-reduce), Adam: everything between two learning-rate updates."""
        for p in self.flat.params:
            p.grad = None
        ret = self.net.render_rays(rays, bc_rgb, aud_feature, None, latent_code, expr, perturb=perturb)
        loss, img_loss, latent_loss = head_loss(ret, target, latent_code, self.args.lc_weight)
        loss.backward()
        self.flat.gather_grads(self.world, self.group)
        self.optimizer.step()
        for m in (self.net.face_nerf_coarse, self.net.face_nerf_fine):      # fused Adam does not bump parameter versions
            m.invalidate_packed()
        return loss, img_loss, latent_loss

    # -- CUDA-graph mode ------------------------------------------------------------------------------------------
    def _graphed(self, rays, bc_rgb, target, aud_feature, expr, index, perturb, aud_window):
        """The whole iteration -- conditioning nets, both render passes, loss, backward, gradient gather, Adam, learning-rate schedule, RNG
        offset -- as ONE cudaGraphLaunch.  Inputs are copied into static buffers; the latent-code row is selected on the device from a
        one-element index tensor; the learning rate and Adam's step count live in device tensors updated inside the graph; the
        stochastic draws come from the in-kernel Philox stream whose offset the graph advances.  The first call runs eagerly (it
        initialises Adam's state and the kernels' per-device constants), the second call captures, every later call replays."""
        dev = self.flat.flat.device
        use_win = aud_window is not None
        key = (tuple(rays.shape), tuple(bc_rgb.shape), use_win, float(perturb))
        if self._static is None or self._static["key"] != key:
            if self._static is not None:
                raise RuntimeError(f"TrainStep(cuda_graph=True) was captured for {self._static['key']}, got {key}: batch shapes, the "
                                   "aud_window / aud_feature choice and perturb are fixed per TrainStep")
            aud_src = aud_window if use_win else aud_feature
            self._static = {"key": key, "rays": torch.empty_like(rays), "bc": torch.empty_like(bc_rgb), "target": torch.empty_like(target),
                            "aud": torch.empty_like(aud_src.detach()), "expr": torch.empty_like(expr),
                            "idx": torch.zeros((1,), dtype=torch.int64, device=dev), "graph": None, "out": None, "calls": 0}
        st = self._static
        st["rays"].copy_(rays, non_blocking=True); st["bc"].copy_(bc_rgb, non_blocking=True); st["target"].copy_(target, non_blocking=True)
        st["aud"].copy_((aud_window if use_win else aud_feature).detach(), non_blocking=True); st["expr"].copy_(expr, non_blocking=True)
        if torch.is_tensor(index):
            st["idx"].copy_(index.reshape(1), non_blocking=True)
        else:
            st["idx"].fill_(int(index))

        def body():
            latent_code = self.latent_codes.index_select(0, st["idx"]).squeeze(0)
            aud = self.net.aud_att_net(self.net.aud_net(st["aud"])) if use_win else st["aud"]
            loss, img_loss, latent_loss = self._body(st["rays"], st["bc"], st["target"], aud, st["expr"], latent_code, perturb)
            # new_lrate = lrate * 0.1 ** (global_step / (lrate_decay * 1500)); global_step += 1      (:554-558), on the device
            self._lr_t.copy_(self.args.lrate * torch.pow(0.1, self._gs_t / (self.args.lrate_decay * 1500)))
            self._gs_t.add_(1.0)
            return loss.detach(), img_loss.detach(), latent_loss.detach()

        if st["calls"] == 0:
            st["out"] = body()                                       # eager: lazy initialisations happen here
        else:
            if st["graph"] is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st["out"] = body()
                st["graph"] = g
            st["graph"].replay()
        st["calls"] += 1
        self.global_step += 1
        loss, img_loss, latent_loss = st["out"]
        return {"loss": loss, "img_loss": img_loss, "latent_code_loss": latent_loss, "lr": self._lr_t}
