"""Golden fixture for the per-frame conditioning nets: outputs of the UNMODIFIED reference modules
models/audio_net.py::AudioNet / AudioAttNet (imported from /root/reference) on seeded inputs and seeded weights
(init_weights of NeRFs/HeadNeRF/train/audio_exp_nerf.py:442-448: xavier-uniform weights, bias 0.01, then a small seeded perturbation
of the biases so that they are not all equal).  Run once in the build container:  python tests/golden/make_golden_audio.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from models.audio_net import AudioNet, AudioAttNet      # noqa: E402  (the reference's own modules)


def init(m, g):
    if isinstance(m, (torch.nn.Linear, torch.nn.Conv1d)):
        torch.nn.init.xavier_uniform_(m.weight, generator=g)
        m.bias.data = 0.01 + 0.05 * torch.randn(m.bias.shape, generator=g)


def main():
    g = torch.Generator().manual_seed(77)
    out = {}
    for dim_aud in (64, 76):
        net = AudioNet(dim_aud, 16)
        net.apply(lambda m: init(m, g))
        x = torch.randn(8, 16, 29, generator=g)                      # the smoothing window: 8 frames of 16 x 29 DeepSpeech features
        with torch.no_grad():
            y = net(x)
            y1 = net(x[:1])                                          # single frame: .squeeze() drops the batch dimension
        for k, v in net.state_dict().items():
            out[f"an{dim_aud}.{k}"] = v.numpy()
        out[f"an{dim_aud}.x"], out[f"an{dim_aud}.y"], out[f"an{dim_aud}.y1"] = x.numpy(), y.numpy(), y1.numpy()
    att = AudioAttNet()                                              # dim_aud = 32, seq_len = 8 as the reference constructs it (:225)
    att.apply(lambda m: init(m, g))
    for d in (64, 76):
        xw = torch.randn(8, d, generator=g)
        with torch.no_grad():
            out[f"att.y{d}"] = att(xw).numpy()
        out[f"att.x{d}"] = xw.numpy()
    for k, v in att.state_dict().items():
        out[f"att.{k}"] = v.numpy()
    # ---- gradients: autograd of the unmodified modules (what loss.backward() computes for them, audio_exp_nerf.py:263-266,:549) ----
    # chain of the training path: window (8,16,29) -> AudioNet -> (8,64) -> AudioAttNet -> (64,) -> <g_aud, .>
    net = AudioNet(64, 16)
    net.load_state_dict({k[len("an64."):]: torch.from_numpy(v) for k, v in out.items()
                         if k.startswith("an64.") and k.split(".")[-1] in ("weight", "bias")})
    x = torch.from_numpy(out["an64.x"])
    g_aud = torch.randn(64, generator=g)
    codes = net(x)
    codes.retain_grad()
    feat = att(codes)
    (feat * g_aud).sum().backward()
    out["grad.g_aud"], out["grad.feat"], out["grad.d_codes"] = g_aud.numpy(), feat.detach().numpy(), codes.grad.numpy()
    for k, p in net.named_parameters():
        out[f"grad.an64.{k}"] = p.grad.numpy()
    for k, p in att.named_parameters():
        out[f"grad.att.{k}"] = p.grad.numpy()
    # the single-frame call before nosmo_iters (:266): AudioNet alone, n = 1
    net.zero_grad()
    g1 = torch.randn(64, generator=g)
    (net(x[3:4]) * g1).sum().backward()
    out["grad1.g"] = g1.numpy()
    for k, p in net.named_parameters():
        out[f"grad1.an64.{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "audio_nets.npz"), **out)
    print("wrote audio_nets.npz:", {k: v.shape for k, v in out.items() if k.endswith((".y", ".y1", "y64", "y76"))})


if __name__ == "__main__":
    main()
