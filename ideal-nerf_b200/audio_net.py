"""Host-side mirror of models/audio_net.py: AudioNet and AudioAttNet with the reference's constructor arguments and state_dict keys
(so `aud_net.*` / `aud_att_net.*` of a head.tar load unchanged); forward runs in the CUDA kernels of csrc/audio_net.cu.

Inference only: the kernels have no backward.  Training the conditioning nets stays with the reference's PyTorch modules (they are
0.02 % of a training step); calling these modules with autograd recording on their parameters raises."""
import ctypes

import torch
import torch.nn as nn

from . import _lib, ops


def _params12(mods):
    arr = (ctypes.c_void_p * 12)()
    keep = []
    for i, m in enumerate(mods):
        for j, p in enumerate((m.weight, m.bias)):
            ops._need_cuda(p, "parameter")
            t = p.detach()
            t = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
            keep.append(t)
            arr[2 * i + j] = t.data_ptr()
    return arr, keep


def _no_grad_only(mod):
    if torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters()):
        raise NotImplementedError(f"{type(mod).__name__}: the CUDA kernel is forward-only; call it under torch.no_grad() / after "
                                  "requires_grad_(False), or train this net with the reference module")


class AudioNet(nn.Module):
    """models/audio_net.py:43-69.  forward(x): x (n, 16, 29) -> (n, dim_aud), squeezed like the reference (n == 1 -> (dim_aud,))."""

    def __init__(self, dim_aud=76, win_size=16):
        super().__init__()
        if win_size != 16:
            raise ValueError("only win_size = 16 (the reference's only configuration) is built")
        self.win_size, self.dim_aud = win_size, dim_aud
        self.encoder_conv = nn.Sequential(
            nn.Conv1d(29, 32, 3, 2, 1), nn.LeakyReLU(0.02, True), nn.Conv1d(32, 32, 3, 2, 1), nn.LeakyReLU(0.02, True),
            nn.Conv1d(32, 64, 3, 2, 1), nn.LeakyReLU(0.02, True), nn.Conv1d(64, 64, 3, 2, 1), nn.LeakyReLU(0.02, True))
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        _no_grad_only(self)
        x = ops.f32c(x, "x")
        if x.dim() != 3 or x.shape[1:] != (16, 29):
            raise ValueError("AudioNet expects (n, 16, 29) DeepSpeech windows")
        n = x.shape[0]
        y = torch.empty((n, self.dim_aud), device=x.device)
        arr, keep = _params12([self.encoder_conv[0], self.encoder_conv[2], self.encoder_conv[4], self.encoder_conv[6],
                               self.encoder_fc1[0], self.encoder_fc1[2]])
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_net_fwd", _lib.lib().inerf_audio_net_fwd, arr, ops.ptr(x), n, self.dim_aud, ops.ptr(y), ops.stream())
        return y.squeeze()


class AudioAttNet(nn.Module):
    """models/audio_net.py:8-36.  forward(x): x (seq_len = 8, dim_feat) -> (dim_feat,)."""

    def __init__(self, dim_aud=32, seq_len=8):
        super().__init__()
        self.seq_len, self.dim_aud = seq_len, dim_aud
        self.attentionConvNet = nn.Sequential(
            nn.Conv1d(dim_aud, 16, 3, 1, 1), nn.LeakyReLU(0.02, True), nn.Conv1d(16, 8, 3, 1, 1), nn.LeakyReLU(0.02, True),
            nn.Conv1d(8, 4, 3, 1, 1), nn.LeakyReLU(0.02, True), nn.Conv1d(4, 2, 3, 1, 1), nn.LeakyReLU(0.02, True),
            nn.Conv1d(2, 1, 3, 1, 1), nn.LeakyReLU(0.02, True))
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len), nn.Softmax(dim=1))

    def forward(self, x):
        _no_grad_only(self)
        x = ops.f32c(x, "x")
        if x.dim() != 2 or x.shape[0] != self.seq_len or x.shape[1] < self.dim_aud:
            raise ValueError("AudioAttNet expects (seq_len, dim_feat >= dim_aud) audio codes")
        y = torch.empty((x.shape[1],), device=x.device)
        c = self.attentionConvNet
        arr, keep = _params12([c[0], c[2], c[4], c[6], c[8], self.attentionNet[0]])
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_att_fwd", _lib.lib().inerf_audio_att_fwd, arr, ops.ptr(x), self.seq_len, x.shape[1], self.dim_aud,
                     ops.ptr(y), ops.stream())
        return y
