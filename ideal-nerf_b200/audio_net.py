"""Host-side mirror of models/audio_net.py: AudioNet and AudioAttNet with the reference's constructor arguments and state_dict keys
(so `aud_net.*` / `aud_att_net.*` of a head.tar load unchanged); forward runs in the CUDA kernels of csrc/audio_net.cu.

Training: both modules are torch.autograd.Functions over the forward kernels and the backward kernels (inerf_audio_net_bwd,
inerf_audio_att_bwd), so the audio code that conditions FaceNeRF carries gradients back into these weights exactly as in the reference,
whose Adam optimises network.parameters() -- AudioNet and AudioAttNet included (audio_exp_nerf.py:263-266, :493)."""
import ctypes

import torch
import torch.nn as nn

from . import _lib, ops


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _f32(ps):
    return [p if (p.dtype == torch.float32 and p.is_contiguous()) else p.float().contiguous() for p in ps]


class _AudioNetFn(torch.autograd.Function):
    """AudioNet.forward as one kernel; backward = inerf_audio_net_bwd (forward recomputed in shared memory, no dx: the DeepSpeech
    features are data)."""

    @staticmethod
    def forward(ctx, x, dim_aud, *params):
        ps = _f32([p.detach() for p in params])
        n = x.shape[0]
        y = torch.empty((n, dim_aud), device=x.device)
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_net_fwd", _lib.lib().inerf_audio_net_fwd, _ptr_array(ps), ops.ptr(x), n, dim_aud, ops.ptr(y), ops.stream())
        ctx.save_for_backward(x, *ps)
        ctx.dim_aud = dim_aud
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *ps = ctx.saved_tensors
        dy = ops.f32c(dy, "dy")
        grads, _ = ops.zero_grads_like(ps)
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_net_bwd", _lib.lib().inerf_audio_net_bwd, _ptr_array(ps), _ptr_array(grads), ops.ptr(x), ops.ptr(dy),
                     x.shape[0], ctx.dim_aud, ops.stream())
        return (None, None, *grads)


class _AudioAttFn(torch.autograd.Function):
    """AudioAttNet.forward as one kernel; backward = inerf_audio_att_bwd (returns dx for the AudioNet codes and the parameter gradients)."""

    @staticmethod
    def forward(ctx, x, seq_len, dim_att, *params):
        ps = _f32([p.detach() for p in params])
        y = torch.empty((x.shape[1],), device=x.device)
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_att_fwd", _lib.lib().inerf_audio_att_fwd, _ptr_array(ps), ops.ptr(x), seq_len, x.shape[1], dim_att,
                     ops.ptr(y), ops.stream())
        ctx.save_for_backward(x, *ps)
        ctx.meta = (seq_len, dim_att)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *ps = ctx.saved_tensors
        seq_len, dim_att = ctx.meta
        dy = ops.f32c(dy, "dy")
        grads, _ = ops.zero_grads_like(ps)
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            ops.call("inerf_audio_att_bwd", _lib.lib().inerf_audio_att_bwd, _ptr_array(ps), _ptr_array(grads), ops.ptr(x), ops.ptr(dy),
                     seq_len, x.shape[1], dim_att, ops.ptr(dx), ops.stream())
        return (dx, None, None, *grads)


def _params_of(mods):
    out = []
    for m in mods:
        for p in (m.weight, m.bias):
            ops._need_cuda(p, "parameter")
            out.append(p)
    return out


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
        x = ops.f32c(x, "x")
        if x.dim() != 3 or x.shape[1:] != (16, 29):
            raise ValueError("AudioNet expects (n, 16, 29) DeepSpeech windows")
        ps = _params_of([self.encoder_conv[0], self.encoder_conv[2], self.encoder_conv[4], self.encoder_conv[6],
                         self.encoder_fc1[0], self.encoder_fc1[2]])
        return _AudioNetFn.apply(x, self.dim_aud, *ps).squeeze()


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
        x = ops.f32c(x, "x")
        if x.dim() != 2 or x.shape[0] != self.seq_len or x.shape[1] < self.dim_aud:
            raise ValueError("AudioAttNet expects (seq_len, dim_feat >= dim_aud) audio codes")
        c = self.attentionConvNet
        return _AudioAttFn.apply(x, self.seq_len, self.dim_aud, *_params_of([c[0], c[2], c[4], c[6], c[8], self.attentionNet[0]]))
