import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import render_oracle as O
torch.set_num_threads(8)
def f16(x): return x.to(torch.float16).to(x.dtype)
def split_rn(x):
    hi = f16(x); lo = f16(x - hi); return hi, lo
def split_trunc(x):
    xf = x.float().contiguous()
    i = xf.view(torch.int32) & ~0x1FFF
    hi32 = i.view(torch.float32)
    hi = f16(hi32).to(x.dtype); lo = f16((xf - hi32)).to(x.dtype); return hi, lo
def emu(sd, pts_enc, dir_enc, aud, expr, lat, asplit, wsplit, dt):
    sd = {k: v.to(dt) for k, v in sd.items()}
    pts_enc, dir_enc, aud, expr, lat = (t.to(dt) for t in (pts_enc, dir_enc, aud, expr, lat))
    cond = torch.cat([aud, expr / 3.0, lat]); W = lambda k: sd[k]
    def mm(a, w):
        if asplit is None: return a @ w.T
        ah, al = asplit(a); wh, wl = wsplit(w)
        return ah @ wh.T + al @ wh.T + ah @ wl.T
    x = pts_enc
    def r32(t): return t.float().to(dt)   # activations are fp32 values in the kernel
    h = r32(torch.relu(mm(x, W("pts_linears.0.weight")[:, :63]) + W("pts_linears.0.weight")[:, 63:] @ cond + W("pts_linears.0.bias")))
    for l in range(1, 8):
        w = W(f"pts_linears.{l}.weight")
        if l == 5:
            C = cond.numel()
            pre = mm(x, w[:, :63]) + w[:, 63:63 + C] @ cond + mm(h, w[:, 63 + C:]) + W(f"pts_linears.{l}.bias")
        else:
            pre = mm(h, w) + W(f"pts_linears.{l}.bias")
        h = r32(torch.relu(pre))
    sigma = h @ W("alpha_linear.weight").T + W("alpha_linear.bias")
    w = W("views_linears.0.weight")
    v = r32(torch.relu(mm(h, w[:, :256]) + dir_enc @ w[:, 256:283].T + w[:, 283:] @ (expr / 3.0) + W("views_linears.0.bias")))
    for l in (1, 2):
        v = r32(torch.relu(mm(v, W(f"views_linears.{l}.weight")) + W(f"views_linears.{l}.bias")))
    rgb = v @ W("rgb_linear.weight").T + W("rgb_linear.bias")
    return torch.cat([rgb, sigma], -1)
b = O.synthetic_train_batch(0)
sd = O.normalise_density(O.init_face_nerf(7), b["rays"], b["aud"], b["expr"], b["latent"])
n, s = 771, 64
rays = b["rays"][:n]
z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], s, n)
pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
dirs = rays[:, None, 8:11].expand(n, s, 3).reshape(-1, 3)
pe, de = O.positional_encoding(pts, 10), O.positional_encoding(dirs, 4)
ref = emu(sd, pe, de, b["aud"], b["expr"], b["latent"], None, None, torch.float64)
scale = ref.abs().amax(0)
for name, (a_, w_), dt in (("fp32 plain", (None, None), torch.float32), ("trunc/rn exact-accum(fp64)", (split_trunc, split_rn), torch.float64),
                        ("rn/rn exact-accum(fp64)", (split_rn, split_rn), torch.float64), ("trunc/rn fp32-accum", (split_trunc, split_rn), torch.float32)):
    out = emu(sd, pe, de, b["aud"], b["expr"], b["latent"], a_, w_, dt).double()
    print(name, [f"{float(e):.2e}" for e in ((out - ref).abs().amax(0) / scale)], "rms", [f"{float(e):.2e}" for e in (((out - ref)**2).mean(0).sqrt() / scale)])
