import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import render_oracle as O
torch.set_num_threads(8)
g = dict(np.load(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'golden', 'render_3072.npz')))
def f16(x): return x.to(torch.float16).float()
def split_rn(x):
    hi = f16(x); lo = f16(x - hi); return hi, lo
def split_trunc(x):
    i = x.contiguous().view(torch.int32) & ~0x1FFF
    hi = i.view(torch.float32); hi = f16(hi)   # exact unless subnormal/overflow
    lo = f16(x - i.view(torch.float32)); return hi, lo
def mm3(a, w, asplit, wsplit):
    ah, al = asplit(a); wh, wl = wsplit(w)
    return ah @ wh.T + al @ wh.T + ah @ wl.T
def emu_forward(sd, cfg, pts_enc, dir_enc, aud, expr, lat):
    asplit, wsplit, pe_single = cfg
    cond = torch.cat([aud, expr / 3.0, lat]); W = lambda k: sd[k]
    mm = lambda a, w: mm3(a, w, asplit, wsplit)
    if pe_single:
        xh = f16(pts_enc); mmx = lambda w: (lambda wh, wl: xh @ wh.T + xh @ wl.T)(*wsplit(w))
    else:
        mmx = lambda w: mm(pts_enc, w)
    h = torch.relu(mmx(W("pts_linears.0.weight")[:, :63]) + W("pts_linears.0.weight")[:, 63:] @ cond + W("pts_linears.0.bias"))
    for l in range(1, 8):
        w = W(f"pts_linears.{l}.weight")
        if l == 5:
            C = cond.numel()
            pre = mmx(w[:, :63]) + w[:, 63:63 + C] @ cond + mm(h, w[:, 63 + C:]) + W(f"pts_linears.{l}.bias")
        else:
            pre = mm(h, w) + W(f"pts_linears.{l}.bias")
        h = torch.relu(pre)
    sigma = h @ W("alpha_linear.weight").T + W("alpha_linear.bias")
    w = W("views_linears.0.weight")
    v = torch.relu(mm(h, w[:, :256]) + dir_enc @ w[:, 256:283].T + w[:, 283:] @ (expr / 3.0) + W("views_linears.0.bias"))
    for l in (1, 2):
        v = torch.relu(mm(v, W(f"views_linears.{l}.weight")) + W(f"views_linears.{l}.bias"))
    rgb = v @ W("rgb_linear.weight").T + W("rgb_linear.bias")
    return torch.cat([rgb, sigma], -1)
def run(tag, cfg):
    c, f = O.init_face_nerf(1), O.init_face_nerf(2)
    T = torch.from_numpy
    c["alpha_linear.weight"], c["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_c"]), T(g[f"{tag}_alpha_b_c"])
    f["alpha_linear.weight"], f["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_f"]), T(g[f"{tag}_alpha_b_f"])
    orig = O.face_nerf_forward
    O.face_nerf_forward = lambda sd, x, aud, expr=None, latent=None: emu_forward(sd, cfg, x[:, :63], x[:, 63:], aud, expr, latent)
    try:
        r = O.render_rays(T(g["rays"]), T(g["bc_rgb"]), c, f, T(g["aud"]), T(g["expr"]), T(g["latent"]))
    finally:
        O.face_nerf_forward = orig
    out = {}
    for k in ("rgb_map", "acc_map", "rgb0", "last_weight", "z_std"):
        out[k] = float((r[k] - T(g[f"{tag}_{k}"])).abs().max())
    out["depth"] = float((1/r["disp_map"] - 1/T(g[f"{tag}_disp_map"])).abs().max())
    return out
for tag in ("dense", "init"):
    for name, cfg in (("rn/rn", (split_rn, split_rn, False)), ("trunc/rn", (split_trunc, split_rn, False)),
                      ("rn/rn pe-single", (split_rn, split_rn, True)), ("trunc/rn pe-single", (split_trunc, split_rn, True))):
        o = run(tag, cfg)
        print(tag, name, " ".join(f"{k}={v:.2e}" for k, v in o.items()), flush=True)
