import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import render_oracle as O
torch.set_num_threads(8)
g = dict(np.load(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'golden', 'render_3072.npz')))
def rnd(mode):
    if mode == 'bf16': return lambda x: x.to(torch.bfloat16).float()
    if mode == 'fp16': return lambda x: x.to(torch.float16).float()
    if mode == 'tf32':   # round to nearest, 10 explicit mantissa bits
        def f(x):
            i = x.contiguous().view(torch.int32)
            i = (i + 0x1000) & ~0x1FFF
            return i.view(torch.float32)
        return f
    if mode == 'tf32t':
        def f(x):
            i = x.contiguous().view(torch.int32) & ~0x1FFF
            return i.view(torch.float32)
        return f
    if mode == 'fp16x2':  # hi+lo
        def f(x):
            hi = x.to(torch.float16).float(); lo = (x-hi).to(torch.float16).float(); return hi+lo
        return f
    return lambda x: x
def emu_forward(sd, q, qa, pts_enc, dir_enc, aud, expr, lat):
    # q rounds weights, qa rounds activations
    cond = torch.cat([aud, expr / 3.0, lat]); W = lambda k: sd[k]
    x = qa(pts_enc)
    h = torch.relu(x @ q(W("pts_linears.0.weight")[:, :63]).T + W("pts_linears.0.weight")[:, 63:] @ cond + W("pts_linears.0.bias"))
    for l in range(1, 8):
        hq = qa(h); w = W(f"pts_linears.{l}.weight")
        if l == 5:
            C = cond.numel()
            pre = x @ q(w[:, :63]).T + w[:, 63:63 + C] @ cond + hq @ q(w[:, 63 + C:]).T + W(f"pts_linears.{l}.bias")
        else:
            pre = hq @ q(w).T + W(f"pts_linears.{l}.bias")
        h = torch.relu(pre)
    sigma = h @ W("alpha_linear.weight").T + W("alpha_linear.bias")
    w = W("views_linears.0.weight")
    v = torch.relu(qa(h) @ q(w[:, :256]).T + dir_enc @ w[:, 256:283].T + w[:, 283:] @ (expr / 3.0) + W("views_linears.0.bias"))
    for l in (1, 2):
        v = torch.relu(qa(v) @ q(W(f"views_linears.{l}.weight")).T + W(f"views_linears.{l}.bias"))
    rgb = v @ W("rgb_linear.weight").T + W("rgb_linear.bias")
    return torch.cat([rgb, sigma], -1)
def run(tag, wmode, amode):
    c, f = O.init_face_nerf(1), O.init_face_nerf(2)
    T = torch.from_numpy
    c["alpha_linear.weight"], c["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_c"]), T(g[f"{tag}_alpha_b_c"])
    f["alpha_linear.weight"], f["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_f"]), T(g[f"{tag}_alpha_b_f"])
    q, qa = rnd(wmode), rnd(amode)
    orig = O.face_nerf_forward
    def fwd(sd, x, aud, expr=None, latent=None):
        return emu_forward(sd, q, qa, x[:, :63], x[:, 63:], aud, expr, latent)
    O.face_nerf_forward = fwd
    try:
        r = O.render_rays(T(g["rays"]), T(g["bc_rgb"]), c, f, T(g["aud"]), T(g["expr"]), T(g["latent"]))
    finally:
        O.face_nerf_forward = orig
    out = {}
    for k in ("rgb_map", "acc_map", "rgb0", "acc0", "last_weight", "z_std"):
        out[k] = float((r[k] - T(g[f"{tag}_{k}"])).abs().max())
    d = float((1/r["disp_map"] - 1/T(g[f"{tag}_disp_map"])).abs().max())
    out["depth"] = d
    ref = T(g[f"{tag}_rgb_map"])
    out["psnr"] = float(-10*torch.log10(((r["rgb_map"]-ref)**2).mean()))
    return out
for tag in ("init", "dense"):
    for wm, am in (("none","none"),("bf16","bf16"),("fp16","fp16"),("tf32","tf32"),("tf32t","tf32t"),("fp16x2","fp16"),("fp16","fp16x2"),("fp16x2","fp16x2")):
        o = run(tag, wm, am)
        print(tag, wm, am, " ".join(f"{k}={v:.2e}" for k, v in o.items()), flush=True)
