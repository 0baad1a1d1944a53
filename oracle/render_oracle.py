"""CPU oracle for the IDEAL-NeRF ``render_rays`` hot path.

TEST INFRASTRUCTURE ONLY -- never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module.  The product
package (``ideal-nerf_b200/``) must not, and fails loudly when its CUDA library is missing.

What this is: a restatement, in torch-CPU fp32 tensor ops (the arithmetic library the reference
itself is written in), of the algorithm in /root/reference for the path SURVEY.md §8 scopes:

    stratified depths  -> NeRFs/HeadNeRF/train/audio_exp_nerf.py:306-328
    positional encoding-> NeRFs/HeadNeRF/helper.py:174-224
    FaceNeRF MLP       -> models/face_nerf.py:40-80
    run_network        -> NeRFs/HeadNeRF/train/audio_exp_nerf.py:376-394
    raw2outputs        -> NeRFs/HeadNeRF/train/baseline.py:325-375 (+ torso NeRFs/TorsoNeRF/test_torso.py:352-402)
    sample_pdf         -> NeRFs/HeadNeRF/helper.py:269-313
    render_rays        -> NeRFs/HeadNeRF/train/audio_exp_nerf.py:297-371
    get_rays           -> NeRFs/HeadNeRF/helper.py:228-243
    head/torso blend   -> NeRFs/TorsoNeRF/train_torso.py:269-270
    loss               -> NeRFs/HeadNeRF/train/audio_exp_nerf.py:540-548

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, produced by importing the unmodified reference in the
build container (``oracle/ref_import.py`` + ``tests/golden/make_golden.py``) and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here against them.

Everything is differentiable torch so that ``torch.autograd`` on this oracle is the backward
oracle as well (the reference's backward *is* autograd of these ops).

One piece is restated at the bit level because the acceptance gate is bit-exact: the
inverse-CDF index search of ``sample_pdf``.  ``torch.sum`` on CPU adds a 62-long fp32 row in a
specific vector order and ``torch.cumsum`` accumulates in fp64; ``sample_pdf_exact`` reproduces
both explicitly in numpy so that the oracle does not depend on the host ISA it happens to run on.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------


def get_rays(H, W, focal, c2w, cx=None, cy=None):
    """Pinhole rays for an HxW frame.  helper.py:228-243.

    Pixel (row j, col i) -> camera dir ((i-cx)/f, -(j-cy)/f, -1), rotated by c2w[:3,:3];
    origin is the camera centre c2w[:3,3] for every ray.
    """
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    cols = torch.linspace(0, W - 1, W)
    rows = torch.linspace(0, H - 1, H)
    jj, ii = torch.meshgrid(rows, cols, indexing="ij")       # jj: row index, ii: col index
    cx = W * .5 if cx is None else cx
    cy = H * .5 if cy is None else cy
    cam = torch.stack([(ii - cx) / focal, -(jj - cy) / focal, -torch.ones_like(ii)], -1)
    rays_d = torch.sum(cam[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def positional_encoding(x, n_freqs):
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    helper.py:174-204 with include_input=True, log_sampling=True; each block is len(x[-1]) wide.
    """
    bands = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs)
    parts = [x]
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def pack_rays(rays_o, rays_d, near, far):
    """rays (N,11) = [o, d, near, far, d/|d|].  audio_exp_nerf.py:409-427."""
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    ones = torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, near * ones, far * ones, viewdirs], -1)


def stratified_z(near, far, n_samples, n_rays, t_rand=None, lindisp=False):
    """Coarse depths.  audio_exp_nerf.py:306-328.

    near/far: (N,1) tensors.  t_rand: None (perturb == 0) or an (N,S) tensor of U[0,1) draws
    whose last column is forced to 1.0, as the reference does at :326.
    """
    t = torch.linspace(0., 1., steps=n_samples)
    if lindisp:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    else:
        z = near * (1. - t) + far * t
    z = z.expand([n_rays, n_samples])
    if t_rand is not None:
        mid = .5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], -1)
        lo = torch.cat([z[..., :1], mid], -1)
        t_rand = t_rand.clone()
        t_rand[..., -1] = 1.0
        z = lo + (hi - lo) * t_rand
    return z


# --------------------------------------------------------------------------------------------
# FaceNeRF
# --------------------------------------------------------------------------------------------

FACE_NERF_LAYERS = (["pts_linears.%d" % i for i in range(8)] + ["views_linears.%d" % i for i in range(3)]
                    + ["feature_linear", "alpha_linear", "rgb_linear"])


def face_nerf_shapes(dim_aud=64, dim_expr=76, dim_latent=32, W=256, in_xyz=63, in_views=27):
    """(out, in) of every nn.Linear in models/face_nerf.py:27-36, keyed like its state_dict."""
    cond = dim_aud + dim_expr + dim_latent
    shp = {"pts_linears.0": (W, in_xyz + cond)}
    for i in range(1, 8):
        shp["pts_linears.%d" % i] = (W, W + (in_xyz + cond if i == 5 else 0))
    shp["views_linears.0"] = (W // 2, in_views + W + dim_expr)
    shp["views_linears.1"] = (W // 2, W // 2)
    shp["views_linears.2"] = (W // 2, W // 2)
    shp["feature_linear"] = (W, W)
    shp["alpha_linear"] = (1, W)
    shp["rgb_linear"] = (3, W // 2)
    return shp


def init_face_nerf(seed, dim_aud=64, dim_expr=76, dim_latent=32):
    """xavier-uniform weights, bias 0.01 -- the reference's init_weights (audio_exp_nerf.py:442-448)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (o, i) in face_nerf_shapes(dim_aud, dim_expr, dim_latent).items():
        bound = math.sqrt(6.0 / (i + o))
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = torch.full((o,), 0.01)
    return sd


def face_nerf_forward(sd, x, aud, expr=None, latent=None):
    """models/face_nerf.py:40-80.  x: (P, 63+27) encoded [xyz | viewdir]; returns (P,4) = [rgb, sigma] raw."""
    in_xyz = 63
    pts, views = x[:, :in_xyz], x[:, in_xyz:]
    P = x.shape[0]
    cols = [pts]
    if aud is not None:
        cols.append(aud[None, :].expand(P, -1))
    if expr is not None:
        expr = expr * 1 / 3                                   # face_nerf.py:49
        cols.append(expr[None, :].expand(P, -1))
    if latent is not None:
        cols.append(latent[None, :].expand(P, -1))
    first = torch.cat(cols, -1)
    h = first
    for i in range(8):
        h = F.relu(F.linear(h, sd["pts_linears.%d.weight" % i], sd["pts_linears.%d.bias" % i]))
        if i == 4:
            h = torch.cat([first, h], -1)                     # skip, face_nerf.py:61-62
    sigma = F.linear(h, sd["alpha_linear.weight"], sd["alpha_linear.bias"])
    h = torch.cat([h, views], -1)
    if expr is not None:
        h = torch.cat([h, expr[None, :].expand(P, -1)], -1)
    for i in range(3):
        h = F.relu(F.linear(h, sd["views_linears.%d.weight" % i], sd["views_linears.%d.bias" % i]))
    rgb = F.linear(h, sd["rgb_linear.weight"], sd["rgb_linear.bias"])
    return torch.cat([rgb, sigma], -1)


def run_network(sd, pts, viewdirs, aud, expr, latent, netchunk=1024 * 64):
    """Encode + MLP in netchunk pieces.  audio_exp_nerf.py:376-394."""
    flat = pts.reshape(-1, 3)
    enc = positional_encoding(flat, 10)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    enc = torch.cat([enc, positional_encoding(dirs, 4)], -1)
    outs = [face_nerf_forward(sd, enc[i:i + netchunk], aud, expr, latent)
            for i in range(0, enc.shape[0], netchunk)]
    return torch.cat(outs, 0).reshape(list(pts.shape[:-1]) + [4])


# --------------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------------


def raw2outputs(raw, z_vals, rays_d, bc_rgb, noise=None, white_bkgd=False, with_fg=False):
    """Alpha compositing with the background injected as the last sample.

    baseline.py:325-375; ``with_fg`` adds the torso variant's rgb_map_fg (test_torso.py:393).
    noise: None or an (N,S) tensor already multiplied by raw_noise_std.
    Returns rgb_map, disp_map, acc_map, weights, depth_map[, rgb_map_fg].
    """
    gaps = z_vals[..., 1:] - z_vals[..., :-1]
    gaps = torch.cat([gaps, torch.full_like(gaps[..., :1], 1e10)], -1)
    gaps = gaps * torch.norm(rays_d[..., None, :], dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    colour = torch.cat((colour[:, :-1, :], bc_rgb.unsqueeze(1)), dim=1)
    sig = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-(F.relu(sig) + 1e-6) * gaps)
    trans = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * colour, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    if with_fg:
        rgb_fg = torch.sum(weights[:, :-1, None] * colour[:, :-1, :], -2)
        return rgb_map, disp_map, acc_map, weights, depth_map, rgb_fg
    return rgb_map, disp_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------------------------
# importance sampling
# --------------------------------------------------------------------------------------------


def sample_pdf(bins, weights, u):
    """helper.py:269-313 in torch ops.  u: (N, n_imp) -- torch.linspace(0,1,n) expanded (det) or draws.

    Returns (samples, inds).  Depends on the host's torch.sum order; the bit-level restatement is
    ``sample_pdf_exact``.
    """
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(inds - 1, min=0)
    hi = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
    bin_lo, bin_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    return bin_lo + t * (bin_hi - bin_lo), inds


def torch_cpu_rowsum_f32(x):
    """Bit-level model of ``torch.sum(x, -1)`` for a contiguous fp32 (N, n) array on CPU.

    ATen's inner-dim sum (aten/src/ATen/native/cpu/SumKernel.cpp, third-party, torch 2.x, as
    executed by helper.py:272) walks each row as 8-float vectors: groups of four vectors feed four
    vector accumulators, left-over vectors go to accumulator 0, accumulators 1..3 are then added to
    0; the scalar tail is summed left to right from 0.0 and finally the 8 lanes are added to it in
    lane order.  Verified equal to torch.sum on 200 000 random rows (tests/test_oracle_golden.py).
    """
    f = np.float32
    x = np.ascontiguousarray(x, dtype=f)
    n = x.shape[1]
    L = 8
    nv = n // L
    acc = [np.zeros((x.shape[0], L), f) for _ in range(4)]
    full = nv // 4
    for it in range(full):
        for k in range(4):
            v = it * 4 + k
            acc[k] = (acc[k] + x[:, v * L:(v + 1) * L]).astype(f)
    for v in range(full * 4, nv):
        acc[0] = (acc[0] + x[:, v * L:(v + 1) * L]).astype(f)
    for k in range(1, 4):
        acc[0] = (acc[0] + acc[k]).astype(f)
    total = np.zeros(x.shape[0], f)
    for j in range(nv * L, n):
        total = (total + x[:, j]).astype(f)
    for lane in range(L):
        total = (total + acc[0][:, lane]).astype(f)
    return total


def sample_pdf_exact(bins, weights, u):
    """Bit-level numpy restatement of helper.py:269-313 (fp32, CPU semantics).

    bins (N,B), weights (N,B-1), u (n_imp,) or (N,n_imp).  Returns samples (N,n_imp) fp32,
    inds (N,n_imp) int64 (= torch.searchsorted(cdf,u,right=True)), cdf (N,B) fp32.
    """
    f = np.float32
    bins = np.asarray(bins, f)
    w = (np.asarray(weights, f) + f(1e-5)).astype(f)
    tot = torch_cpu_rowsum_f32(w)
    pdf = (w / tot[:, None]).astype(f)
    run = np.zeros(w.shape[0], np.float64)                    # torch.cumsum on CPU carries fp64
    cdf = np.zeros((w.shape[0], w.shape[1] + 1), f)
    for j in range(w.shape[1]):
        run = run + pdf[:, j].astype(np.float64)
        cdf[:, j + 1] = run.astype(f)
    u = np.asarray(u, f)
    if u.ndim == 1:
        u = np.broadcast_to(u, (w.shape[0], u.shape[0]))
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)   # searchsorted right=True
    lo = np.maximum(inds - 1, 0)
    hi = np.minimum(inds, cdf.shape[1] - 1)
    rows = np.arange(w.shape[0])[:, None]
    cdf_lo, cdf_hi = cdf[rows, lo], cdf[rows, hi]
    bin_lo, bin_hi = bins[rows, lo], bins[rows, hi]
    denom = (cdf_hi - cdf_lo).astype(f)
    denom = np.where(denom < f(1e-5), f(1.0), denom).astype(f)
    t = ((u - cdf_lo).astype(f) / denom).astype(f)
    samples = (bin_lo + (t * (bin_hi - bin_lo).astype(f)).astype(f)).astype(f)
    return samples, inds, cdf


# --------------------------------------------------------------------------------------------
# the path
# --------------------------------------------------------------------------------------------


def render_rays(rays, bc_rgb, sd_coarse, sd_fine, aud, expr, latent, n_samples=64, n_importance=128,
                t_rand=None, u_rand=None, lindisp=False, white_bkgd=False, noise0=None, noise1=None,
                retraw=False, with_fg=False, netchunk=1024 * 64):
    """audio_exp_nerf.py:297-371 (torso extras: test_torso.py:269-349).

    rays (N,11).  perturb == 0  <=>  t_rand is None and u_rand is None (det=True).
    """
    N = rays.shape[0]
    rays_o, rays_d, viewdirs = rays[:, 0:3], rays[:, 3:6], rays[:, -3:]
    near, far = rays[:, 6:7], rays[:, 7:8]
    z = stratified_z(near, far, n_samples, N, t_rand, lindisp)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = run_network(sd_coarse, pts, viewdirs, aud, expr, latent, netchunk)
    out0 = raw2outputs(raw, z, rays_d, bc_rgb, noise0, white_bkgd, with_fg)
    ret = {}
    out = out0
    if n_importance > 0:
        w0 = out0[3]
        mid = .5 * (z[..., 1:] + z[..., :-1])
        u = u_rand if u_rand is not None else torch.linspace(0., 1., steps=n_importance).expand(N, n_importance)
        zs, _ = sample_pdf(mid, w0[..., 1:-1], u)
        zs = zs.detach()
        z, _ = torch.sort(torch.cat([z, zs], -1), -1)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        raw = run_network(sd_fine, pts, viewdirs, aud, expr, latent, netchunk)
        out = raw2outputs(raw, z, rays_d, bc_rgb, noise1, white_bkgd, with_fg)
        ret.update(rgb0=out0[0], disp0=out0[1], acc0=out0[2],
                   z_std=torch.std(zs, dim=-1, unbiased=False), last_weight=out[3][..., -1])
        if with_fg:
            ret.update(rgb_map_fg0=out0[5], last_weight0=out0[3][..., -1])
    ret.update(rgb_map=out[0], disp_map=out[1], acc_map=out[2])
    if with_fg:
        ret["rgb_map_fg"] = out[5]
    if retraw:
        ret["raw"] = raw
    ret["_z_vals"] = z
    ret["_weights"] = out[3]
    ret["_depth_map"] = out[4]
    return ret


def head_torso_blend(rgb_head, last_weight_torso, rgb_fg_torso):
    """train_torso.py:269-270 / test_torso.py:523."""
    return rgb_head * last_weight_torso[..., None] + rgb_fg_torso


def pose_to_euler_trans(poses):
    """run_nerf_helpers.py:26-47.  poses (B,3|4,4) -> (B,6) = [euler(3), trans(3)]."""
    R = poses[:, :3, :3]
    e = torch.stack([torch.atan2(R[:, 2, 2], R[:, 1, 2]), torch.asin(-R[:, 0, 2]),
                     torch.atan2(R[:, 0, 0], -R[:, 0, 1])], 1)
    return torch.cat((e, poses[:, :3, 3]), dim=1)


def head_loss(rgb, rgb0, target, latent, lc_weight=0.0005):
    """audio_exp_nerf.py:540-548: mse(rgb)+mse(rgb0)+10*lc_weight*||latent||_2."""
    return (torch.mean((rgb - target) ** 2) + torch.mean((rgb0 - target) ** 2)
            + 10 * lc_weight * torch.norm(latent))


# --------------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests, smoke and bench so all sides see the same data
# --------------------------------------------------------------------------------------------

NEAR = 0.5772005200386048
FAR = 1.1772005200386046


def synthetic_camera():
    c2w = torch.eye(4)[:3, :4].clone()
    c2w[2, 3] = 0.7772
    return dict(H=450, W=450, focal=1200., cx=225., cy=225., c2w=c2w)


def synthetic_frame(seed=0):
    """Full 450x450 frame: rays (202500,11), bc_rgb (202500,3), aud(64), expr(76), latent(32)."""
    cam = synthetic_camera()
    g = torch.Generator().manual_seed(seed)
    ro, rd = get_rays(cam["H"], cam["W"], cam["focal"], cam["c2w"], cam["cx"], cam["cy"])
    rays = pack_rays(ro, rd, NEAR, FAR)
    bc = torch.rand(cam["H"] * cam["W"], 3, generator=g)
    aud = torch.randn(64, generator=g)
    expr = torch.randn(76, generator=g)
    latent = torch.ones(32)
    return dict(rays=rays, bc_rgb=bc, aud=aud, expr=expr, latent=latent)


def synthetic_train_batch(seed=0, n_rand=3072, mouth_rays=512):
    """N_rand pixels: face box 95 %, outside 5 %, mouth box (mirrors audio_exp_nerf.py:163-187)."""
    fr = synthetic_frame(seed)
    g = torch.Generator().manual_seed(seed + 1)
    H = W = 450
    ii, jj = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    face = ((ii - 200).abs() < 100) & ((jj - 225).abs() < 100)
    mouth = ((ii - 250).abs() < 25) & ((jj - 225).abs() < 40)
    n_rest = n_rand - mouth_rays
    n_face = int(n_rest * 0.95)
    flat = torch.arange(H * W)

    def pick(mask, k):
        cand = flat[mask.reshape(-1)]
        return cand[torch.randperm(cand.numel(), generator=g)[:k]]

    idx = torch.cat([pick(face & ~mouth, n_face), pick(~face, n_rest - n_face), pick(mouth, mouth_rays)])
    target = torch.rand(n_rand, 3, generator=g)
    return dict(rays=fr["rays"][idx].contiguous(), bc_rgb=fr["bc_rgb"][idx].contiguous(), aud=fr["aud"],
                expr=fr["expr"], latent=fr["latent"], target=target, pixel_index=idx)


def normalise_density(sd, rays, aud, expr, latent, n_samples=64, target_std=8.0, max_rays=1024):
    """SURVEY.md §7-7 preset: rescale alpha_linear so raw sigma on the coarse samples is ~N(0.01, 8^2).

    Random-init FaceNeRF gives sigma ~ -0.3 +- 0.1 => every ray is pure background and parity would
    pass trivially.  Returns a new state dict.
    """
    sub = rays[:: max(1, rays.shape[0] // max_rays)][:max_rays]
    z = stratified_z(sub[:, 6:7], sub[:, 7:8], n_samples, sub.shape[0])
    pts = sub[:, None, 0:3] + sub[:, None, 3:6] * z[..., None]
    with torch.no_grad():
        s = run_network(sd, pts, sub[:, -3:], aud, expr, latent)[..., 3]
    k = target_std / float(s.std())
    out = dict(sd)
    out["alpha_linear.weight"] = sd["alpha_linear.weight"] * k
    out["alpha_linear.bias"] = (sd["alpha_linear.bias"] - float(s.mean())) * k + 0.01
    return out
