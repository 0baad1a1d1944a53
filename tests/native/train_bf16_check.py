"""Debug driver (GPU): bf16 tensor-core training path against the fp32 kernels on the same network, stage by stage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, _lib
from oracle import render_oracle as O

dev = torch.device("cuda", 0)
torch.manual_seed(0)
b = O.synthetic_train_batch(0)
n, s = int(os.environ.get("N", 601)), int(os.environ.get("S", 64))
rays = b["rays"][:n].to(dev)
sd = O.init_face_nerf(7)
def mk(mode):
    net = M.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, dim_latent=32, dim_expr=76)
    net.load_state_dict(sd); net.mlp_mode = mode
    return net.to(dev)
n16, n32 = mk("bf16"), mk("fp32")
aud, expr, lat = b["aud"].to(dev), b["expr"].to(dev), b["latent"].to(dev)
z = ops.sample_coarse(rays, s, torch.rand(n, s, device=dev))
G = torch.randn(n, s, 4, device=dev) * torch.tensor([1., 1., 1., 0.3], device=dev)
P = n * s
n_tiles = ((P + 255) // 256) * 2

params = [p.detach() for p in n16.kernel_params()]
dims = n16._dims
cond = ops.fold_cond(dims, params, aud, expr, lat)
# ---- forward -----------------------------------------------------------------------------------------------------------------
raw16, acts, mask, _ = ops.mlp_fwd_train_bf16(dims, params, n16.packed_weights(n16.kernel_params()), cond, rays, z)
raw32, acts32, _ = ops.mlp_fwd_train(dims, params, cond, rays=rays, z=z)
torch.cuda.synchronize()
print("raw  bf16-train vs fp32: max-abs", float((raw16 - raw32).abs().max()), " vs bf16 inference:",
      float((raw16 - ops.mlp_fwd(_lib.INERF_MLP_BF16, dims, params, n16.packed_weights(n16.kernel_params()), cond, rays, z)).abs().max()))
A = ops.decode_images(acts, n_tiles)                                   # (T, 40, 128, 64)
def img_cols(A, first, n_img):                                         # images -> (T*128, 64*n_img)
    return A[:, first:first + n_img].permute(0, 2, 1, 3).reshape(-1, 64 * n_img)
def lay_img(l): return (4 * l, 4) if l < 8 else (32 + 2 * (l - 8), 2)
mw = mask.view(torch.int32).reshape(n_tiles, 76, 128)
def mask_bits(l):
    w0, nw = (8 * l, 8) if l < 8 else (64 + 4 * (l - 8), 4)
    words = mw[:, w0:w0 + nw].permute(0, 2, 1).reshape(-1, nw)        # (T*128, nw)
    sh = 31 - torch.arange(32, device=dev)
    return ((words[:, :, None] >> sh[None, None, :]) & 1).reshape(-1, nw * 32).bool()
for l in range(11):
    a16 = img_cols(A, *lay_img(l))[:P]
    agree = (mask_bits(l)[:P] == (a16 > 0)).float().mean()
    print(f"  layer {l}: mask agrees with saved activation > 0: {float(agree):.6f}   (act scale {float(a16.abs().max()):.2f}, zero frac {float((a16 == 0).float().mean()):.3f})")
# ---- backward ----------------------------------------------------------------------------------------------------------------
g16, dc16 = ops.mlp_bwd_bf16(dims, params, n16.packed_weights_bwd(n16.kernel_params()), aud, expr, lat, acts, mask, G, P, keep_deltas=True)
g32, dc32 = ops.mlp_bwd(dims, params, aud, expr, lat, acts32, G, P)
torch.cuda.synchronize()
D = ops.decode_images(ops.mlp_bwd_bf16.deltas, n_tiles)
# ---- self-consistency of the chain: delta_{l-1} = (delta_l . bf16(W_l)) * mask, from the kernel's own stored deltas ------------
sdc = {k: v.to(dev) for k, v in sd.items()}
C = 64 + 76 + 32
def Wl(l):
    if l < 8: return sdc[f"pts_linears.{l}.weight"]
    return sdc[f"views_linears.{l - 8}.weight"]
def bf(x): return x.to(torch.bfloat16).float()
Gp = torch.zeros(n_tiles * 128, 4, device=dev); Gp[:P] = G.reshape(-1, 4)
d10 = (bf(Gp[:, :3]) * 0 + Gp[:, :3]) @ sdc["rgb_linear.weight"]                       # fp32 d_rgb . W_rgb
d10 = torch.where(mask_bits(10), d10, torch.zeros_like(d10))
got = img_cols(D, *lay_img(10))
print(f"  chain d_v2 : max-abs err {float((got - bf(d10)).abs().max()):.3e}  (scale {float(d10.abs().max()):.2f})")
for l in range(10, 0, -1):                                                             # delta_l -> delta_{l-1}
    dl = img_cols(D, *lay_img(l))                                                      # stored (bf16) delta_l
    W = Wl(l)
    if l == 8: Wa = W[:, :256]
    elif l == 5: Wa = W[:, 63 + C:63 + C + 256]
    else: Wa = W
    dh = dl @ bf(Wa)
    if l == 8:                                                                         # + d_sigma (x) alpha weight (hi+lo = ~fp32)
        dh = dh + Gp[:, 3:4] * sdc["alpha_linear.weight"]
    ref = torch.where(mask_bits(l - 1), dh, torch.zeros_like(dh))
    got = img_cols(D, *lay_img(l - 1))
    err = float((got - bf(ref)).abs().max()); sc = float(ref.abs().max())
    print(f"  chain delta layer {l - 1}: max-abs err {err:.3e} (scale {sc:.2f})   rel-L2 {float((got - ref).norm() / ref.norm()):.3e}")
# ---- self-consistency of dW: grads = delta^T X from the stored images ---------------------------------------------------------
def chk(name, got, ref):
    print(f"  dW {name:26s} rel-L2 {float((got - ref).norm() / (ref.norm() + 1e-30)):.3e}  max-abs {float((got - ref).abs().max()):.3e} |ref| {float(ref.norm()):.2e}")
PE, DIR = img_cols(A, 38, 1), img_cols(A, 39, 1)
for l in range(11):
    dl = img_cols(D, *lay_img(l))
    if l == 0: X = PE[:, :63]; ref = dl.T @ X; got = g16[0][:, :63]
    elif l == 5: X = torch.cat([PE[:, :63], img_cols(A, *lay_img(4))], 1); ref = dl.T @ X; got = torch.cat([g16[10][:, :63], g16[10][:, 63 + C:]], 1)
    elif l == 8: X = torch.cat([img_cols(A, *lay_img(7)), DIR[:, :27]], 1); ref = dl.T @ X; got = g16[16][:, :283]
    else: X = img_cols(A, *lay_img(l - 1)); ref = dl.T @ X; got = g16[2 * l if l < 8 else 16 + 2 * (l - 8)]
    chk(f"layer {l} weight", got, ref)
    chk(f"layer {l} bias", g16[(2 * l if l < 8 else 16 + 2 * (l - 8)) + 1], dl.sum(0))
dout = img_cols(D, 38, 1)
chk("alpha weight", g16[22], (dout[:, 3:4].T @ img_cols(A, *lay_img(7))))
chk("rgb weight", g16[24], (dout[:, :3].T @ img_cols(A, *lay_img(10))))
chk("alpha bias", g16[23], dout[:, 3].sum(0, keepdim=True)); chk("rgb bias", g16[25], dout[:, :3].sum(0))
names = [k for k in n16.state_dict().keys() if not k.startswith("feature_linear")]
order = []
for i in range(8): order += [f"pts_linears.{i}.weight", f"pts_linears.{i}.bias"]
for i in range(3): order += [f"views_linears.{i}.weight", f"views_linears.{i}.bias"]
order += ["alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]
worst = 1.0
for nm, a, r in zip(order, g16, g32):
    cos = float(torch.nn.functional.cosine_similarity(a.flatten().double(), r.flatten().double(), dim=0))
    rel = float((a - r).norm() / (r.norm() + 1e-30))
    worst = min(worst, cos)
    print(f"  grad {nm:26s} cos {cos:.6f}  rel-L2 {rel:.3e}  |ref| {float(r.norm()):.3e}")
cosc = float(torch.nn.functional.cosine_similarity(dc16.double(), dc32.double(), dim=0))
print(f"  d_cond cos {cosc:.6f} rel {float((dc16 - dc32).norm() / dc32.norm()):.3e}")
print("delta_out image col 0..3 vs G:", float((img_cols(D, 38, 1)[:P, :4] - G.reshape(-1, 4)).abs().max()))
print("WORST COS", worst)
