"""Calibration of the fp16x2 kernel's accumulator compensation: raw (n, s, 4) of the kernel against an fp64 torch evaluation of the same
network on the same points, for a sweep of INERF_F16X2_COMP (main accumulator scaled by 1 + comp in the epilogue).  The tensor core adds
into its accumulator with round-toward-zero: every MMA shrinks the running sum by half an ulp on average, a multiplicative bias per layer.
python tests/native/f16x2_comp_sweep.py > profiles/r02_f16x2_comp.txt   (test infrastructure: uses oracle/ as the fp64 checker)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ideal_nerf_b200 as M
from oracle import render_oracle as O      # checker only (fp64 evaluation of the reference forward)

dev = "cuda:0"
b = O.synthetic_train_batch(0)
sd = O.normalise_density(O.init_face_nerf(7), b["rays"], b["aud"], b["expr"], b["latent"])
n, s = 1024, 64
rays = b["rays"][:n]
z = O.stratified_z(rays[:, 6:7], rays[:, 7:8], s, n)
pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None])
sd64 = {k: v.double().to(dev) for k, v in sd.items()}
pe = O.positional_encoding(pts.reshape(-1, 3).double().to(dev), 10)
de = O.positional_encoding(rays[:, None, 8:11].expand(n, s, 3).reshape(-1, 3).double().to(dev), 4)
ref = O.face_nerf_forward(sd64, torch.cat([pe, de], -1), b["aud"].double().to(dev), b["expr"].double().to(dev), b["latent"].double().to(dev)).reshape(n, s, 4)
scale = ref.abs().amax((0, 1))


def net(mode):
    f = M.FaceNeRF(dim_aud=64, dim_latent=32, dim_expr=76, mlp_mode=mode)
    f.load_state_dict(sd)
    return f.to(dev)


def report(tag, raw):
    e = (raw.double() - ref)
    print(f"{tag:28s} max/scale {[f'{float(v):.2e}' for v in e.abs().amax((0, 1)) / scale]}  rms/scale {[f'{float(v):.2e}' for v in (e ** 2).mean((0, 1)).sqrt() / scale]}"
          f"  mean signed sigma err / scale {float(e[..., 3].mean() / scale[3]):+.2e}")


args = (rays.to(dev), z.to(dev), b["aud"].to(dev), b["expr"].to(dev), b["latent"].to(dev))
with torch.no_grad():
    report("fp32 FFMA kernel", net("fp32").query(*args))
    fx = net("fp16x2")
    os.environ.pop("INERF_F16X2_COMP", None)
    report("fp16x2 per-layer table", fx.query(*args))
    for comp in ("0", "2e-7", "4e-7", "6e-7", "8e-7", "1e-6", "1.3e-6", "1.6e-6"):
        os.environ["INERF_F16X2_COMP"] = comp
        report(f"fp16x2 comp={comp}", fx.query(*args))
    os.environ.pop("INERF_F16X2_COMP", None)
    print("per-group ulps {L0, L1-7, V0, V1-2}:")
    best = []
    for l0 in (0, 1, 2):
        for tr in (3, 4, 5, 6):
            for v0 in (1, 2, 3, 4):
                for v12 in (0, 1, 2):
                    os.environ["INERF_F16X2_ULPS"] = f"{l0},{tr},{v0},{v12}"
                    e = fx.query(*args).double() - ref
                    rms = ((e ** 2).mean((0, 1)).sqrt() / scale)
                    best.append((float(rms.max()), float(rms.mean()), (l0, tr, v0, v12), [float(v) for v in e.abs().amax((0, 1)) / scale],
                                 [float(v) for v in e.mean((0, 1)) / scale]))
    best.sort()
    for b_ in best[:12]:
        print(f"  ulps {b_[2]}: rms max {b_[0]:.2e} mean {b_[1]:.2e}; max/scale {[f'{v:.2e}' for v in b_[3]]}; mean signed {[f'{v:+.1e}' for v in b_[4]]}")
