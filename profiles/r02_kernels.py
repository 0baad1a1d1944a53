"""One short program that launches every hot kernel of round 2 once or twice, for `ncu` (launch list and --set full captures):

    python profiles/r02_kernels.py [render|f16x2|sampling|train]

render: one 450x450 frame, bf16 mode (coarse + fine FaceNeRF launches, both compositor launches, both sampling kernels)
f16x2 : the same frame in the fp32-gate tensor-core mode
sampling: the two sampling kernels stand-alone on 202 500 rays
train : two training steps (config 3, bf16 kernels, eager TrainStep)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "render"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
if what in ("render", "f16x2"):
    mode = "bf16" if what == "render" else "fp16x2"
    _, net, fr, cam = bench.build_network(mode, dev)
    res = {k: fr[k].to(dev) for k in ("pose", "aud", "expr", "latent")}
    bc = fr["bc_rgb"].to(dev)
    with torch.no_grad():
        rays = ops.get_rays_packed(450, 450, net.focal, res["pose"][:3, :4], net.near, net.far)
        for _ in range(2):
            net.render_rays(rays, bc, res["aud"], None, res["latent"], res["expr"], perturb=1.0)
    torch.cuda.synchronize()
elif what == "sampling":
    with torch.no_grad():
        print(bench.sampling_standalone(M, dev, 202500, reps=2))
elif what == "train":
    print(bench.train_step_bench(M, dev, 2, "bf16"))
