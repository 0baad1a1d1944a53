import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from ideal_nerf_b200 import ops
dev = torch.device("cuda", 0)
M, net, fr, cam = bench.build_network("bf16", dev)
from ideal_nerf_b200.frame import FrameRenderer
r = FrameRenderer(net, 0, 1)
res = {k: fr[k].to(dev) for k in ("pose", "aud", "expr", "latent")}
bc = fr["bc_rgb"].to(dev)
with torch.no_grad():
    for _ in range(3):
        r.render_frame(res["pose"], res["aud"], res["expr"], res["latent"], bc, perturb=1.0)
    torch.cuda.synchronize()
    # host time to enqueue one frame (GPU far behind: enqueue only)
    t0 = time.perf_counter()
    for _ in range(20):
        r.render_frame(res["pose"], res["aud"], res["expr"], res["latent"], bc, perturb=1.0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("host ms per render_frame enqueue:", (t1 - t0) / 20 * 1e3)
    t0 = time.perf_counter()
    for _ in range(20):
        net.render_dynamic_face(450, 450, net.focal, res["expr"], res["pose"], res["latent"], render_poses=res["pose"][:3, :4], chunk=1 << 20, near=net.near, far=net.far, bc_rgb=bc.reshape(450, 450, 3), aud_para=res["aud"], perturb=1.0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("host ms per render_dynamic_face enqueue:", (t1 - t0) / 20 * 1e3)
    # single-frame latency with sync
    lat = []
    for _ in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r.render_frame(res["pose"], res["aud"], res["expr"], res["latent"], bc, perturb=1.0)
        torch.cuda.synchronize(); lat.append((time.perf_counter() - t0) * 1e3)
    print("latency ms (sync per frame):", sorted(lat)[5])
