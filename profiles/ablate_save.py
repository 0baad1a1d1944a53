"""What does the activation saving cost the training forward?  Times inerf_mlp_fwd_train_bf16 on the fine pass of an N_rand = 3072 step
(3072 x 192 points) with parts of the saving switched off (csrc/mlp_bf16.cu, SAVE_OFF; the saved buffers are then garbage, only time counts),
then checks whether the cost depends on the data (zeroed weights; SM clock and board power sampled with nvidia-smi during 2 s loops).
Results: profiles/r01e_save_ablation.txt.

    build here :  INERF_SO=$PWD/build/libinerf_abl.so INERF_EXTRA_NVCC=-DINERF_ABLATION python profiles/ablate_save.py --build
    run on GPU :  INERF_SO=$PWD/build/libinerf_abl.so python profiles/ablate_save.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

if "--build" in sys.argv:
    M.build(force=True)
    print("built", M._lib.SO_PATH)
    sys.exit(0)

dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
a = M.default_args(dim_aud=64, dim_expr=76, perturb=1.0, mlp_mode="bf16", N_samples=64, N_importance=128, near=S.NEAR, far=S.FAR)
net = M.Network(450, 450, cam["focal"], S.NEAR, S.FAR, 1 << 20, None, 64, 128, args=a)
torch.manual_seed(1)
net.apply(M.init_weights)
net = net.to(dev).eval()
n_rand = int(os.environ.get("N_RAND", "3072"))
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)[:n_rand].contiguous()
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
f = net.face_nerf_fine
kp = f.kernel_params()
params = [p.detach() for p in kp]
packed = f.packed_weights(kp)
cond = ops.fold_cond(f._dims, params, aud, expr, lat)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    ms = timed(lambda: f.query(rays, z, aud, expr, lat))
print(f"render forward (no saving)                         {ms:7.3f} ms")
names = {0: "training forward, everything saved", 1: "no mask stores", 2: "no bulk copies of the activation images", 4: "no PE / view images",
         3: "no masks, no bulk copies", 7: "nothing saved (SAVE code paths only)"}
if "--waits" in sys.argv:      # round 3: is it the wait for the bulk copies' shared-memory reads (latency) or their bandwidth?
    names.update({8: "everything saved, NO wait for the copies' smem reads (rows overwritten early: garbage images)", 10: "no bulk copies, no waits"})
    for abl in (0, 8, 2, 1, 9):
        os.environ["INERF_SAVE_ABL"] = str(abl)
        ms = timed(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed, cond, rays, z))
        print(f"SAVE_ABL={abl} {names.get(abl, 'masks off + no waits'):64s} {ms:7.3f} ms")
    sys.exit(0)
for kabl, ktag in ((0, ""), (2, " + weight streaming off"), (4, " + sincosf off")):
    os.environ["INERF_ABL"] = str(kabl)
    for abl in (0, 1, 2, 4, 3, 7) if kabl == 0 else (0, 2, 7):
        os.environ["INERF_SAVE_ABL"] = str(abl)
        ms = timed(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed, cond, rays, z))
        print(f"SAVE_ABL={abl} {names[abl] + ktag:64s} {ms:7.3f} ms")

# ---- is the cost of saving data dependent (power)?  same kernel, zeroed weights; SM clock / board power sampled during a 2 s loop of each ----
import subprocess
import threading


def sampled(fn, secs=2.0):
    ms1 = timed(fn, 5)
    reps = max(10, int(secs * 1e3 / ms1))
    samples = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
            try:
                c, p = r.stdout.strip().split(",")
                samples.append((float(c), float(p)))
            except ValueError:
                pass
    th = threading.Thread(target=poll)
    th.start()
    ms = timed(fn, reps)
    stop.set()
    th.join()
    samples = samples[len(samples) // 3:] or [(0., 0.)]
    return ms, sorted(s[0] for s in samples)[len(samples) // 2], sorted(s[1] for s in samples)[len(samples) // 2]


os.environ["INERF_ABL"] = "0"
os.environ["INERF_SAVE_ABL"] = "0"
with torch.no_grad():
    ms, mhz, w = sampled(lambda: f.query(rays, z, aud, expr, lat))
    print(f"render forward, trained-scale weights      {ms:7.3f} ms   SM {mhz:.0f} MHz  {w:.0f} W")
ms, mhz, w = sampled(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed, cond, rays, z))
print(f"training forward, trained-scale weights    {ms:7.3f} ms   SM {mhz:.0f} MHz  {w:.0f} W")
os.environ["INERF_SAVE_ABL"] = "7"
ms, mhz, w = sampled(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed, cond, rays, z))
print(f"training forward, nothing saved            {ms:7.3f} ms   SM {mhz:.0f} MHz  {w:.0f} W")
os.environ["INERF_SAVE_ABL"] = "0"
with torch.no_grad():
    for p in params:
        p.zero_()
f.invalidate_packed()
packed0 = f.packed_weights(kp)
cond0 = ops.fold_cond(f._dims, params, aud, expr, lat)
ms, mhz, w = sampled(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed0, cond0, rays, z))
print(f"training forward, all-zero weights         {ms:7.3f} ms   SM {mhz:.0f} MHz  {w:.0f} W")
with torch.no_grad():
    ms, mhz, w = sampled(lambda: f.query(rays, z, aud, expr, lat))
    print(f"render forward, all-zero weights           {ms:7.3f} ms   SM {mhz:.0f} MHz  {w:.0f} W")
