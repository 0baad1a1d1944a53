"""Pure-write, pure-read and copy HBM rates on this GPU (torch library kernels; context for the training kernels, which only write or only read).
    python profiles/hbm_rw.py"""
import json
import torch

dev = torch.device("cuda", 0)
n = 4 << 30
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
xf = x.view(torch.float32)


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
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
t = timed(lambda: x.zero_())
out["write_only_GBps"] = n / t / 1e9
t = timed(lambda: xf.fill_(1.5))
out["fill_f32_GBps"] = n / t / 1e9
t = timed(lambda: xf.sum())
out["read_only_GBps"] = n / t / 1e9
t = timed(lambda: y.copy_(x))
out["copy_read_plus_write_GBps"] = 2 * n / t / 1e9
print(json.dumps(out))
# non-constant data: broadcast a 1 MiB (L2-resident) pattern over the 4 GiB buffer
pat = torch.randint(0, 255, (1 << 20,), dtype=torch.uint8, device=dev)
t = timed(lambda: x.view(-1, 1 << 20).copy_(pat.expand(n >> 20, 1 << 20)))
print(json.dumps({"write_only_pattern_GBps": n / t / 1e9}))
