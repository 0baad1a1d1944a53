"""World-size-2 checks of the ray partition + image gather (gloo on CPU; the render kernels themselves are GPU-only)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ideal_nerf_b200.frame import FrameRenderer, band
    lo, hi = band(n_rays, rank, world)
    full = torch.arange(n_rays * 3, dtype=torch.float32).reshape(n_rays, 3)     # the "rendered image"
    fr = FrameRenderer(network=None, rank=rank, world=world)
    out = fr.gather_image(full[lo:hi].clone(), n_rays)
    ok = bool(torch.equal(out, full))                     # all-gather: every rank holds the image
    # asynchronous, double-buffered form used by render_video: three frames in flight order, each consumed one frame late
    hs = [fr.gather_image((full[lo:hi] + k).clone(), n_rays, async_op=True) for k in range(2)]
    for k in range(2, 5):
        ok = ok and bool(torch.equal(hs[k - 2].wait(), full + (k - 2)))
        hs.append(fr.gather_image((full[lo:hi] + k).clone(), n_rays, async_op=True))
    ok = ok and bool(torch.equal(hs[3].wait(), full + 3)) and bool(torch.equal(hs[4].wait(), full + 4))
    if rank == 0:
        q.put((tuple(out.shape), ok))
    else:
        assert ok
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == world
    dist.destroy_process_group()


def test_band_partition_covers_every_ray_once():
    from ideal_nerf_b200.frame import band
    for n in (202500, 3072, 7, 1):
        for world in (1, 2, 3, 4, 8):
            spans = [band(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= hi - lo <= (n + world - 1) // world for lo, hi in spans)


def test_gather_image_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    for n_rays in (202500, 11):           # even split, and a ragged tail (6 + 5)
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        shape, ok = q.get(timeout=10)
        assert shape == (n_rays, 3) and ok
        port = _free_port()


def _train_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ideal_nerf_b200.frame import allreduce_grads, band
    torch.manual_seed(0)                                   # identical replica on every rank
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    x, y = torch.randn(3072, 5), torch.randn(3072, 3)      # the N_rand batch
    lo, hi = band(3072, rank, world)
    torch.mean((net(x[lo:hi]) - y[lo:hi]) ** 2).backward()
    allreduce_grads(net.parameters(), world)
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
        ref.load_state_dict(net.state_dict())
        torch.mean((ref(x) - y) ** 2).backward()
        q.put(max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters())))
    dist.destroy_process_group()


def test_allreduce_grads_world2_gloo_matches_full_batch():
    """Training rays sharded over 2 ranks + one flat gradient all-reduce == the single-process gradient (mean loss, equal bands)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=10) <= 1e-6


def _flat_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ideal_nerf_b200.frame import band
    from ideal_nerf_b200.train import FlatParams
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    ref.load_state_dict(net.state_dict())
    fp = FlatParams(list(net.parameters()))
    opt, opt_ref = torch.optim.Adam([fp.flat], lr=1e-2), torch.optim.Adam(ref.parameters(), lr=1e-2)
    x, y = torch.randn(3072, 5), torch.randn(3072, 3)
    lo, hi = band(3072, rank, world)
    for _ in range(3):
        for p in net.parameters():
            p.grad = None
        torch.mean((net(x[lo:hi]) - y[lo:hi]) ** 2).backward()
        fp.gather_grads(world)
        opt.step()
        opt_ref.zero_grad(set_to_none=True)
        torch.mean((ref(x) - y) ** 2).backward()
        opt_ref.step()
    if rank == 0:
        q.put(max(float((a - b).abs().max()) for a, b in zip(net.parameters(), ref.parameters())))
    dist.destroy_process_group()


def test_flat_params_data_parallel_adam_world2_gloo():
    """train.FlatParams: 2 ranks x half the batch, all-reduce of the flat gradient, one-tensor Adam == single-process Adam on the full batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=10) <= 1e-5
