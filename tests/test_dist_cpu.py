"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gradient buckets, gathered logits."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uncltmo_b200 import dist as udist


def test_shard_range_partitions():
    for n in (1, 7, 60, 61, 1798):
        for world in (1, 2, 4, 8):
            spans = [udist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert udist.shard_tiles(60, 3, 8) == list(range(24, 32)) and udist.shard_tiles(60, 7, 8) == list(range(53, 60))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((300,), (17, 9), (5,), (64, 64))]
    params[2].requires_grad_(False)
    for i, p in enumerate(params):
        if p.requires_grad and i != 1:
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    buckets = udist.GradientBuckets(params, bucket_bytes=1024)   # params[1] has no grad yet: must be treated as zeros
    buckets.allreduce()
    ok = all(torch.allclose(params[i].grad, torch.full_like(params[i], 3.0 * (i + 1))) for i in (0, 3))
    ok = ok and torch.count_nonzero(params[1].grad) == 0 and params[2].grad is None
    # the same sums through the backward hooks (buckets launched as soon as their gradients exist)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))
    for i, p in enumerate(net.parameters()):
        torch.manual_seed(100 + i)
        p.data.normal_()
    net[2].bias.requires_grad_(False)
    hooked = udist.GradientBuckets(net.parameters(), bucket_bytes=64).install_hooks()
    xs = torch.arange(12.0).reshape(2, 6) / 10 + rank
    hooked.arm()
    net(xs).sum().backward()
    launched_early = len(hooked._inflight)
    hooked.finish()
    got = [p.grad.clone() for p in net.parameters() if p.requires_grad]
    for p in net.parameters():
        p.grad = None
    all_x = torch.cat([torch.arange(12.0).reshape(2, 6) / 10 + r for r in range(world)])
    net(all_x).sum().backward()
    want = [p.grad for p in net.parameters() if p.requires_grad]
    ok = ok and launched_early == len(hooked.buckets) and len(hooked.buckets) >= 2
    ok = ok and all(torch.allclose(a, b, atol=1e-5) for a, b in zip(got, want)) and net[2].bias.grad is None
    # gathered logits: global loss, local gradient
    t = torch.tensor([[float(rank)], [float(rank) + 0.5]], requires_grad=True)
    g = udist.all_gather_cat(t)
    (g * torch.arange(1, 5.0).view(4, 1)).sum().backward()
    ok = ok and g.shape == (4, 1) and torch.equal(g.detach().flatten(), torch.tensor([0.0, 0.5, 1.0, 1.5]))
    ok = ok and torch.equal(t.grad.flatten(), torch.tensor([1.0, 2.0]) + 2.0 * rank)
    # video scene sharded by tile chain: 7 tiles over 2 ranks (4 + 3), T = 3 frames
    full = torch.arange(3 * 7 * 2, dtype=torch.float32).reshape(3, 7, 2)
    lo, hi = udist.shard_range(7, rank, world)
    got = udist.gather_tile_chains(full[:, lo:hi].clone(), 7)
    ok = ok and torch.equal(got, full)
    try:
        udist.gather_tile_chains(full[:, :1].clone(), 7)
        ok = False
    except ValueError:
        pass
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_buckets_and_gather_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]


def _global_selection_worker(rank, world, port, out):
    """Global-batch selection under data parallelism (GanTrainerImg.py:341-408: arg-max / arg-min over the WHOLE batch):
    world-2 sum of the per-rank gradients must equal the single-process gradient of the same loss on the full batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    b, d = 3, 5
    full = torch.randn(world * b, d)
    scores = torch.tensor([0.3, 0.9, 0.1, 0.5, 0.95, 0.2])[:world * b]          # best = 4 (rank 1), worst = 2 (rank 0)

    def sim(a, q):       # stand-in for the nce similarity (the CUDA kernel cannot run here); any smooth pair function
        return (a * q / (1.0 + (a - q).abs())).sum(-1)

    def nce_like(anchor, pos, neg):
        lg = torch.stack([sim(anchor, pos), sim(anchor, neg)], 1)
        return torch.nn.functional.cross_entropy(lg, torch.zeros(anchor.shape[0], dtype=torch.long))

    # single process, full batch
    xf = full.clone().requires_grad_(True)
    gi = torch.stack([scores.argmax(), scores.argmin()])
    nce_like(xf, xf[gi[0]:gi[0] + 1], xf[gi[1]:gi[1] + 1]).backward()
    want = xf.grad[rank * b:(rank + 1) * b]
    # data parallel: local slice, global indices, rows broadcast from their owners, loss scaled by 1/world, grads summed
    x = full[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    g = udist.all_gather_flat(scores[rank * b:(rank + 1) * b])
    g_idx = torch.stack([g.argmax(), g.argmin()])
    rows = udist.BroadcastRowsFn.apply(x, g_idx)
    ok = torch.equal(rows.detach(), full[gi])
    (nce_like(x, rows[0:1], rows[1:2]) / world).backward()
    ok = ok and torch.allclose(x.grad, want, atol=1e-6)
    # pseudo-label form: mean_i |s_i - s_best| with the label's gradient owed to its owner
    stat = full[:, :1].clone()
    sf = stat.clone().requires_grad_(True)
    (sf - sf[gi[0]:gi[0] + 1]).abs().mean().backward()
    want_s = sf.grad[rank * b:(rank + 1) * b]
    s = stat[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    gs = udist.all_gather_cat(s)
    lab = gs.index_select(0, g_idx[:1])
    a_r = (s - lab.detach()).abs().mean()
    b_all = world * (gs.detach() - lab).abs().mean()
    ((a_r + (b_all - b_all.detach())) / world).backward()
    ok = ok and torch.allclose(s.grad, want_s, atol=1e-6)
    vals = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(vals, (a_r.detach() / world).reshape(1))
    ok = ok and torch.allclose(sum(vals), (stat - stat[gi[0]]).abs().mean().reshape(1), atol=1e-6)   # values add up too
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_global_batch_selection_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_global_selection_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]
