"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL over NVLink for the only collective the path needs.

* Inference: image frames (and tiles) are independent - ranks take disjoint frames, no collective.
  Video scenes shard by TILE CHAIN, not by frame: the generator hands channel slices from frame k-1 to frame k of the
  same tile (Unet.py:229-286), so a rank owns a subset of the tiles for all frames of the scene.
* Training: data parallel over the batch; gradients are summed with bucketed all-reduces.  The reference's
  counterpart is nn.DataParallel (utils/model_save_util.py:50-54), i.e. scatter / replicate / gather every step.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [begin, end) of n items for `rank` (the first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_tiles(ntiles, rank, world):
    """Tile-chain ownership for a video scene."""
    return list(range(*shard_range(ntiles, rank, world)))


def gather_tile_chains(local, ntiles, group=None):
    """Video scene sharded by tile chain: every rank computed `local` = [T, n_local, ...] outputs for its contiguous
    `shard_range` of the `ntiles` tiles (all T frames); returns the full [T, ntiles, ...] on every rank.
    One all-gather of equal-sized (zero-padded) blocks - the only collective of the inference path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(ntiles, rank, world)
    if local.shape[1] != hi - lo:
        raise ValueError("rank %d owns tiles [%d, %d) but holds %d" % (rank, lo, hi, local.shape[1]))
    per = (ntiles + world - 1) // world
    block = local.new_zeros((local.shape[0], per) + tuple(local.shape[2:]))
    block[:, :hi - lo] = local
    block = block.contiguous()
    out = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(out, block, group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(ntiles, r, world)
        parts.append(out[r][:, :e - b])
    return torch.cat(parts, dim=1)


class GradientBuckets:
    """Flat fp32 buckets over a parameter list, built in REVERSE registration order (decoder first: the order
    in which backward produces gradients), all-reduced (sum) with one collective per bucket."""

    def __init__(self, params, bucket_bytes=8 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat = None
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._hooks, self._pending, self._inflight, self._comm = [], None, {}, None

    # ---------------------------------------------------------------- overlap with backward
    def install_hooks(self, group=None):
        """Launch a bucket's all-reduce as soon as backward has produced all of its gradients (buckets are in backward
        order), on a communication stream next to the remaining backward kernels.  Use `arm()` before backward() and
        `finish()` after it instead of `allreduce()`.  Works inside a CUDA-graph capture (NCCL branches are captured)."""
        if self._hooks:
            return self
        self._group = group
        for b in self.buckets:
            for p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def arm(self):
        self._pending = [len(b) for b in self.buckets]
        self._inflight = {}

    def _on_grad(self, p):
        if self._pending is None or not (dist.is_available() and dist.is_initialized()):
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        bucket = self.buckets[bi]
        cuda = bucket[0].is_cuda
        if cuda:
            if self._comm is None:
                self._comm = torch.cuda.Stream(bucket[0].device)
            main = torch.cuda.current_stream(bucket[0].device)
            self._comm.wait_stream(main)
            ctx = torch.cuda.stream(self._comm)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self._group, async_op=True)
        self._inflight[bi] = (flat, work)

    def finish(self, average=False):
        """Wait for the in-flight buckets, reduce the ones backward never completed (parameters without a gradient this
        step count as zeros) and write the sums back into .grad."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self._group) == 1:
            self._pending = None
            return
        world = dist.get_world_size(self._group)
        for bi in range(len(self.buckets)):
            if bi not in self._inflight:
                self._launch(bi)
        cuda = self.buckets[0][0].is_cuda
        main = torch.cuda.current_stream(self.buckets[0][0].device) if cuda else None
        for bi, (flat, work) in sorted(self._inflight.items()):
            if cuda:
                with torch.cuda.stream(self._comm):
                    work.wait()
            else:
                work.wait()
        if cuda:
            main.wait_stream(self._comm)
        for bi, (flat, work) in sorted(self._inflight.items()):
            if cuda:
                flat.record_stream(main)
            if average:
                flat /= world
            off = 0
            for p in self.buckets[bi]:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._pending, self._inflight = None, {}

    def allreduce(self, group=None, average=False, async_op=False):
        """Sum (or average) .grad over the process group.  Returns the list of work handles if async_op."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return []
        world = dist.get_world_size(group)
        works, flats = [], []
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
            flats.append((flat, bucket))
        for w in works:
            w.wait()
        for flat, bucket in flats:
            if average:
                flat /= world
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        return works


def all_gather_cat(t, group=None):
    """Concatenate a small per-rank tensor over ranks along dim 0 (logits / naturalness scores of the global batch).
    The local slice keeps its autograd history; the remote slices are constants - each rank differentiates the global
    loss with respect to its own samples and the gradient all-reduce adds the pieces up."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.detach().contiguous(), group=group)
    parts[rank] = t
    return torch.cat(parts, dim=0)


def is_parallel(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_gather_flat(t, group=None):
    """[world * len(t)] concatenation of a small 1-D per-rank tensor, no autograd (scores that only drive a selection)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * t.numel(), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t.detach().contiguous().reshape(-1), group=group)
    return out


class BroadcastRowsFn(torch.autograd.Function):
    """rows[j] = x[g_idx[j]] of the GLOBAL batch, on every rank, without the host learning who owns them.

    x: this rank's [b, ...] slice of a batch split evenly over the ranks; g_idx: device int64 [R] global sample indices
    (identical on all ranks).  Forward: every rank contributes its own rows (zeros for the others) to one all-reduce.
    Backward: the partial gradients of the rows are summed over the ranks (all-reduce) and the owner adds them to its
    samples - so a loss that uses a sample chosen over the global batch (infoNCE2's arg-max / arg-min TMQI naturalness,
    GanTrainerImg.py:384-408) back-propagates into that sample on the rank that holds it, as under nn.DataParallel."""

    @staticmethod
    def forward(ctx, x, g_idx, group=None):
        b = x.shape[0]
        rank = dist.get_rank(group)
        local = g_idx - rank * b
        mine = (local >= 0) & (local < b)
        rows = x.detach().index_select(0, local.clamp(0, b - 1))
        rows = rows * mine.view(-1, *([1] * (x.dim() - 1))).to(rows.dtype)
        dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)
        ctx.save_for_backward(local.clamp(0, b - 1), mine)
        ctx.meta = (x.shape, x.dtype, group)
        return rows

    @staticmethod
    def backward(ctx, d_rows):
        local, mine = ctx.saved_tensors
        shape, dtype, group = ctx.meta
        d_rows = d_rows.contiguous().float()
        dist.all_reduce(d_rows, op=dist.ReduceOp.SUM, group=group)
        dx = torch.zeros(shape, device=d_rows.device, dtype=torch.float32)
        dx.index_add_(0, local, d_rows * mine.view(-1, *([1] * (d_rows.dim() - 1))).to(d_rows.dtype))
        return dx.to(dtype), None, None
