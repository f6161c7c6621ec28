"""Frame path: HDR frame -> tone-mapped LDR frame, entirely on the GPU.

Host-side mirror of the reference's inference driver (utils/model_save_util.py:293-407 run_model_on_single_image2,
:409-486 test_big_size_image2, :567-614 run_model_on_video, :488-565 test_big_size_image) with every per-pixel
operation done by kernels of libuncltmo_b200.so:  log-lambda normalise + replicate pad -> gather 256x256 tiles ->
generator on ALL tiles in one batch -> closed-form cross-fade blend -> on-device percentiles -> clamp / stretch /
back-to-colour / crop (-> optional 8-bit stretch).  No host synchronisation inside the path.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call

PATCH = 256


def padded_extent(n):
    """data_loader_util.py:135-157 (resize_im): always pads to 16*floor(n/16)+16."""
    return 16 * (n // 16) + 16


def tile_starts(length, overlap=64, patch=PATCH):
    """Tile start offsets along one axis (regular tiles, then one anchored at the far edge).
    model_save_util.py:416-440: `while patch*i - overlap*(i-1) < L`."""
    if length <= patch:
        raise ValueError("tiling needs a padded extent > %d px (got %d): the reference fails on such inputs too "
                         "(undersized edge tile, SURVEY.md §4)" % (patch, length))
    starts, i = [], 1
    while patch * i - overlap * (i - 1) < length:
        starts.append((patch - overlap) * (i - 1))
        i += 1
    return starts + [length - patch]


def axis_blend_table(length, overlap=64, patch=PATCH):
    """Closed form of the sequential cross-fade along one axis.

    Returns (starts, idx [L][K] int32, w [L][K] float32): position p of the blended result equals
    sum_k w[p][k] * tile[idx[p][k]][p - starts[idx[p][k]]].  Derived by pushing unit weights through the
    reference's update order (model_save_util.py:428-446 for columns, :448-481 for rows)."""
    starts = tile_starts(length, overlap, patch)
    nt = len(starts)
    wt = np.zeros((nt, length), dtype=np.float64)
    end = 0
    for j, s in enumerate(starts[:-1]):
        if j == 0:
            wt[0, s:s + patch] = 1.0
        else:
            for i in range(overlap):
                wt[:, s + i] *= (overlap - 1 - i) / (overlap - 1)
                wt[j, s + i] += i / (overlap - 1)
            wt[:, s + overlap:s + patch] = 0.0
            wt[j, s + overlap:s + patch] = 1.0
        end = s + patch
    last = starts[-1]
    rng = end - last
    if rng <= 1:
        raise ValueError("degenerate edge tile overlap (last_range=%d divides by zero in the reference, "
                         "model_save_util.py:443)" % rng)
    for i in range(rng):
        wt[:, last + i] *= (rng - 1 - i) / (rng - 1)
        wt[nt - 1, last + i] += i / (rng - 1)
    wt[:, end:] = 0.0
    wt[nt - 1, end:] = 1.0
    k = int((wt != 0).sum(axis=0).max())
    idx = np.zeros((length, k), dtype=np.int32)
    w = np.zeros((length, k), dtype=np.float32)
    for p in range(length):
        nz = np.nonzero(wt[:, p])[0]
        idx[p, :len(nz)] = nz
        w[p, :len(nz)] = wt[nz, p]
    return starts, idx, w


class _Plan:
    """Per-resolution constants: padded size, tile origins, blend tables (device resident)."""

    def __init__(self, h, w, overlap, device):
        self.h, self.w = h, w
        self.h1, self.w1 = padded_extent(h), padded_extent(w)
        ys, yidx, yw = axis_blend_table(self.h1, overlap)
        xs, xidx, xw = axis_blend_table(self.w1, overlap)
        k = max(yidx.shape[1], xidx.shape[1])

        def widen(a):
            out = np.zeros((a.shape[0], k), dtype=a.dtype)
            out[:, :a.shape[1]] = a
            return out

        self.k = k
        self.fused_ok = None   # uncl_frame_fused_supported, asked once
        self.ty, self.tx = len(ys), len(xs)
        self.ntiles = self.ty * self.tx
        origins = np.array([(y, x) for y in ys for x in xs], dtype=np.int32)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self.origins = dev(origins)
        self.yidx, self.yw, self.ystart = dev(widen(yidx)), dev(widen(yw)), dev(np.array(ys, dtype=np.int32))
        self.xidx, self.xw, self.xstart = dev(widen(xidx)), dev(widen(xw)), dev(np.array(xs, dtype=np.int32))


class FramePipeline:
    """HDR rgb frame [3,H,W] (fp32, CUDA) -> tone-mapped colour frame, through `generator` (uncltmo_b200 UNet)."""

    def __init__(self, generator, factor_coeff=0.1, overlap=64, max_tiles_per_batch=256):
        self.g = generator
        self.factor_coeff = factor_coeff
        self.overlap = overlap
        self.max_tiles = max_tiles_per_batch
        self.fused = True   # one cooperative launch per frame stage when the device / geometry allow it
        self.fused_stages = ("norm", "blend", "post")   # which stages take the fused kernel (all; tools/ measure subsets)
        self._plans = {}
        self._ws = None

    def plan(self, h, w, device):
        key = (h, w, self.overlap, str(device))
        if key not in self._plans:
            self._plans[key] = _Plan(h, w, self.overlap, device)
        return self._plans[key]

    def _workspace(self, device):
        if self._ws is None or self._ws.device != device:
            self._ws = torch.zeros(_lib.lib().uncl_frame_workspace_bytes(), dtype=torch.uint8, device=device)
        return self._ws

    # -- stages (each usable on its own; tests compare them one by one with the oracle) --
    def normalise_pad(self, rgb, lam):
        """-> (gray_log padded [H1,W1] fp32, stats [4] = (min rgb, min Y, max Y, -)).
        model_save_util.py:232-239 + data_loader_util.resize_im."""
        _, h, w = rgb.shape
        pl = self.plan(h, w, rgb.device)
        out = torch.empty((pl.h1, pl.w1), device=rgb.device, dtype=torch.float32)
        stats = torch.empty(4, device=rgb.device, dtype=torch.float32)
        call("uncl_frame_normalise_pad", rgb, h, w, float(lam * 255 * self.factor_coeff), out, pl.h1, pl.w1, stats,
             self._workspace(rgb.device))
        return out, stats

    def gather_tiles(self, gray_p, pl):
        tiles = torch.empty((pl.ntiles, 1, PATCH, PATCH), device=gray_p.device, dtype=torch.float32)
        call("uncl_tiles_gather", gray_p, pl.h1, pl.w1, pl.origins, pl.ntiles, tiles)
        return tiles

    def blend(self, tiles, pl):
        out = torch.empty((pl.h1, pl.w1), device=tiles.device, dtype=torch.float32)
        call("uncl_tiles_blend", tiles, pl.yidx, pl.yw, pl.ystart, pl.xidx, pl.xw, pl.xstart, pl.tx, pl.k, out,
             pl.h1, pl.w1)
        return out

    def run_generator(self, tiles):
        if tiles.shape[0] <= self.max_tiles:
            return self.g.tonemap_tiles(tiles)
        return torch.cat([self.g.tonemap_tiles(tiles[i:i + self.max_tiles])
                          for i in range(0, tiles.shape[0], self.max_tiles)])

    def percentiles(self, x, p_lo, p_hi, clamp=(-3.0e38, 3.0e38)):
        out = torch.empty(2, device=x.device, dtype=torch.float32)
        call("uncl_percentile_pair", x, x.numel(), float(clamp[0]), float(clamp[1]), float(p_lo), float(p_hi), out,
             self._workspace(x.device))
        return out

    def postprocess(self, fake_p, rgb, stats, pl):
        """percentile(0.5, 99.5) clamp -> stretch -> back to colour -> crop.  model_save_util.py:389-402."""
        pct = self.percentiles(fake_p, 0.5, 99.5)
        out = torch.empty((3, pl.h, pl.w), device=rgb.device, dtype=torch.float32)
        call("uncl_frame_postprocess", fake_p, pl.h1, pl.w1, rgb, pl.h, pl.w, stats, pct, out)
        return out

    def to_uint8(self, col):
        """hdr_image_util.save_gray_tensor_as_numpy_stretch without the file write: HWC uint8."""
        _, h, w = col.shape
        pct = self.percentiles(col, 0.1, 99.0, clamp=(0.0, 1.0))
        out = torch.empty((h, w, 3), device=col.device, dtype=torch.uint8)
        call("uncl_frame_to_u8", col, h, w, pct, out)
        return out

    # -- the same stages fused into one cooperative launch each (bit-identical values, 3 launches instead of 14) --
    def fused_ok(self, pl):
        """True when the fused stage kernels can run for this plan on the current device (uncl_frame_fused_supported)."""
        if not self.fused:
            return False
        if pl.fused_ok is None:
            pl.fused_ok = bool(_lib.lib().uncl_frame_fused_supported(pl.h, pl.w, pl.h1, pl.w1, pl.k))
        return pl.fused_ok

    def normalise_tiles(self, rgb, lam, pl, tiles=None):
        """normalise_pad + gather_tiles in one launch -> (tiles [T,1,256,256], stats [4]).  tiles: optional destination
        (a contiguous [T,1,256,256] slice of a larger batch)."""
        if tiles is None:
            tiles = torch.empty((pl.ntiles, 1, PATCH, PATCH), device=rgb.device, dtype=torch.float32)
        stats = torch.empty(4, device=rgb.device, dtype=torch.float32)
        call("uncl_frame_normalise_tiles", rgb, pl.h, pl.w, float(lam * 255 * self.factor_coeff), pl.h1, pl.w1, pl.origins,
             pl.ntiles, tiles, stats, self._workspace(rgb.device))
        return tiles, stats

    def blend_percentiles(self, tiles, pl, p_lo=0.5, p_hi=99.5):
        """blend + percentiles of the blended plane in one launch -> (fake_p [H1,W1], pct [2])."""
        out = torch.empty((pl.h1, pl.w1), device=tiles.device, dtype=torch.float32)
        pct = torch.empty(2, device=tiles.device, dtype=torch.float32)
        call("uncl_frame_blend_percentiles", tiles, pl.yidx, pl.yw, pl.ystart, pl.xidx, pl.xw, pl.xstart, pl.tx, pl.k, out,
             pl.h1, pl.w1, float(p_lo), float(p_hi), pct, self._workspace(tiles.device))
        return out, pct

    def post_uint8(self, fake_p, pct, rgb, stats, pl, want_col=False):
        """post-process + colour percentiles + 8-bit stretch in one launch -> HWC uint8 (and the fp32 colour frame)."""
        col = torch.empty((3, pl.h, pl.w), device=rgb.device, dtype=torch.float32) if want_col else None
        u8 = torch.empty((pl.h, pl.w, 3), device=rgb.device, dtype=torch.uint8)
        pct2 = torch.empty(2, device=rgb.device, dtype=torch.float32)
        call("uncl_frame_post_u8", fake_p, pl.h1, pl.w1, rgb, pl.h, pl.w, stats, pct, col, 0.1, 99.0, pct2, u8,
             self._workspace(rgb.device))
        return (u8, col) if want_col else u8

    def finish(self, out_tiles, rgb, stats, pl, uint8):
        """generator tiles -> tone-mapped frame: blend, percentile clamp / stretch, back to colour, (8-bit stretch)."""
        ok = self.fused_ok(pl)
        if ok and "blend" in self.fused_stages:
            fake_p, pct = self.blend_percentiles(out_tiles, pl)
        else:
            fake_p = self.blend(out_tiles, pl)
            pct = self.percentiles(fake_p, 0.5, 99.5)
        if ok and uint8 and "post" in self.fused_stages:
            return self.post_uint8(fake_p, pct, rgb, stats, pl)
        col = torch.empty((3, pl.h, pl.w), device=rgb.device, dtype=torch.float32)
        call("uncl_frame_postprocess", fake_p, pl.h1, pl.w1, rgb, pl.h, pl.w, stats, pct, col)
        return self.to_uint8(col) if uint8 else col

    # -- whole path --
    def tonemap(self, rgb, lam, uint8=False):
        """rgb [3,H,W] fp32 CUDA, lam = the image's lambda (f = lam*255*factor_coeff) -> [3,H,W] fp32 or HWC uint8."""
        if not (rgb.is_cuda and rgb.dtype == torch.float32 and rgb.dim() == 3 and rgb.shape[0] == 3):
            raise ValueError("tonemap expects a CUDA fp32 [3,H,W] tensor")
        rgb = rgb.contiguous()
        pl = self.plan(rgb.shape[1], rgb.shape[2], rgb.device)
        if self.fused_ok(pl) and "norm" in self.fused_stages:
            tiles, stats = self.normalise_tiles(rgb, lam, pl)
        else:
            gray_p, stats = self.normalise_pad(rgb, lam)
            tiles = self.gather_tiles(gray_p, pl)
        return self.finish(self.run_generator(tiles), rgb, stats, pl, uint8)

    def tonemap_frames(self, frames, lam, uint8=False):
        """Several frames of ONE resolution in one generator call: their tiles are independent, and at 60 tiles the deeper
        half of the network (12^2 ... 59^2 resolutions, the graph block) does not fill 148 SMs - 120 tiles per call cost
        1.71 ms per frame of generator time instead of 1.85 (tools/tiles_batch_sweep.py).  frames: list of [3,H,W] fp32
        CUDA tensors; lam: one lambda or one per frame.  Returns the list of results, identical to tonemap() of each."""
        if len(frames) == 1:
            return [self.tonemap(frames[0], lam if not isinstance(lam, (list, tuple)) else lam[0], uint8)]
        lams = list(lam) if isinstance(lam, (list, tuple)) else [lam] * len(frames)
        f0 = frames[0]
        for f in frames:
            if not (f.is_cuda and f.dtype == torch.float32 and f.dim() == 3 and f.shape == f0.shape):
                raise ValueError("tonemap_frames expects CUDA fp32 [3,H,W] tensors of one shape")
        frames = [f.contiguous() for f in frames]
        pl = self.plan(f0.shape[1], f0.shape[2], f0.device)
        nt = pl.ntiles
        tiles = torch.empty((len(frames) * nt, 1, PATCH, PATCH), device=f0.device, dtype=torch.float32)
        stats = []
        for i, (f, l) in enumerate(zip(frames, lams)):
            dst = tiles[i * nt:(i + 1) * nt]
            if self.fused_ok(pl) and "norm" in self.fused_stages:
                stats.append(self.normalise_tiles(f, l, pl, dst)[1])
            else:
                gray_p, st = self.normalise_pad(f, l)
                dst.copy_(self.gather_tiles(gray_p, pl))
                stats.append(st)
        out_tiles = self.run_generator(tiles)
        return [self.finish(out_tiles[i * nt:(i + 1) * nt], f, stats[i], pl, uint8) for i, f in enumerate(frames)]

    def tonemap_host_frames(self, host_frames, lam, out=None, sync=True, frames_per_batch=1):
        """Stream pinned HOST frames through the path: [3,H,W] fp32 each -> HWC uint8 host tensors.

        The host->device copy of frame i+1 and the device->host copy of result i-1 run on their own streams while
        frame i computes (double-buffered device staging), so a long sequence costs max(copy, compute) per frame.
        Every frame's input and output still cross PCIe inside the call.

        frames_per_batch > 1: that many consecutive frames share one generator call (tonemap_frames); the staging is
        double-buffered per batch.

        sync=True (default): the call returns when the LAST device->host copy has landed - the returned host tensors are
        complete and may be read or saved at once.  sync=False returns while copies are still in flight; the caller must
        wait on `self.last_d2h_event` (or synchronise the device) before touching the results."""
        if not host_frames:
            return []
        dev = next(self.g.parameters()).device
        compute = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_streams", None) is None or self._copy_streams[0].device != dev:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        h2d, d2h = self._copy_streams
        fpb = max(1, int(frames_per_batch))
        groups = [list(range(i, min(i + fpb, len(host_frames)))) for i in range(0, len(host_frames), fpb)]
        stage = [[torch.empty(host_frames[0].shape, device=dev, dtype=torch.float32) for _ in range(fpb)] for _ in range(2)]
        # the staging buffers come from the compute stream's allocator pool: kernels queued there by an earlier call may
        # still be reading the blocks they were carved from, so the first upload must order itself after them
        h2d.wait_stream(compute)
        for grp in stage:
            for t in grp:
                t.record_stream(h2d)
        ready = [torch.cuda.Event() for _ in range(2)]     # staging set i holds a fresh batch
        free = [torch.cuda.Event() for _ in range(2)]      # staging set i has been consumed
        results = []
        if out is None:
            h, w = host_frames[0].shape[1], host_frames[0].shape[2]
            out = [torch.empty((h, w, 3), dtype=torch.uint8).pin_memory() for _ in host_frames]

        def upload(g, b):
            for k, idx in enumerate(groups[g]):
                stage[b][k].copy_(host_frames[idx], non_blocking=True)
            ready[b].record(h2d)

        with torch.cuda.stream(h2d):
            upload(0, 0)
        for g, idxs in enumerate(groups):
            b = g & 1
            if g + 1 < len(groups):
                with torch.cuda.stream(h2d):
                    if g >= 1:
                        h2d.wait_event(free[1 - b])
                    upload(g + 1, 1 - b)
            compute.wait_event(ready[b])
            u8s = self.tonemap_frames(stage[b][:len(idxs)], lam, uint8=True)
            free[b].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                for idx, u8 in zip(idxs, u8s):
                    out[idx].copy_(u8, non_blocking=True)
                    u8.record_stream(d2h)
            results.extend(out[idx] for idx in idxs)
        self.last_d2h_event = torch.cuda.Event()
        self.last_d2h_event.record(d2h)
        compute.wait_stream(d2h)
        if sync:
            self.last_d2h_event.synchronize()   # blocks the HOST: the pinned results are complete on return
        return results

    def tonemap_clip(self, frames, lam, uint8=False, shard_tiles=False, shard_frames_out=False):
        """Video path (run_model_on_video, model_save_util.py:567-614): frames [T,3,H,W] fp32 CUDA of ONE scene, one
        lambda per scene.  `self.g` must be the video generator (UNetVideo): every tile is a chain over the T frames
        that hands its recurrent channel slices from frame to frame; tiles are independent of each other.

        shard_tiles=True (torch.distributed initialised, every rank holding the same frames): the scene is split by tile
        chain - a rank runs its contiguous share of the tiles through all T frames, one all-gather returns every rank
        the full set (splitting by frame would change the results), blend / post-process follow.

        shard_frames_out=True (with shard_tiles): blend / percentiles / back-to-colour / 8-bit stretch are per-frame work
        with no coupling between frames, so rank r finishes only frames r, r + world, ... and returns those (a list of
        (frame index, result)); nothing of the post-process is replicated.  Normalisation stays on every rank (it needs
        each frame's global min / max and costs one 25 MB read per frame), the recurrent tile chains are what is sharded."""
        if not (frames.is_cuda and frames.dtype == torch.float32 and frames.dim() == 4 and frames.shape[1] == 3):
            raise ValueError("tonemap_clip expects a CUDA fp32 [T,3,H,W] tensor")
        import torch.distributed as tdist
        from .dist import gather_tile_chains, shard_range
        frames = frames.contiguous()
        t_len = frames.shape[0]
        pl = self.plan(frames.shape[2], frames.shape[3], frames.device)
        if self.fused_ok(pl):
            norm = [self.normalise_tiles(frames[t], lam, pl) for t in range(t_len)]
            tiles = [tl for tl, _ in norm]
        else:
            norm = [self.normalise_pad(frames[t], lam) for t in range(t_len)]
            tiles = [self.gather_tiles(g, pl) for g, _ in norm]
        lo, hi = 0, pl.ntiles
        sharded = shard_tiles and tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
        if sharded:
            lo, hi = shard_range(pl.ntiles, tdist.get_rank(), tdist.get_world_size())
        outs = [[] for _ in range(t_len)]
        for i in range(lo, hi, self.max_tiles):
            j = min(i + self.max_tiles, hi)
            chain = self.g.tonemap_clip_tiles([tl[i:j] for tl in tiles])
            for t in range(t_len):
                outs[t].append(chain[t])
        per_frame = [torch.cat(o) if len(o) > 1 else o[0] for o in outs] if hi > lo else \
            [tiles[0].new_empty((0, 1, 256, 256)) for _ in range(t_len)]
        if sharded:
            full = gather_tile_chains(torch.stack(per_frame), pl.ntiles)
            per_frame = [full[t] for t in range(t_len)]
        res = []
        mine = range(t_len)
        if sharded and shard_frames_out:
            mine = range(tdist.get_rank(), t_len, tdist.get_world_size())
        for t in mine:
            res.append(self.finish(per_frame[t], frames[t], norm[t][1], pl, uint8))
        if sharded and shard_frames_out:
            return list(zip(mine, res))
        return torch.stack(res)
