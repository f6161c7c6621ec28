"""Training data path: the reference's dataset loaders with the per-crop work moved to the GPU.

Reference: utils/ProcessedDatasetFolderImg.py:44-168 and utils/ProcessedDatasetFolder.py:43-215 (`npy_loader`: .npy decode,
random square resize, random 256x256 crop, RGB->Y, LDR / log-lambda HDR normalisation, two crops per file),
:171-260 (`ProcessedDatasetFolder`), utils/data_loader_util.py (DataLoader construction).

What stays on the host: the `.npy` decode and the random draws - `draw_augment` makes the same `np.random` calls in the
same order as the reference, so a seeded run picks the same sizes and crop origins.  What moves to the device: resize,
crop, layout change, luminance, reductions and normalisation, one launch per batch instead of ~10 torch/cv2 calls per
crop (`uncl_sample_crop_resize`, `uncl_sample_normalise`).

`TrainBatchLoader` reads the next batch's files on a background thread into pinned staging memory, copies it on a side
stream and runs the kernels there, double-buffered, so the trainer's stream never waits for disk or PCIe.
`ShardedSampler` gives every rank an equal, disjoint slice of a per-epoch permutation (torch DistributedSampler
semantics; the reference uses nn.DataParallel with one loader).
"""
import os
import threading

import numpy as np
import torch

from ._lib import F32, call  # noqa: F401

PATCH = 256
NORMALISATIONS = {"hdr": 0, "max_normalization": 1, "bugy_max_normalization": 2, "stretch": 3}


def draw_augment(h, w, always_resize, rng=np.random):
    """The random draws of one crop, in the reference's call order (ProcessedDatasetFolderImg.py:61-78, 103-127).
    Returns (RH, RW, xx, yy): the size the image is resized to (== (h, w) if it is not) and the crop origin."""
    rh, rw = h, w
    if always_resize or h != PATCH:
        mode = rng.randint(0, 2)
        rh = PATCH if mode == 0 else int(rng.uniform(256, 512))
        rw = rh
    xx = yy = 0
    if rh != PATCH:
        xx = rng.randint(0, rw - PATCH)
        yy = rng.randint(0, rh - PATCH)
    elif rw != PATCH:
        raise ValueError("a 256-row image must be 256 wide (the reference crops nothing when h == 256)")
    return rh, rw, xx, yy


def draw_video_crop(h, w, rng=np.random):
    """real_video branch (ProcessedDatasetFolder.py:116-119): the frame keeps its size, only x is cropped."""
    if h != PATCH:
        raise ValueError("video frames are stored 256 rows high")
    return h, w, rng.randint(0, w - PATCH), 0


class ShardedSampler:
    """Per-epoch permutation, padded to a multiple of the world size, strided over ranks."""

    def __init__(self, n, world=1, rank=0, seed=0, shuffle=True):
        if not 0 <= rank < world:
            raise ValueError("rank %d outside world %d" % (rank, world))
        self.n, self.world, self.rank, self.seed, self.shuffle = n, world, rank, seed, shuffle
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return (self.n + self.world - 1) // self.world

    def indices(self):
        order = np.random.RandomState(self.seed + self.epoch).permutation(self.n) if self.shuffle else np.arange(self.n)
        total = len(self) * self.world
        order = np.concatenate([order, order[:total - self.n]])
        return order[self.rank:total:self.world].tolist()


def prepare_crops(src_base, src_off, meta, mode, f_per_sample=None, max_stretch=1.0, min_stretch=0.0):
    """Device part of `npy_loader` for S crops at once.

    src_base: 1-D fp32 CUDA buffer holding the source images (HWC) back to back; src_off int64 [S] (offsets in floats);
    meta int32 [S, 8] rows {H, W, RH, RW, xx, yy, 0, 0}; mode: key of NORMALISATIONS; f_per_sample fp32 [S] (HDR).
    Returns (input [S,1,256,256], color [S,3,256,256], gray_norm, gray_shift) - the last two are None for LDR modes."""
    if not src_base.is_cuda:
        raise ValueError("prepare_crops runs on CUDA tensors only")
    s = meta.shape[0]
    dev = src_base.device
    color = torch.empty((s, 3, PATCH, PATCH), device=dev, dtype=torch.float32)
    inp = torch.empty((s, 1, PATCH, PATCH), device=dev, dtype=torch.float32)
    hdr = NORMALISATIONS[mode] == 0
    gnorm = torch.empty_like(inp) if hdr else None
    gshift = torch.empty_like(inp) if hdr else None
    stats = torch.empty(2 * s, device=dev, dtype=torch.float32)
    call("uncl_sample_crop_resize", src_base, src_off, meta, color, s, PATCH)
    call("uncl_sample_normalise", color, s, PATCH, NORMALISATIONS[mode], f_per_sample, float(max_stretch), float(min_stretch),
         inp, gnorm, gshift, stats)
    return inp, color, gnorm, gshift


class TrainBatchLoader:
    """Iterates batches shaped like the reference DataLoader's: dicts with `input_im` [B,2,1,256,256], `color_im`
    [B,2,3,256,256], `original_gray_norm`, `original_gray` [B,2,1,256,256] (HDR mode; aliases of `input_im` otherwise, as in
    the reference) and `gamma_factor` [B] (the brightness factor lambda*255*factor_coeff, 0 in LDR mode)."""

    def __init__(self, paths, batch_size, hdr_mode, ldr_neg_mode=False, lambdas=None, factor_coeff=0.1,
                 normalization="max_normalization", max_stretch=1.0, min_stretch=0.0, device="cuda", world=1, rank=0, seed=0,
                 shuffle=True, drop_last=True, loader=np.load):
        if hdr_mode and lambdas is None:
            raise ValueError("HDR mode needs the per-image lambda dictionary (f_train_dict_path of the reference)")
        self.paths = list(paths)
        self.batch_size, self.hdr_mode, self.ldr_neg_mode = batch_size, hdr_mode, ldr_neg_mode
        self.lambdas, self.factor_coeff = lambdas, factor_coeff
        self.mode = "hdr" if hdr_mode else normalization
        if self.mode not in NORMALISATIONS:
            raise ValueError("unknown normalization %r" % (normalization,))
        self.max_stretch, self.min_stretch = max_stretch, min_stretch
        self.device = torch.device(device)
        self.sampler = ShardedSampler(len(self.paths), world, rank, seed, shuffle)
        self.drop_last = drop_last
        self.loader = loader
        self.copy_stream = torch.cuda.Stream(self.device)
        self._stage = [None, None]      # pinned staging buffers, grown on demand

    def __len__(self):
        n = len(self.sampler)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def set_epoch(self, epoch):
        self.sampler.set_epoch(epoch)

    # -------------------------------------------------------------- host side of one batch
    def _brightness(self, path):
        name = os.path.splitext(os.path.basename(path))[0]
        if name not in self.lambdas:
            raise KeyError("no lambda found for file %s" % name)   # get_f, ProcessedDatasetFolderImg.py:25-35
        return float(self.lambdas[name]) * 255 * self.factor_coeff

    def _read(self, idxs, slot):
        """Decode the files of one batch and lay them out back to back in pinned memory; two crops per file."""
        arrays, metas, offs, fs = [], [], [], []
        off = 0
        for i in idxs:
            img = np.ascontiguousarray(self.loader(self.paths[i]), dtype=np.float32)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("%s: expected an [H, W, 3] array" % self.paths[i])
            h, w = img.shape[0], img.shape[1]
            f = self._brightness(self.paths[i]) if self.hdr_mode else 0.0
            for _ in range(2):   # `for k in range(2)` of npy_loader: two independent crops of the same file
                rh, rw, xx, yy = draw_augment(h, w, always_resize=self.ldr_neg_mode)
                metas.append((h, w, rh, rw, xx, yy, 0, 0))
                offs.append(off)
                fs.append(f)
            arrays.append(img)
            off += img.size
        need = off
        if self._stage[slot] is None or self._stage[slot].numel() < need:
            self._stage[slot] = torch.empty(max(need, 1 << 20), dtype=torch.float32).pin_memory()
        flat = self._stage[slot].numpy()
        o = 0
        for a in arrays:
            flat[o:o + a.size] = a.reshape(-1)
            o += a.size
        return need, np.asarray(metas, np.int32), np.asarray(offs, np.int64), np.asarray(fs, np.float32)

    def _upload_and_prepare(self, host, slot):
        need, metas, offs, fs = host
        with torch.cuda.stream(self.copy_stream):
            src = self._stage[slot][:need].to(self.device, non_blocking=True)
            meta = torch.from_numpy(metas).pin_memory().to(self.device, non_blocking=True)
            off = torch.from_numpy(offs).pin_memory().to(self.device, non_blocking=True)
            f = torch.from_numpy(fs).pin_memory().to(self.device, non_blocking=True)
            inp, color, gnorm, gshift = prepare_crops(src, off, meta, self.mode, f, self.max_stretch, self.min_stretch)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        b = len(fs) // 2
        shape = lambda t: t.reshape(b, 2, *t.shape[1:])  # noqa: E731
        batch = {"input_im": shape(inp), "color_im": shape(color),
                 "original_gray_norm": shape(gnorm) if gnorm is not None else shape(inp),
                 "original_gray": shape(gshift) if gshift is not None else shape(inp),
                 "gamma_factor": f[::2]}
        return batch, done, (src, meta, off)

    def __iter__(self):
        idx = self.sampler.indices()
        bs = self.batch_size
        batches = [idx[i:i + bs] for i in range(0, len(idx), bs)]
        if self.drop_last and batches and len(batches[-1]) < bs:
            batches.pop()
        if not batches:
            return
        result = {}

        def read(k, slot):
            try:
                result[k] = self._read(batches[k], slot)
            except BaseException as e:  # surfaced on the consumer thread
                result[k] = e

        thread = threading.Thread(target=read, args=(0, 0))
        thread.start()
        staged_free = [None, None]   # event after which the pinned slot may be overwritten
        for k in range(len(batches)):
            thread.join()
            host = result.pop(k)
            if isinstance(host, BaseException):
                raise host
            slot = k & 1
            batch, done, keep = self._upload_and_prepare(host, slot)
            staged_free[slot] = done
            if k + 1 < len(batches):
                nslot = (k + 1) & 1
                if staged_free[nslot] is not None:
                    staged_free[nslot].synchronize()   # the H2D copy out of that pinned slot has finished
                thread = threading.Thread(target=read, args=(k + 1, nslot))
                thread.start()
            torch.cuda.current_stream(self.device).wait_event(done)
            for t in batch.values():
                t.record_stream(torch.cuda.current_stream(self.device))
            yield batch
