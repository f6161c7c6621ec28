#!/usr/bin/env python
"""Headline benchmark: 1080p tone-mapped frames/s through the B200 UnCLTMO hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32]

One "step" = one synthetic 1080p HDR frame through the whole frame path: log-lambda normalise + pad ->
60 tiles of 256x256 -> U-Net generator (incl. bottleneck graph block) on all tiles -> cross-fade blend ->
percentile clamp / stretch / back-to-colour / crop -> 8-bit stretch.  N > 1 shards frames over ranks (one
process per GPU, no data-path collective: image frames are independent - SURVEY.md §8e), weak scaling.

Printed JSON (rank 0, one line):
  value      frames/s with the HDR frame already resident in HBM (device-timed, max over ranks)
  e2e        frames/s through the public API with HOST buffers: pinned H2D of the frame + D2H of the 8-bit result
             inside the timed region
  roofline   tensor-core roofline of the dominant kernel (conv3x3_tc, all 3x3 conv / ConvTranspose layers):
             algorithmic FLOP per frame of those layers / device time spent in them, against the measured peak
  cpu_baseline  the CPU oracle (PyTorch-on-CPU restatement of the reference, `oracle/`) on this box's host cores,
             on a bounded sample of the same workload
--impl reference times that CPU arm alone (the reference has no GPU-independent build to install: it IS the
PyTorch code path the oracle restates).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
TILES = 60
LAMBDA = 371.4          # median of lambda_data/input_images_lambdas.npy
GFLOP_TILE = 18.2858    # generator forward per 256x256 tile
# MACs per tile of the layers that run in conv3x3_tc: total 9.1429 GMAC minus inc.conv (0.0186), the four
# ConvTranspose k2 s2 (0.0377+0.0514+0.0610+0.0650), the graph block (0.0566) and outc (0.0021)  
GFLOP_TILE_TC = 2 * (9.1429 - 0.0186 - 0.2151 - 0.0566 - 0.0021)
METRIC = "1080p tone-mapped frames/s (UNet fwd)"
CONFIG = {"workload": "image TMO inference, one 1920x1080 HDR frame per step per GPU = 60 tiles of 256x256 "
                      "(pad to 1088x1936, overlap 64), random-init weights, lambda 371.4",
          "frames_per_generator_call": "4 (default --frames-per-call): consecutive frames share one generator call of 240 "
                                       "independent tiles; every frame still runs its own normalise / blend / percentile / "
                                       "post-process stages and K steps = exactly K frames",
          "tiles_per_frame": TILES, "gflop_per_frame": TILES * GFLOP_TILE,
          "l2": "3 frames rotate per rank and each step writes/reads >2 GB of activations (>> 126 MB L2), "
                "so no input survives in L2 between steps",
          "parallelism": "frames sharded over ranks, no collective",
          "second_metric": "256^2 train steps/s is reported in the `train` object of the same line"}
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


def measured_peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, source).  A kernel timed per launch in the instrumented pass
    runs under burst conditions (short, cool, full clocks): its roofline denominator is the BURST figure."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        burst = d.get("bf16_tflops", d.get("bf16_tflops_sustained"))
        return burst, d.get("bf16_tflops_sustained", burst), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
    return 1650.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index, power=False):
        self.power = power
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has produced its first sample: its start-up (NVML initialisation) can stall kernel
        launches for 100-200 ms, which must not fall into a timed region of a few tens of milliseconds."""
        if self.p is None:
            return self
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    break
            except OSError:
                break
            time.sleep(0.02)
        return self

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, watts = [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1]))
                watts.append(float(r[3]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        if self.power and watts:
            out["power_w_median"], out["power_w_max"] = float(np.median(watts)), float(np.max(watts))
            out["sm_mhz_min"] = float(np.min(sm))
        return out


def cpu_reference_arm(steps, warmup, threads=None):
    """The CPU oracle on the host cores.  One step = one WHOLE 1080p frame through the oracle's restatement of
    run_model_on_single_image2: log-lambda normalise, pad, the generator tile by tile at batch 1 (60 tiles, as the
    reference runs them, utils/model_save_util.py:417-427), sequential cross-fade, np.percentile, back-to-colour, 8-bit
    stretch.  Nothing is extrapolated."""
    import oracle
    from uncltmo_b200 import synth
    from uncltmo_b200.weights import make_generator_state_dict
    torch.set_grad_enabled(False)
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = make_generator_state_dict()
    frames = [torch.from_numpy(synth.hdr_frame(H, W, seed=i)) for i in range(2)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        col = oracle.tonemap_frame(frames[i % 2], LAMBDA, lambda t: oracle.unet_forward(sd, t)[0])
        oracle.frame_path.to_uint8_stretch(col)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t_frame = float(np.mean(times))
    return {"value": 1.0 / t_frame, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": "%d whole 1080p frames (60 tiles at batch 1 + the sequential frame path each), mean %.2f s/frame, "
                      "%d warm-up frame(s)" % (len(times), t_frame, warmup),
            "ms_per_frame": t_frame * 1e3}, float(np.sum(times))


TRAIN_GFLOP_STEP = 4 * GFLOP_TILE * 16   # G fwd (D step) + G fwd (G step) + one merged backward, 16 images


def video_inference_workload(dev, precision, world, rank, frames=8, iters=2):
    """BASELINE.json configs[3]: video TMO on a synthetic 1080p clip (frame k = frame 0 translated + 1 % noise, SURVEY.md
    section 8d).  The recurrent generator couples the frames of a scene, not its tiles: with N GPUs the scene is split
    by tile chain (one all-gather per scene), so this is STRONG scaling of one scene."""
    from uncltmo_b200 import synth
    from uncltmo_b200.frame import FramePipeline
    from uncltmo_b200.generator import UNetVideo
    from uncltmo_b200.weights import make_generator_state_dict
    import torch.distributed as dist
    net = UNetVideo(*G_ARGS, up_mode=0, precision=precision).to(dev).eval()
    net.load_state_dict(make_generator_state_dict())
    pipe = FramePipeline(net)
    clip = torch.from_numpy(synth.hdr_clip(frames, H, W, seed=0)).to(dev)
    with torch.no_grad():
        pipe.tonemap_clip(clip, LAMBDA, uint8=True, shard_tiles=world > 1, shard_frames_out=world > 1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            pipe.tonemap_clip(clip, LAMBDA, uint8=True, shard_tiles=world > 1, shard_frames_out=world > 1)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return {"metric": "1080p video tone-mapped frames/s (recurrent UNet fwd, one scene)", "value": frames * iters / (ms / 1e3),
            "unit": "frames/s", "ms_per_frame": ms / (frames * iters), "scaling": "strong", "dtype": precision,
            "config": {"clip": "%d frames of 1920x1080, 60 tile chains" % frames,
                       "parallelism": "tile chains sharded over %d rank(s), one NCCL all-gather per scene, blend / post-process of the frames "
                                      "dealt round-robin over the ranks (each 8-bit frame is finished on one rank)" % world}}


def _release():
    """Drop the previous workload's networks, captured graphs and their private memory pools (reference cycles through
    autograd contexts keep them alive until a collection) before the next workload allocates its own."""
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def train_workload(dev, precision, steps, warmup, world=1, rank=0, video=False, weak=False):
    """256x256 image-TMO training step, global batch 8x2 = 16 images (GanTrainerImg.train_D + train_G, epoch-0 loss
    schedule, Adam as main_train_image.py builds it).  Strong scaling: ranks split the 16 images."""
    from uncltmo_b200 import _lib, synth
    from uncltmo_b200.discriminator import SimpleDiscriminator
    from uncltmo_b200.generator import UNet, UNetVideo
    from uncltmo_b200.trainer import GanTrainerStep
    from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
    import torch.distributed as dist
    _release()
    netG = (UNetVideo if video else UNet)(*G_ARGS, up_mode=0, precision=precision).to(dev).train()
    netG.load_state_dict(make_generator_state_dict())
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).to(dev).train()
    netD.load_state_dict(make_discriminator_state_dict())
    flat = precision == "bf16" and not video
    if flat:   # Adam on the flat parameter buffer of the bf16 training path: one launch (uncltmo_b200/optim.py)
        from uncltmo_b200.optim import FlatAdam
        optG = FlatAdam(netG, lr=1e-5, betas=(0.5, 0.999))
    else:
        optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999), capturable=True, fused=True)
    optD = torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999), capturable=True, fused=True)
    tr = GanTrainerStep(netG, netD, optG, optD)
    b_local = 8 if weak else max(1, 8 // world)
    mk = lambda a: torch.from_numpy(a).reshape(b_local, 2, 1, 256, 256)  # noqa: E731
    host = [(mk(synth.normalised_batch(2 * b_local, seed=40 + 10 * rank + i)).pin_memory(),
             mk(synth.ldr_batch(2 * b_local, seed=50 + 10 * rank + i)).pin_memory(),
             mk(synth.ldr_batch(2 * b_local, seed=60 + 10 * rank + i)).pin_memory()) for i in range(2)]
    devb = [tuple(t.to(dev) for t in h) for h in host]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(fn):
        for i in range(warmup):
            fn(i)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    def resident(i):
        h, p, n = devb[i % 2]
        tr.step(h, None, p, n, 0)

    loss_host = torch.empty(2).pin_memory()

    def e2e(i):
        h, p, n = (t.to(dev, non_blocking=True) for t in host[i % 2])
        g, s = tr.step(h, None, p, n, 0)
        loss_host.copy_(torch.stack([g.detach(), s.detach()]), non_blocking=True)

    # eager launches on a side stream: autograd binds the gradient-accumulation nodes of the parameters to the stream of
    # their first use, and the legacy default stream cannot take part in the capture that follows
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        ms_eager = run(resident)
    torch.cuda.current_stream(dev).wait_stream(side)
    # the whole iteration (D step, G step, optimizers, all-reduces) as one CUDA graph: what the step costs once the
    # ~1000 launches per iteration no longer go through Python
    _lib.reset_launch_count()
    tr.capture(devb[0][0], None, devb[0][1], devb[0][2], 0, warmup=1)
    launches = _lib.launch_count() // 2 * steps     # warm-up iteration + captured iteration were counted

    def resident(i):  # noqa: F811
        h, p, n = devb[i % 2]
        tr.replay(h, None, p, n, 0)

    def e2e(i):  # noqa: F811
        g, s = tr.replay(*(host[i % 2][0], None, host[i % 2][1], host[i % 2][2]), 0)   # pinned host -> captured buffers
        loss_host.copy_(torch.stack([g.detach(), s.detach()]), non_blocking=True)

    ms = run(resident)
    ms_e2e = run(e2e)
    bytes_in = 3 * b_local * 2 * 256 * 256 * 4
    what = "video TMO, 8 clips x 2 consecutive frames (recurrent generator, GanTrainer.py)" if video else \
        "16 images/step: train_D + train_G"
    return {"metric": "256^2 train steps/s (%s)" % what, "value": steps / (ms / 1e3), "unit": "steps/s",
            "ms_per_step": ms / steps, "scaling": "weak (16 images per GPU)" if weak else "strong",
            "images_per_s": steps * 2 * b_local * world / (ms / 1e3), "execution": "CUDA graph replay of the whole iteration",
            "eager_launch_path": {"value": steps / (ms_eager / 1e3), "unit": "steps/s", "ms_per_step": ms_eager / steps,
                                  "note": "same trainer without capture; launch-bound, and Adam(capturable=True) is slower eagerly"}, "dtype": "f32" if precision == "fp32" else ("f32 (3x3 convolutions as three-term bf16 split tensor-core GEMMs, ~2^-16)" if precision == "fp32_tc" else ("bf16 activations and gradient operands, f32 accumulation / parameters / optimizer" if flat else "bf16 operands / f32 tensors (mixed)")),
            "tflops_algorithmic": steps * TRAIN_GFLOP_STEP / ms, "gpu_launches": launches,
            "e2e": {"value": steps / (ms_e2e / 1e3), "unit": "steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 8},
            "config": {"global_batch": "8 x 2 crops = 16 images of 256x256", "per_gpu_images": 2 * b_local, "loss_schedule": "epoch 0",
                       "optimizer": ("uncltmo_b200.optim.FlatAdam (G, one launch) + torch Adam fused (D)" if flat else "torch.optim.Adam(fused=True, capturable=True)") + ", lr 1e-5 / 1.5e-5, betas (0.5, 0.999)",
                       "droppath": "p = 0.05, fresh masks in both generator passes (reference schedule)",
                       "parallelism": "dp%d, NCCL gradient all-reduce%s" % (world, " of the flat gradient buffer" if flat else "")}}


def cpu_train_arm(threads=None):
    """The CPU oracle's training step (same schedule, torch autograd on the host cores) on a 2-image sample, x8."""
    import oracle
    from uncltmo_b200 import synth
    from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    g_sd, d_sd = make_generator_state_dict(), make_discriminator_state_dict()
    hdr = torch.from_numpy(synth.normalised_batch(16, seed=4))
    pos, neg = torch.from_numpy(synth.ldr_batch(16, seed=5)), torch.from_numpy(synth.ldr_batch(16, seed=6))
    with torch.enable_grad():
        oracle.train_step_losses(g_sd, d_sd, hdr[:2], pos[:2], neg[:2], 0)          # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        oracle.train_step_losses(g_sd, d_sd, hdr, pos, neg, 0)                       # the whole 16-image step, nothing extrapolated
        dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": "one whole D+G step of the oracle on the 16-image batch (%.2f s), after a 2-image warm-up step" % dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, _ = cpu_reference_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_frame"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload; the reference runs its generator tile by tile at batch 1 (utils/model_save_util.py:417-427)
            "config": dict({k: v for k, v in CONFIG.items() if k != "frames_per_generator_call"},
                           frames_per_generator_call="n/a: the reference calls its generator once per tile (batch 1)"),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--frames-per-call", type=int, default=4,
                    help="consecutive frames that share one generator call (their tiles are independent); 1 = frame by frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    ap.add_argument("--train-precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sustain-s", type=float, default=2.5, help="seconds of back-to-back frames for the `sustained` field (0: skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    from uncltmo_b200 import _lib, synth
    from uncltmo_b200.frame import FramePipeline
    from uncltmo_b200.generator import UNet
    from uncltmo_b200.weights import make_generator_state_dict
    torch.set_grad_enabled(False)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints one JSON line
        # rank 0 prints exactly one JSON line on stdout: NCCL / torch print their version banner on file descriptor 1 when
        # the first communicator is created, so that happens here with fd 1 pointed at stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    net = UNet(*G_ARGS, up_mode=0, precision=args.precision).to(dev).eval()
    net.load_state_dict(make_generator_state_dict())
    pipe = FramePipeline(net)
    # a few distinct frames per rank so no step re-reads a frame that is still warm in L2 from the previous one
    nframes = 3
    host_frames = [torch.from_numpy(synth.hdr_frame(H, W, seed=100 * rank + i)).pin_memory() for i in range(nframes)]
    dev_frames = [f.to(dev) for f in host_frames]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    loop = {"n": 0}     # length of the loop that is running (the last generator call of a loop takes the frames that are left)

    def timed(fn, steps):
        loop["n"] = args.warmup
        for i in range(args.warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loop["n"] = steps
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    FPC = max(1, args.frames_per_call)

    def batch_call(i, count=None):
        """`count` (default FPC) consecutive frames (steps i .. i + count - 1) through the path with ONE generator call
        (their tiles are independent; at 60 tiles the deeper half of the network does not fill the GPU)."""
        pipe.tonemap_frames([dev_frames[(i + k) % nframes] for k in range(count or FPC)], LAMBDA, uint8=True)

    def step_resident(i):
        # a step is one frame; frames are processed FPC at a time, so every FPC-th step does the work of up to FPC steps -
        # exactly K frames are processed in a loop of K steps
        if i % FPC == 0:
            batch_call(i, min(FPC, loop["n"] - i))

    host_outs = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(8)]

    def e2e_run(steps):
        """K frames from pinned host memory to 8-bit results in pinned host memory through the streaming API."""
        frames = [host_frames[i % nframes] for i in range(steps)]
        outs = [host_outs[i % len(host_outs)] for i in range(steps)]
        pipe.tonemap_host_frames(frames, LAMBDA, out=outs, frames_per_batch=FPC)

    def timed_e2e(steps):
        e2e_run(args.warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local).wait_first() if rank == 0 else None
    _lib.reset_launch_count()
    ms_res = timed(step_resident, args.steps)
    launches = _lib.launch_count() * args.steps // (args.steps + args.warmup)   # (kernel launches, averaged over warm-up + timed frames)
    # guard against a one-off host / driver stall inside the short timed region (seen once: 180 ms in a 36 ms loop right
    # after another process had released the GPU): time the same K steps once more; if the first try was > 1.3x slower, the
    # second one is reported and the first is kept in the line as `first_try_ms_per_step`
    first_try = None
    ms_again = timed(step_resident, args.steps)
    if ms_res > 1.3 * ms_again:
        first_try, ms_res = ms_res / args.steps, ms_again
    ms_e2e = timed_e2e(args.steps)
    clocks = sampler.stop() if sampler else None

    # endurance: the same step back to back for >= args.sustain_s seconds of device time, clocks and power sampled through it
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / (ms_res / args.steps)) + 1)
        sus_sampler = ClockSampler(local, power=True).wait_first() if rank == 0 else None
        ms_sus = timed(step_resident, n_sus)
        sus_clocks = sus_sampler.stop() if sus_sampler else None
        sustained = {"value": world * n_sus / (ms_sus / 1e3), "unit": "frames/s", "frames_per_gpu": n_sus, "seconds": ms_sus / 1e3,
                     "ms_per_step": ms_sus / n_sus, "clocks": sus_clocks}

    # per-kernel device time (separate instrumented pass: an event pair around every C-ABI call)
    roof = roofline_hbm = None
    if rank == 0:
        per_call = {}
        reps = 3
        for i in range(reps + 1):
            _lib.start_call_timing()
            batch_call(i * FPC)
            torch.cuda.synchronize()
            rec = _lib.stop_call_timing()
            if i == 0:
                continue
            for name, ms in rec:
                per_call.setdefault(name, []).append(ms)
        tot = {k: sum(v) / reps for k, v in per_call.items()}
        cnt = {k: len(v) // reps for k, v in per_call.items()}
        # (two of the 17 conv launches go through uncl_conv3x3_tc_skipcat: the same kernel family with the skip operators fused)
        # ... and the C_out = 32 layers at 124..256 pixels through uncl_conv3x3_tc_rows / _rows_skipcat (conv_tc_rows.cu; a call
        # covers its trailing-column launch of the older kernel as well)
        # uncl_conv_first_conv3x3_tc_rows = inc.conv1 with inc.conv (1 -> 32) computed inside the same launch: its whole time counts
        # as conv time although the first conv's 0.04 GFLOP per tile are not in GFLOP_TILE_TC
        TC_CALLS = ("uncl_conv3x3_tc", "uncl_conv3x3_tc_skipcat", "uncl_conv3x3_tc_rows", "uncl_conv3x3_tc_rows_skipcat",
                    "uncl_conv_first_conv3x3_tc_rows")
        tc_ms = sum(tot.get(k, 0.0) for k in TC_CALLS) if args.precision == "bf16" else tot.get("uncl_conv3x3_simt")
        burst, sus_peak, hbm, how = measured_peaks()
        if tc_ms:
            name = "conv3x3_tc" if args.precision == "bf16" else "conv3x3_simt"
            n_launch = sum(cnt.get(k, 0) for k in TC_CALLS) if args.precision == "bf16" else cnt["uncl_" + name]
            achieved = FPC * TILES * GFLOP_TILE_TC / tc_ms  # GFLOP/ms == TFLOP/s (tot / cnt are per generator call = FPC frames)
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
            if os.path.exists(tpath):   # captured at 60 tiles per launch; a launch of FPC frames moves FPC times the bytes
                traffic = FPC * json.load(open(tpath)).get("dram_bytes_per_launch_avg")
            roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                    "frac": achieved / burst, "frac_of_sustained_peak": achieved / sus_peak, "traffic": traffic,
                    "peak_source": how + ": burst figure (the kernel is timed per launch in a short instrumented pass)",
                    "traffic_source": "profiles/conv_tc_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch)",
                    "launches_per_call": n_launch, "frames_per_call": FPC, "tiles_per_launch": FPC * TILES,
                    "avg_launch_ms": tc_ms / n_launch,
                    "algorithmic_gflop_per_launch_avg": FPC * TILES * GFLOP_TILE_TC / n_launch,
                    "share_of_step": tc_ms / sum(tot.values()),
                    "step_breakdown_ms": {k: round(v, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}}
        # memory-bound kernels of the frame path: ALGORITHMIC bytes per call (SURVEY.md 8d / DESIGN.md section 3) over the
        # device time of the call, against the measured copy bandwidth
        h1, w1 = 16 * (H // 16) + 16, 16 * (W // 16) + 16
        # fused stage calls (one cooperative launch each): normalise_tiles reads rgb twice (statistics, then values) and writes
        # the tiles; blend_percentiles reads the tiles, writes the plane (its keys stay in shared memory); post_u8 reads
        # plane + rgb and writes the 8-bit image (keys in shared memory)
        algo = {"uncl_frame_normalise_tiles": 24 * H * W + TILES * 65536 * 4,
                "uncl_frame_blend_percentiles": TILES * 65536 * 4 + h1 * w1 * 4,
                "uncl_frame_post_u8": h1 * w1 * 4 + 12 * H * W + 3 * H * W,
                "uncl_frame_normalise_pad": 28 * H * W, "uncl_tiles_gather": TILES * 65536 * 8,
                "uncl_tiles_blend": TILES * 65536 * 4 + h1 * w1 * 4, "uncl_frame_postprocess": h1 * w1 * 4 + 24 * H * W,
                "uncl_frame_to_u8": 15 * H * W, "uncl_percentile_pair": 4 * (h1 * w1 + 3 * H * W) // 2,
                "uncl_conv_first": TILES * (256 * 256 * 4 + 254 * 254 * 32 * 2),
                "uncl_maxpool2": int(TILES * 2 * 1.25 * (32 * 252 ** 2 + 64 * 122 ** 2 + 128 * 57 ** 2 + 256 * 24 ** 2) / 4)}
        hbm_rows = []
        for k, nbytes in algo.items():
            if k in tot and cnt.get(k):
                if k in ("uncl_conv_first", "uncl_maxpool2"):
                    nbytes *= FPC          # generator kernels see the tiles of all FPC frames in one call
                us = tot[k] / cnt[k] * 1e3
                hbm_rows.append({"kernel": k[5:], "calls_per_generator_call": cnt[k], "bytes": int(nbytes), "us": round(us, 2),
                                 "frac": round(nbytes / (us * 1e-6) / 1e9 / hbm, 3)})
        roofline_hbm = {"peak_gbs": hbm, "note": "bytes = algorithmic bytes per CALL (percentile_pair / maxpool2: mean over the "
                        "calls of a step); us = device time per call, events around the C-ABI call", "kernels": hbm_rows}

    # BASELINE.json configs[2]: the other image resolutions (HDR-Survey 1/4 resolution, 4K), same path, frames resident in HBM
    other_res = None
    if not args.no_train and world == 1 and rank == 0 and args.precision == "bf16":
        other_res = {}
        for (h2, w2, per_call, reps) in ((768, 1024, 8, 6), (2160, 3840, 1, 6)):
            fr = [torch.from_numpy(synth.hdr_frame(h2, w2, seed=200 + i)).to(dev) for i in range(2)]
            batch = [fr[k % 2] for k in range(per_call)]
            for _ in range(2):
                pipe.tonemap_frames(batch, LAMBDA, uint8=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                pipe.tonemap_frames(batch, LAMBDA, uint8=True)
            e1.record()
            torch.cuda.synchronize()
            pl2 = pipe.plan(h2, w2, dev)
            other_res["%dx%d" % (w2, h2)] = {"frames_per_s": reps * per_call / (e0.elapsed_time(e1) / 1e3), "tiles_per_frame": pl2.ntiles,
                                             "frames_per_generator_call": per_call,
                                             "frame_stages": "fused cooperative kernels" if pipe.fused_ok(pl2) else "staged kernels (order keys exceed shared memory)"}
            del fr, batch
        torch.cuda.empty_cache()
    train = video = None
    fp32_exact = {}
    if not args.no_train and world == 1 and args.precision == "bf16":
        # the exact path on the tensor cores (fp32 activations, three-term bf16 split convolutions, fp32 accumulation:
        # generator rel-L2 4e-7 vs the reference; tests/test_gpu_generator.py): the same frame step
        net32 = UNet(*G_ARGS, up_mode=0, precision="fp32_tc").to(dev).eval()
        net32.load_state_dict(make_generator_state_dict())
        pipe32 = FramePipeline(net32)
        for _ in range(2):
            pipe32.tonemap(dev_frames[0], LAMBDA, uint8=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(4):
            pipe32.tonemap(dev_frames[i % nframes], LAMBDA, uint8=True)
        e1.record()
        torch.cuda.synchronize()
        fp32_exact["frames_per_s"] = 4 / (e0.elapsed_time(e1) / 1e3)
        fp32_exact["inference_path"] = "precision='fp32_tc' (tcgen05, three-term bf16 split)"
        fp32_exact["training_path"] = "precision='fp32_tc' (3x3 convolutions forward / data / weight gradient as three-term bf16 split GEMMs on tcgen05, the rest on fp32 CUDA cores)"
        net32 = pipe32 = None
    if not args.no_train:
        net = pipe = None
        dev_frames = None
        torch.cuda.empty_cache()
        with torch.enable_grad():
            train = train_workload(dev, args.train_precision, max(10, args.steps), args.warmup, world, rank)
            exact = train_workload(dev, "fp32_tc", 5, 3, 1, 0) if world == 1 else None
            if world > 1:
                weak = train_workload(dev, args.train_precision, max(10, args.steps), args.warmup, world, rank, weak=True)
                train["weak_16_images_per_gpu"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "images_per_s", "scaling")}
        if exact is not None:
            train["fp32_exact_path"] = {"value": exact["value"], "unit": "steps/s", "ms_per_step": exact["ms_per_step"]}
            fp32_exact["steps_per_s"] = exact["value"]
            if not args.no_cpu_baseline:
                train["cpu_baseline"] = cpu_train_arm()
        # BASELINE.json configs[3] / configs[4]: the video generator (inference on a 1080p clip, 256^2 training step)
        video = {"inference": video_inference_workload(dev, args.precision, world, rank)}
        torch.cuda.empty_cache()
        with torch.enable_grad():
            video["train"] = train_workload(dev, args.train_precision, 5, args.warmup, world, rank, video=True)

    if rank == 0:
        base = None
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_reference_arm(steps=8, warmup=1)
            base = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        fps = world * args.steps / (ms_res / 1e3)
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": CONFIG,
                "tflops_generator": world * args.steps * TILES * GFLOP_TILE / ms_res,
                "e2e": {"value": world * args.steps / (ms_e2e / 1e3), "unit": "frames/s",
                        "h2d_bytes_per_step": 3 * H * W * 4, "d2h_bytes_per_step": H * W * 3},
                "first_try_ms_per_step": first_try,
                "gpu_launches": launches, "roofline": roof, "roofline_hbm": roofline_hbm, "cpu_baseline": base,
                "clocks": clocks, "sustained": sustained, "fp32_exact": fp32_exact or None, "other_resolutions": other_res,
                "train": train, "video": video}
        # short scalars LAST: a reader that keeps only the tail of the line still sees every headline number
        line["summary"] = {
            "fps": round(fps, 1), "e2e_fps": round(line["e2e"]["value"], 1),
            "sustained_fps": round(sustained["value"], 1) if sustained else None,
            "conv_frac_of_burst_peak": round(roof["frac"], 3) if roof else None,
            "train_steps_per_s": round(train["value"], 1) if train else None,
            "train_weak_steps_per_s": round(train["weak_16_images_per_gpu"]["value"], 1) if train and "weak_16_images_per_gpu" in train else None,
            "video_train_steps_per_s": round(video["train"]["value"], 1) if video else None,
            "video_fps": round(video["inference"]["value"], 1) if video else None,
            "fp32_exact_fps": round(fp32_exact["frames_per_s"], 1) if fp32_exact.get("frames_per_s") else None,
            "fp32_exact_steps_per_s": round(fp32_exact["steps_per_s"], 1) if fp32_exact.get("steps_per_s") else None,
            "n_gpus": world}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
