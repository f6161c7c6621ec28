"""torchrun --nproc-per-node 2 tools/clip_shard_check.py : a video scene sharded by tile chain over the ranks gives the
same frames as the single-GPU run (bit-exact: the same kernels see the same tiles)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torch.distributed as dist
from uncltmo_b200 import synth
from uncltmo_b200.frame import FramePipeline
from uncltmo_b200.generator import UNetVideo
from uncltmo_b200.weights import make_generator_state_dict

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
net = UNetVideo(*G_ARGS, up_mode=0, precision="bf16").cuda().eval()
net.load_state_dict(make_generator_state_dict())
pipe = FramePipeline(net)
clip = torch.from_numpy(synth.hdr_clip(3, 540, 700, seed=5)).cuda()
with torch.no_grad():
    single = pipe.tonemap_clip(clip, 371.4, uint8=True)
    sharded = pipe.tonemap_clip(clip, 371.4, uint8=True, shard_tiles=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for flag in (False, True):
        dist.barrier(); torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            pipe.tonemap_clip(clip, 371.4, uint8=True, shard_tiles=flag)
        e1.record(); torch.cuda.synchronize(); times.append(e0.elapsed_time(e1) / 3)
print("rank %d/%d: identical=%s  tiles=%d  single %.2f ms  sharded %.2f ms per 3-frame scene" % (
    rank, world, torch.equal(single, sharded), pipe.plan(540, 700, clip.device).ntiles, times[0], times[1]), flush=True)
dist.destroy_process_group()
