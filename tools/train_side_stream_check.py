"""Steps/s of the benchmarked training step (bench.train_workload, CUDA-graph replay) with the weight gradients on a side
stream (train_graph.SIDE_STREAM_WGRAD) and in line."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.argv = ["bench"]
import torch
import bench
from uncltmo_b200 import train_graph, trainer
dev = torch.device("cuda:0")
_init = trainer.GanTrainerStep.__init__
OVERLAP = [True]


def _patched(self, *a, **k):
    _init(self, *a, **k)
    self.overlap_g_forward = OVERLAP[0]


trainer.GanTrainerStep.__init__ = _patched
print("stream priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "?")
with torch.enable_grad():
    for rep in range(3):
        for pm, pg in ((-3, -2),):
            trainer.PRIO_MAIN, trainer.PRIO_G = pm, pg
            r = bench.train_workload(dev, "bf16", 20, 5, 1, 0)
            import gc; gc.collect(); torch.cuda.empty_cache()
            print("mem allocated GB %.1f" % (torch.cuda.memory_allocated() / 2**30), flush=True)
            print("priorities main %d generator %d  %.1f steps/s  %.3f ms" % (pm, pg, r["value"], r["ms_per_step"]), flush=True)
    r = bench.train_workload(dev, "bf16", 5, 3, 1, 0, video=True)
    print("video, side stream on: %.1f steps/s" % r["value"])
