"""Steps/s of the benchmarked training step (bench.train_workload, CUDA-graph replay) with the weight gradients on a side
stream (train_graph.SIDE_STREAM_WGRAD) and in line."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.argv = ["bench"]
import torch
import bench
from uncltmo_b200 import train_graph
dev = torch.device("cuda:0")
with torch.enable_grad():
    for rep in range(2):
        for flag in (False, True):
            train_graph.SIDE_STREAM_WGRAD = flag
            r = bench.train_workload(dev, "bf16", 20, 5, 1, 0)
            print("side stream %-5s  %.1f steps/s  %.3f ms" % (flag, r["value"], r["ms_per_step"]), flush=True)
    r = bench.train_workload(dev, "bf16", 5, 3, 1, 0, video=True)
    print("video, side stream on: %.1f steps/s" % r["value"])
