"""Adam on the generator's flat parameter buffer: one kernel launch per step.

`torch.optim.Adam` semantics (GanTrainerImg.py builds `optim.Adam(netG.parameters(), lr, betas=(0.5, 0.999))`; no weight
decay, no amsgrad), applied to the whole flat fp32 buffer `uncltmo_b200.train_graph.FlatParams` keeps the parameters in
(`uncl_adam_flat`).  Padding elements and frozen parameters (`gcn.module.0.0.relative_pos`) have zero gradient, hence zero
moments and a zero update.  The step counter lives on the device, so the optimizer is CUDA-graph capturable.
"""
import torch

from ._lib import call
from .train_graph import flat_params


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.net = net
        self.fp = flat_params(net)
        params = [p for p in net.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, capturable=True))
        dev = self.fp.flat.device
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.float32)

    @torch.no_grad()
    def step(self, closure=None):
        fp = flat_params(self.net)
        if fp is not self.fp:
            raise RuntimeError("FlatAdam: the network's parameters were moved out of the flat buffer it was built on")
        g = self.param_groups[0]
        if not fp.grads_attached():
            return None      # nothing was back-propagated since zero_grad(set_to_none=True)
        call("uncl_adam_flat", fp.flat, fp.grad, self.m, self.v, fp.total, float(g["lr"]), float(g["betas"][0]),
             float(g["betas"][1]), float(g["eps"]), self.step_count)
        fp._packed_version = None    # in-place update behind autograd's version counters: invalidate the packed operands
        self.net._packed = None
        return None

    def zero_grad(self, set_to_none=True):
        """Keeps the p.grad views attached and clears the flat buffer with one memset (set_to_none is accepted for
        signature compatibility; detaching 60 views only to re-attach them next step would buy nothing)."""
        self.fp.grad.zero_()
        self.fp.attach_grads()
