"""Parity gates on the configurations bench.py reports.

(a) the MIXED-PRECISION (bf16 tensor-core) training step on the reference recipe's 16-image batch, epochs 0 / 7 / 10,
    image and video trainer: errD / errG_d / errG_struct within 1e-3 (BASELINE.json's gate) of what the reference's own
    train_D + train_G produced on CPU (tests/golden/reference_train_outputs.npz `ref/`) and of the float64 oracle
    (`o64/`); every generator gradient against the oracle evaluated with the same bf16 rounding points (`obf/`).
(d) the reference's trainer call sequence (tests/ref_style_trainer.py, validated against the reference on the CPU in
    tests/test_oracle_train_golden.py) driving the drop-in netG / netD / StructLoss: INTEGRATION.md section 1.
"""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import oracle
from ref_style_trainer import RefStyleTrainer
from uncltmo_b200 import losses as drop_in_losses
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.generator import UNet, UNetVideo
from uncltmo_b200.struct_loss import StructLoss
from uncltmo_b200.trainer import GanTrainerStep
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
LOSS_TOL = 1e-3   # BASELINE.json north_star: "loss values within 1e-3 relative"


@pytest.fixture(scope="module")
def train_golden():
    return np.load(os.path.join(HERE, "golden", "reference_train_outputs.npz"))


class RecordingSGD(torch.optim.SGD):
    """lr = 0 optimizer that keeps the gradients it was handed."""

    def step(self):
        self.seen = {id(p): p.grad.detach().clone() for g in self.param_groups for p in g["params"] if p.grad is not None}


def _stats(g):
    g = g.detach().double().reshape(-1).cpu()
    step = max(1, g.numel() // 64)
    return np.concatenate([[g.norm().item(), g.sum().item(), g.abs().sum().item()], g[::step][:64].numpy()])


def _nets(video, precision, droppath=None):
    netG = (UNetVideo if video else UNet)(*G_ARGS, up_mode=0, precision=precision).cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netG.drop_path_prob = 0.0
    if droppath is not None:   # train-mode DropPath with the masks of the fixture instead of fresh draws
        masks = [m.cuda() for m in droppath]
        netG._droppath_scale = lambda n, device: masks
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    return netG, netD


# Gradient tolerances.  What separates the CUDA path from the bf16-operand oracle is fp32 (vs float64) accumulation -
# but bf16 arithmetic makes that matter far more than its 1e-7 size suggests: a 1e-7 difference flips the bf16 rounding
# of ~1e-4 of the activations, those flips move pre-activations by ~1e-5, and every ReLU / max-pool / KNN decision that
# flips near a tie switches a gradient spike of the skip operator sqrt(x + 1e-8) (derivative up to 5000) on or off.
# Measured IN THE ORACLE ALONE (float64 accumulation, two bf16 evaluations whose inputs differ by 1e-7 relative,
# 2 images): rel-L2 0.32 on down_path.2.*, 0.02 on down_path.1.*, 0.04-0.08 on the graph block, <= 3e-3 on
# up_path.2/3 and the out conv; the same perturbation moves the float64 gradients by 1e-4.  (bf16-operand vs float64
# gradients: 0.42 / 0.23 / 0.10-0.18 / 0.01.)  A fixed per-tensor bound of 1e-2 is therefore not a property any bf16
# implementation of this network can have; the full-tensor test below calibrates its bound on that intrinsic
# sensitivity instead, and this test holds the gradient NORMS of the 16-image step to the stated bounds.
# (measured at 16 images, worst case over epochs 0 / 7 / 10 and the DropPath case: down_path.2 6.4e-2, down_path.1 1.7e-2,
# every other group <= 1.6e-2)
GRAD_TOL_NORM = {"up_path.2": 2e-2, "up_path.3": 2e-2, "outc": 2e-2, "up_path.1": 3e-2, "up_path.0": 5e-2, "gcn": 5e-2,
                 "down_path.3": 5e-2, "down_path.2": 1.5e-1, "down_path.1": 5e-2, "down_path.0": 5e-2, "inc": 5e-2}


def _norm_tol(k):
    for prefix, tol in GRAD_TOL_NORM.items():
        if k.startswith(prefix):
            return tol
    raise KeyError(k)


@pytest.mark.parametrize("video,epoch,dp", [(False, 0, False), (False, 7, False), (False, 10, False), (False, 0, True),
                                            (True, 0, False), (True, 7, False), (True, 10, False)])
def test_mixed_precision_step_16_images(video, epoch, dp, train_golden):
    tag = "%s/e%d/b16%s/" % ("vid" if video else "img", epoch, "dp" if dp else "")
    hdr, pos, neg = gi.train_batch(16, video)
    netG, netD = _nets(video, "bf16", gi.droppath_masks(16) if dp else None)
    optG = RecordingSGD([p for p in netG.parameters() if p.requires_grad], lr=0.0)
    optD = RecordingSGD(netD.parameters(), lr=0.0)
    tr = GanTrainerStep(netG, netD, optG, optD)
    err_g, err_s = tr.step(hdr.cuda(), None, pos.cuda(), neg.cuda(), epoch)
    torch.cuda.synchronize()
    got = {"errD": tr.errD.item(), "errG_d": err_g.item(), "errG_struct": err_s.item()}
    report = {}
    for anchor in (("o64",) if dp else ("ref", "o64")):
        for k, v in got.items():
            want = float(train_golden[anchor + "/" + tag + k])
            report[anchor + "." + k] = abs(v - want) / abs(want)
    print("mixed step %s: relative loss deviations %s" % (tag, {k: "%.1e" % v for k, v in report.items()}))
    assert max(report.values()) <= LOSS_TOL, report
    # discriminator gradients (fp32 kernels, but D(fake) sees the bf16 generator's output: measured 5e-3): float64 oracle
    big = max(float(train_golden["o64/" + tag + "gD/" + k][0]) for k, _ in netD.named_parameters())
    for k, p in netD.named_parameters():
        want, have = train_golden["o64/" + tag + "gD/" + k], _stats(optD.seen[id(p)])
        # (the bias gradients are sums that cancel to ~3 % of their terms: floor relative to the largest gradient)
        assert np.linalg.norm(have[3:] - want[3:]) <= 1e-2 * np.linalg.norm(want[3:]) + 1e-3 * big, k
    if video:
        return
    # generator gradients against the oracle with the same bf16 rounding points
    dev = {}
    for k, p in netG.named_parameters():
        key = "obf/" + tag + "gG/" + k
        if key not in train_golden.files:
            continue
        want, have = train_golden[key], _stats(optG.seen[id(p)])
        dev[k] = (abs(have[0] - want[0]) / (want[0] + 1e-30),
                  np.linalg.norm(have[3:] - want[3:]) / (np.linalg.norm(want[3:]) + 1e-30))
    worst = sorted(dev.items(), key=lambda kv: -kv[1][0])[:5]
    print("mixed step %s: worst gradient-norm deviations vs bf16-operand oracle %s"
          % (tag, [(k, "%.1e" % v[0], "%.1e" % v[1]) for k, v in worst]))
    bad = {k: v for k, v in dev.items() if v[0] > _norm_tol(k)}
    assert not bad, bad


@pytest.mark.parametrize("video", [False, True])
def test_mixed_precision_gradients_full_tensors(video):
    """2 images (video: one clip of two frames, gradients flow through the recurrent hand-over), every element of every generator gradient: rel-L2 per tensor against the bf16-operand oracle run live
    (float64 accumulation), with the bound CALIBRATED on the oracle itself: a second bf16-operand evaluation whose input
    is perturbed by 1e-6 relative (the size of fp32 accumulation error) gives, per tensor, the distance between two
    equally valid bf16 evaluations of the same step; the CUDA path must be within 3x that distance (floor 2e-2)."""
    hdr, pos, neg = gi.train_batch(2, video)
    g_sd = {k: v.double() for k, v in make_generator_state_dict().items()}
    d_sd = {k: v.double() for k, v in make_discriminator_state_dict().items()}
    h = hdr.double() if video else hdr[0].double()
    with oracle.bf16_operands(True):
        ref = oracle.train_step_losses(g_sd, d_sd, h, pos[0].double(), neg[0].double(), 0)
        gen = torch.Generator().manual_seed(0)
        h2 = h * (1 + 1e-6 * torch.randn(h.shape, generator=gen, dtype=torch.float64))
        ref2 = oracle.train_step_losses(g_sd, d_sd, h2, pos[0].double(), neg[0].double(), 0)
    netG, netD = _nets(video, "bf16")
    optG = RecordingSGD([p for p in netG.parameters() if p.requires_grad], lr=0.0)
    tr = GanTrainerStep(netG, netD, optG, RecordingSGD(netD.parameters(), lr=0.0))
    assert tr.video == video
    err_g, err_s = tr.step(hdr.cuda(), None, pos.cuda(), neg.cuda(), 0)
    for name, v, w in (("errD", tr.errD.item(), ref["errD"]), ("errG_d", err_g.item(), ref["errG_d"]),
                       ("errG_struct", err_s.item(), ref["errG_struct"])):
        assert abs(v - w) <= 2e-4 * abs(w), (name, v, w)   # same rounding points: tighter than the 1e-3 gate
    rels, noise = {}, {}
    for k, p in netG.named_parameters():
        if k in ref["grads_G"]:
            a, b = optG.seen[id(p)].double().cpu(), ref["grads_G"][k]
            rels[k] = ((a - b).norm() / (b.norm() + 1e-30)).item()
            noise[k] = ((ref2["grads_G"][k] - b).norm() / (b.norm() + 1e-30)).item()
    print("mixed gradients (%s) vs bf16-operand oracle (rel-L2, oracle's own bf16 sensitivity):" % ("video" if video else "image"),
          [(k, "%.1e" % rels[k], "%.1e" % noise[k]) for k in sorted(rels, key=lambda k: -rels[k])[:10]])
    bad = {k: (v, noise[k]) for k, v in rels.items() if v > max(2e-2, 3.0 * noise[k])}
    assert not bad, bad


@pytest.mark.parametrize("video,epoch,precision,swap_losses", [(False, 0, "fp32", False), (False, 7, "bf16", False),
                                                               (False, 10, "bf16", True), (True, 0, "bf16", True),
                                                               (True, 7, "fp32", False)])
def test_reference_call_sequence_over_drop_in_modules(video, epoch, precision, swap_losses, train_golden):
    """The reference trainer's train_D / train_G call sequence (two backward passes with retain_graph, torch loss glue,
    host TMQI on .cpu().numpy() copies) over the drop-in netG / netD / StructLoss; optionally the loss methods swapped
    for uncltmo_b200.losses as well.  Values against the reference trainer's own (4 images)."""
    tag = "ref/%s/e%d/b4/" % ("vid" if video else "img", epoch)
    hdr, pos, neg = (t.cuda() for t in gi.train_batch(4, video))
    netG, netD = _nets(video, precision)
    sl = StructLoss(torch.tensor([1.0, 1.0, 1.0]))
    optG = RecordingSGD([p for p in netG.parameters() if p.requires_grad], lr=0.0)
    t = RefStyleTrainer(netG, netD, sl, optG, RecordingSGD(netD.parameters(), lr=0.0), video=video,
                        losses=drop_in_losses if swap_losses else None)
    t.adv_weight_list = t.adv_weight_list.cuda()
    t.pyramid_weight_list = t.pyramid_weight_list.cuda()
    t.train_D(hdr, pos, neg, epoch)
    t.train_G(hdr, hdr, pos, neg, epoch)
    for k in ("errD", "errG_d", "errG_struct"):
        want = float(train_golden[tag + k])
        assert abs(getattr(t, k).item() - want) <= LOSS_TOL * abs(want), (k, getattr(t, k).item(), want)
    # both backward calls reached the generator: decoder gradient norms agree with the reference's
    for k, p in netG.named_parameters():
        if k.startswith("up_path.3") and (tag + "gG/" + k) in train_golden.files:
            want = train_golden[tag + "gG/" + k][0]
            assert abs(optG.seen[id(p)].double().norm().item() - want) <= (3e-2 if precision == "bf16" else 2e-3) * want, k
