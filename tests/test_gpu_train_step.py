"""One GAN training step (train_D + train_G of GanTrainerImg) on the B200 path against the CPU oracle's step:
loss values within 1e-3 relative (BASELINE.json), gradients of the well-conditioned parameter groups within 1e-3."""
import numpy as np
import pytest
import torch

import oracle
from uncltmo_b200 import synth
from uncltmo_b200.discriminator import SimpleDiscriminator
from uncltmo_b200.generator import UNet
from uncltmo_b200.trainer import GanTrainerStep
from uncltmo_b200.weights import make_discriminator_state_dict, make_generator_state_dict

pytestmark = pytest.mark.gpu
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


class RecordingSGD(torch.optim.SGD):
    """lr = 0 optimizer that keeps a copy of the gradients it was handed (zero_grad follows immediately in the trainer)."""

    def step(self):
        self.seen = {id(p): p.grad.detach().clone() for g in self.param_groups for p in g["params"] if p.grad is not None}


@pytest.mark.parametrize("epoch,precision", [(0, "fp32"), (7, "fp32"), (10, "fp32"), (0, "fp32_tc"), (10, "fp32_tc")])
def test_train_step_matches_oracle(epoch, precision):
    """precision 'fp32': CUDA-core kernels; 'fp32_tc': the exact path on the tensor cores (3x3 convolutions - forward, data
    and weight gradients - as three-term bf16 split GEMMs, ~2^-16 per product): the same gates on losses and D gradients."""
    g_sd, d_sd = make_generator_state_dict(), make_discriminator_state_dict()
    hdr = torch.from_numpy(synth.normalised_batch(2, seed=4)).reshape(1, 2, 1, 256, 256)
    pos = torch.from_numpy(synth.ldr_batch(2, seed=5)).reshape(1, 2, 1, 256, 256)
    neg = torch.from_numpy(synth.ldr_batch(2, seed=6)).reshape(1, 2, 1, 256, 256)
    # float64 oracle: several gradients (e.g. the out-conv bias = a signed sum over 131072 pixels) cancel heavily, and an
    # fp32 reference would carry as much rounding noise as the path under test
    ref = oracle.train_step_losses({k: v.double() for k, v in g_sd.items()}, {k: v.double() for k, v in d_sd.items()},
                                   hdr[0].double(), pos[0].double(), neg[0].double(), epoch)

    netG = UNet(*G_ARGS, up_mode=0, precision=precision).cuda().train()
    netG.load_state_dict(g_sd)
    netG.drop_path_prob = 0.0
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(d_sd)
    optG = RecordingSGD([p for p in netG.parameters() if p.requires_grad], lr=0.0)
    optD = RecordingSGD(netD.parameters(), lr=0.0)
    tr = GanTrainerStep(netG, netD, optG, optD)
    err_g, err_s = tr.step(hdr.cuda(), None, pos.cuda(), neg.cuda(), epoch)
    assert abs(tr.errD.item() - ref["errD"]) <= 1e-3 * abs(ref["errD"])
    assert abs(err_g.item() - ref["errG_d"]) <= 1e-3 * abs(ref["errG_d"])
    assert abs(err_s.item() - ref["errG_struct"]) <= 1e-3 * abs(ref["errG_struct"])
    biggest = max(g.double().norm().item() for g in ref["grads_D"].values())
    for k, p in netD.named_parameters():
        # model.4.bias shifts every logit alike and the contrastive loss only sees differences: its exact gradient
        # is 0 and both sides hold rounding noise there, hence the floor relative to the largest gradient
        got, want = optD.seen[id(p)].double().cpu(), ref["grads_D"][k].double()
        assert (got - want).norm().item() <= 1e-3 * want.norm().item() + 1e-6 * biggest, k
    # decoder / graph block / deepest encoder stage are well conditioned (tests/test_gpu_backward.py explains why the
    # shallow encoder stages are not, in the reference itself)
    bad = {}
    for k, p in netG.named_parameters():
        if k in ref["grads_G"] and not (k.startswith("inc.") or k[:11] in ("down_path.0", "down_path.1", "down_path.2")):
            e = rel(optG.seen[id(p)], ref["grads_G"][k])
            # outc.conv.bias = sum over all pixels of d(logit): the struct-loss part of that sum cancels to ~0 window by
            # window while its terms are ~1e4 larger, so fp32 leaves ~1e-2 of noise on the small remainder
            tol = 5e-2 if k == "outc.conv.bias" else 2e-3
            # fp32_tc: the features that reach the graph block carry ~2^-16 instead of ~2^-23, which flips a handful of
            # near-tie KNN choices (discrete).  The block's output, and the gradient that flows back through it, then differ
            # by a fraction of a per cent in the layers next to it (measured 2e-3 ... 7e-3: deepest encoder stage, the block,
            # the first two decoder stages); everything further away stays at the 2e-3 gate
            if precision == "fp32_tc" and k.startswith(("down_path.3", "gcn.", "up_path.0", "up_path.1")):
                tol = 2e-2
            if e > tol:
                bad[k] = e
    assert not bad, "\n".join("%s %.3e" % kv for kv in sorted(bad.items()))


def test_adam_training_reduces_struct_loss():
    """A few real optimizer steps (Adam, the reference's lr 1e-5 x100 to see movement): losses stay finite, G changes."""
    torch.manual_seed(0)
    netG = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-3, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999))
    tr = GanTrainerStep(netG, netD, optG, optD)
    hdr = torch.from_numpy(synth.normalised_batch(4, seed=4)).reshape(2, 2, 1, 256, 256).cuda()
    pos = torch.from_numpy(synth.ldr_batch(4, seed=5)).reshape(2, 2, 1, 256, 256).cuda()
    neg = torch.from_numpy(synth.ldr_batch(4, seed=6)).reshape(2, 2, 1, 256, 256).cuda()
    hist = []
    for _ in range(6):
        _, s = tr.step(hdr, None, pos, neg, 0)
        hist.append(s.item())
    assert all(np.isfinite(hist)) and hist[-1] < hist[0]


@pytest.mark.parametrize("epoch", [0, 10])
def test_video_train_step_matches_oracle(epoch):
    """GanTrainer (video): 5-D clips through the recurrent generator, features [B*T,64,1,1] into infoNCE2."""
    from uncltmo_b200.generator import UNetVideo
    g_sd, d_sd = make_generator_state_dict(), make_discriminator_state_dict()
    hdr = torch.from_numpy(synth.normalised_batch(2, seed=4)).reshape(1, 2, 1, 256, 256)
    pos = torch.from_numpy(synth.ldr_batch(2, seed=5)).reshape(1, 2, 1, 256, 256)
    neg = torch.from_numpy(synth.ldr_batch(2, seed=6)).reshape(1, 2, 1, 256, 256)
    ref = oracle.train_step_losses({k: v.double() for k, v in g_sd.items()}, {k: v.double() for k, v in d_sd.items()},
                                   hdr.double(), pos[0].double(), neg[0].double(), epoch)
    netG = UNetVideo(*G_ARGS, up_mode=0, precision="fp32").cuda().train()
    netG.load_state_dict(g_sd)
    netG.drop_path_prob = 0.0
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(d_sd)
    optG = RecordingSGD([p for p in netG.parameters() if p.requires_grad], lr=0.0)
    optD = RecordingSGD(netD.parameters(), lr=0.0)
    tr = GanTrainerStep(netG, netD, optG, optD)
    assert tr.video
    err_g, err_s = tr.step(hdr.cuda(), None, pos.cuda(), neg.cuda(), epoch)
    assert abs(tr.errD.item() - ref["errD"]) <= 1e-3 * abs(ref["errD"])
    assert abs(err_g.item() - ref["errG_d"]) <= 1e-3 * abs(ref["errG_d"])
    assert abs(err_s.item() - ref["errG_struct"]) <= 1e-3 * abs(ref["errG_struct"])
    bad = {}
    for k, p in netG.named_parameters():
        if k in ref["grads_G"] and not (k.startswith("inc.") or k[:11] in ("down_path.0", "down_path.1", "down_path.2")):
            e = rel(optG.seen[id(p)], ref["grads_G"][k])
            if e > (5e-2 if k == "outc.conv.bias" else 2e-3):
                bad[k] = e
    assert not bad, bad


def test_graph_replay_matches_eager_steps():
    """GanTrainerStep.capture / replay: the whole iteration as one CUDA graph gives the same training trajectory as the
    eager launches (atomics make the sums order-dependent, hence a tolerance instead of bit equality)."""
    hdr = torch.from_numpy(synth.normalised_batch(4, seed=4)).reshape(2, 2, 1, 256, 256).cuda()
    pos = torch.from_numpy(synth.ldr_batch(4, seed=5)).reshape(2, 2, 1, 256, 256).cuda()
    neg = torch.from_numpy(synth.ldr_batch(4, seed=6)).reshape(2, 2, 1, 256, 256).cuda()
    hdr2 = torch.from_numpy(synth.normalised_batch(4, seed=14)).reshape(2, 2, 1, 256, 256).cuda()

    def make():
        netG = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train()
        netG.load_state_dict(make_generator_state_dict())
        netG.drop_path_prob = 0.0
        netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
        netD.load_state_dict(make_discriminator_state_dict())
        optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-4, betas=(0.5, 0.999), capturable=True)
        optD = torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=True)
        return netG, netD, GanTrainerStep(netG, netD, optG, optD)

    gA, dA, trA = make()
    for b in (hdr, hdr, hdr2, hdr):
        eg, es = trA.step(b, None, pos, neg, 0)
    gB, dB, trB = make()
    trB.capture(hdr, None, pos, neg, 0, warmup=2)      # two real iterations on `hdr`
    trB.replay(hdr2, None, pos, neg, 0)
    rg, rs = trB.replay(hdr, None, pos, neg, 0)
    torch.cuda.synchronize()
    assert abs(rg.item() - eg.item()) <= 2e-3 * abs(eg.item()) and abs(rs.item() - es.item()) <= 2e-3 * abs(es.item())
    assert abs(trB.errD.item() - trA.errD.item()) <= 1e-2 * abs(trA.errD.item()), (trB.errD.item(), trA.errD.item())
    # (parameters are not compared element-wise: Adam normalises every element's step, so elements whose gradient is
    # rounding noise move differently from run to run - two EAGER runs differ by 35 % of the movement of the noisiest
    # tensor after four steps, exactly as eager vs replay does; tools/graph_check.py prints both)
    # the inference path sees the replayed parameters (the packing cache is keyed on version counters replay bypasses)
    with torch.no_grad():
        gA.eval(), gB.eval()
        assert rel(gB(hdr[0])[0], gA(hdr[0])[0]) < 2e-2
    with pytest.raises(ValueError):
        GanTrainerStep(gA, dA, torch.optim.Adam(gA.parameters()), torch.optim.Adam(dA.parameters())).capture(hdr, None, pos, neg, 0)


def test_capture_after_eager_steps_on_the_default_stream():
    """The same trainer first steps eagerly on the legacy default stream and is then captured (what bench.py does): the
    nested graph-block autograd of the bf16 path must not tie parameter bookkeeping to the legacy stream."""
    from uncltmo_b200.optim import FlatAdam
    hdr = torch.from_numpy(synth.normalised_batch(4, seed=4)).reshape(2, 2, 1, 256, 256).cuda()
    pos = torch.from_numpy(synth.ldr_batch(4, seed=5)).reshape(2, 2, 1, 256, 256).cuda()
    neg = torch.from_numpy(synth.ldr_batch(4, seed=6)).reshape(2, 2, 1, 256, 256).cuda()
    netG = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    tr = GanTrainerStep(netG, netD, FlatAdam(netG, lr=1e-4, betas=(0.5, 0.999)),
                        torch.optim.Adam(netD.parameters(), lr=1e-4, betas=(0.5, 0.999), capturable=True, fused=True))
    for _ in range(2):
        eg, es = tr.step(hdr, None, pos, neg, 0)
    torch.cuda.synchronize()
    first = es.item()
    tr.capture(hdr, None, pos, neg, 0, warmup=1)
    for _ in range(3):
        rg, rs = tr.replay(hdr, None, pos, neg, 0)
    torch.cuda.synchronize()
    assert np.isfinite(rs.item()) and rs.item() < first      # training continued through the replays


@pytest.mark.parametrize("video,precision", [(False, "bf16"), (True, "bf16"), (False, "fp32_tc")])
def test_eager_steps_do_not_accumulate_device_memory(video, precision):
    """Regression: the bf16 generator's single autograd node kept its saved state on `ctx` together with the very tensor
    objects it returned (ctx.S -> out -> grad_fn -> node -> ctx, a cycle through C++ the garbage collector cannot see), so
    every eager step left its activations - 0.8 GB at 16 images - allocated for ever.  Steps must be memory-neutral."""
    import gc
    from uncltmo_b200.generator import UNetVideo
    netG = (UNetVideo if video else UNet)(*G_ARGS, up_mode=0, precision=precision).cuda().train()
    netG.load_state_dict(make_generator_state_dict())
    netD = SimpleDiscriminator(256, 1, 16, "none", "none", 0, 0).cuda().train()
    netD.load_state_dict(make_discriminator_state_dict())
    optG = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=1e-5, betas=(0.5, 0.999))
    optD = torch.optim.Adam(netD.parameters(), lr=1.5e-5, betas=(0.5, 0.999))
    tr = GanTrainerStep(netG, netD, optG, optD)
    hdr = torch.from_numpy(synth.normalised_batch(4, seed=4)).reshape(2, 2, 1, 256, 256).cuda()
    pos = torch.from_numpy(synth.ldr_batch(4, seed=5)).reshape(2, 2, 1, 256, 256).cuda()
    neg = torch.from_numpy(synth.ldr_batch(4, seed=6)).reshape(2, 2, 1, 256, 256).cuda()
    held = []
    for i in range(5):
        tr.step(hdr, None, pos, neg, 0)
        torch.cuda.synchronize()
        gc.collect()
        held.append(torch.cuda.memory_allocated())
    assert held[4] - held[2] <= 8 << 20, [h >> 20 for h in held]     # (Adam state appears in the first steps)
    # the drop-in surface (parameter gradients handed to autograd) as well
    x = hdr.reshape(-1, 2, 1, 256, 256) if video else hdr.reshape(-1, 1, 256, 256)
    held = []
    for i in range(4):
        netG.zero_grad(set_to_none=True)
        out, fea = netG(x)
        (out.mean() + fea.float().mean()).backward()
        del out, fea
        torch.cuda.synchronize()
        gc.collect()
        held.append(torch.cuda.memory_allocated())
    assert held[3] - held[1] <= 8 << 20, [h >> 20 for h in held]
