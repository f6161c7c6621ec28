"""GPU parity of the backward kernels: gradients through the CUDA generator (autograd Functions over the C ABI)
against torch autograd through the CPU oracle on the same weights and inputs.
Gate: gradient rel-L2 <= 1e-3 per parameter tensor on the fp32 path."""
import numpy as np
import pytest
import torch

import golden_inputs as gi
import oracle
from uncltmo_b200.generator import UNet
from uncltmo_b200.weights import make_generator_state_dict

pytestmark = pytest.mark.gpu
G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def oracle_grads(sd, x, proj_out, proj_feat, scale=None):
    params = {k: v.clone().requires_grad_(k != "gcn.module.0.0.relative_pos") for k, v in sd.items()}
    out, feats = oracle.unet_forward(params, x, droppath_masks=scale)
    loss = (out * proj_out).sum() + (feats * proj_feat).sum()
    loss.backward()
    return loss.item(), {k: v.grad for k, v in params.items() if v.grad is not None}


def _run_cuda(sd, x, proj_out, proj_feat, scale):
    net = UNet(*G_ARGS, up_mode=0, precision="fp32").cuda().train()
    net.load_state_dict(sd)
    net.drop_path_prob = 0.0
    if scale is not None:
        out, feats = net._forward_train(x.cuda(), [s.cuda() for s in scale])
    else:
        out, feats = net(x.cuda())
    loss = (out * proj_out.cuda()).sum() + (feats * proj_feat.cuda()).sum()
    loss.backward()
    return loss.item(), {k: p.grad for k, p in net.named_parameters()}


def _problem():
    sd = make_generator_state_dict()
    x = gi.generator_input()[:1]
    rng = np.random.default_rng(3)
    proj_out = torch.from_numpy(rng.standard_normal((1, 1, 256, 256)).astype(np.float32))
    proj_feat = torch.from_numpy((rng.standard_normal((1, 32, 256, 256)) * 0.01).astype(np.float32))
    return sd, x, proj_out, proj_feat


@pytest.mark.parametrize("with_droppath", [False, True])
def test_generator_gradients_match_oracle_well_conditioned(with_droppath):
    """Whole-network gradient parity at the 1e-3 gate.  The skip operator sqrt(x2 + 1e-8) has derivative up to 5000 at
    post-ReLU zeros (SURVEY.md 'hard parts'): pre-activations within rounding of 0 flip that term on or off, so the
    reference's OWN fp32 and fp64 gradients differ by up to 1e-2 in the encoder.  Zeroing the decoder weights that read
    the sqrt channels removes the ill-conditioning while every kernel still runs; the unmodified network is checked in
    the next test."""
    sd, x, proj_out, proj_feat = _problem()
    for i in range(4):
        w = sd["up_path.%d.conv.conv.weight" % i]
        w[3 * w.shape[0] // 4:] = 0.0
    scale = [torch.tensor([1 / 0.95]), torch.tensor([1 / 0.95])] if with_droppath else None
    ref_loss, ref = oracle_grads({k: v.double() for k, v in sd.items()}, x.double(), proj_out.double(), proj_feat.double(),
                                 [s.double() for s in scale] if scale else None)
    loss, got = _run_cuda(sd, x, proj_out, proj_feat, scale)
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    worst = {k: rel(got[k], g) for k, g in ref.items() if not k.endswith("conv.conv.weight") or "up_path" not in k}
    for i in range(4):  # the zeroed slice still receives a (well defined) gradient; compare the live three quarters
        k = "up_path.%d.conv.conv.weight" % i
        q = 3 * ref[k].shape[0] // 4
        worst[k] = rel(got[k][:q], ref[k][:q])
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    assert not bad, bad
    assert max(worst.values()) < 2e-4, max(worst.values())
    assert got["gcn.module.0.0.relative_pos"] is None


def test_generator_gradients_match_oracle_shipped_network():
    """Unmodified network against the float64 oracle: decoder / graph block / deepest encoder stage at 1e-4; the
    encoder stages behind a sqrt skip are ill-conditioned in the reference itself (see above), so they are held to
    the spread between the reference's fp32 and fp64 runs."""
    sd, x, proj_out, proj_feat = _problem()
    _, r64 = oracle_grads({k: v.double() for k, v in sd.items()}, x.double(), proj_out.double(), proj_feat.double())
    _, r32 = oracle_grads(sd, x, proj_out, proj_feat)
    _, got = _run_cuda(sd, x, proj_out, proj_feat, None)
    for k, g in r64.items():
        e_cuda, e_ref = rel(got[k], g), rel(r32[k], g)
        shallow = k.startswith("inc.") or any(k.startswith("down_path.%d" % i) for i in range(3))
        if shallow:
            assert e_cuda <= max(10 * e_ref, 5e-2), (k, e_cuda, e_ref)
        else:
            assert e_cuda <= 1e-4, (k, e_cuda)


def test_mixed_precision_training_path():
    """precision='bf16' under autograd: tensor-core conv forward / data gradient with bf16-rounded operands.  Checked on
    the well-conditioned network (see above) against the float64 oracle at the bf16 tolerance."""
    sd, x, proj_out, proj_feat = _problem()
    for i in range(4):
        w = sd["up_path.%d.conv.conv.weight" % i]
        w[3 * w.shape[0] // 4:] = 0.0
    ref_loss, ref = oracle_grads({k: v.double() for k, v in sd.items()}, x.double(), proj_out.double(), proj_feat.double())
    net = UNet(*G_ARGS, up_mode=0, precision="bf16").cuda().train()
    net.load_state_dict(sd)
    net.drop_path_prob = 0.0
    out, feats = net(x.cuda())
    loss = (out * proj_out.cuda()).sum() + (feats * proj_feat.cuda()).sum()
    loss.backward()
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss)
    # bf16 operand rounding (2^-9) times the conditioning of this signed random-projection loss (~25x, the same factor
    # that takes the fp32 path from 6e-8 to ~1e-6) gives 1e-2 ... 0.17 from the last to the deepest layer
    worst, cos = {}, {}
    for k, p in net.named_parameters():
        if k in ref:
            g, r = p.grad, ref[k]
            if "up_path" in k and k.endswith("conv.conv.weight"):
                q = 3 * r.shape[0] // 4
                g, r = g[:q], r[:q]
            worst[k] = rel(g, r)
            cos[k] = torch.nn.functional.cosine_similarity(g.detach().double().cpu().flatten(), r.double().flatten(), dim=0).item()
    assert max(worst.values()) <= 0.3, worst
    assert min(cos.values()) >= 0.97, cos
    assert worst["outc.conv.weight"] <= 1e-2 and worst["up_path.3.conv.conv1.weight"] <= 5e-2


def test_video_generator_gradients_match_oracle():
    """Recurrent (video) generator, T=2, loss on the LAST frame only: the hand-over slices are not detached in the
    reference (Unet.py:244, 270), so frame 0 is reached through them alone.

    Tolerance.  A ReLU whose pre-activation lies within fp32 rounding of 0 takes a different side in the fp32 CUDA
    path and the float64 oracle (about one element per pass of ~3e7); that single decision changes the gradient of a
    3x3xC patch, i.e. ~1/sqrt(N) = 7e-4 relative on a 126^2 plane and more on the 12^2 ones, and everything upstream
    inherits it (torch's own fp32 run shows the same against float64).  The image test above happens to draw no such
    element; the clip used here does.  So the gate here is 2e-2 + cosine, and the recurrent part is pinned separately:
    it must be reproduced to a fraction of its size, measured against the oracle with the hand-over detached."""
    from uncltmo_b200.generator import UNetVideo
    sd, _, _, _ = _problem()
    for i in range(4):
        w = sd["up_path.%d.conv.conv.weight" % i]
        w[3 * w.shape[0] // 4:] = 0.0
    x = gi.video_input()[:1, :2].contiguous()
    rng = np.random.default_rng(5)
    proj_out = torch.from_numpy(rng.standard_normal((1, 2, 1, 256, 256)).astype(np.float32))
    proj_feat = torch.from_numpy(rng.standard_normal((1, 2, 64, 1, 1)).astype(np.float32))
    proj_out[:, 0] = 0
    proj_feat[:, 0] = 0

    def oracle_run(detach):
        params = {k: v.double().clone().requires_grad_(k != "gcn.module.0.0.relative_pos") for k, v in sd.items()}
        out, feats = oracle.unet_video_forward(params, x.double(), detach_state=detach)
        loss = (out * proj_out.double()).sum() + (feats * proj_feat.double()).sum()
        loss.backward()
        return out, feats, loss.item(), {k: v.grad for k, v in params.items() if v.grad is not None}

    out, feats, ref_loss, ref = oracle_run(False)
    _, _, _, cut = oracle_run(True)

    net = UNetVideo(*G_ARGS, up_mode=0, precision="fp32").cuda().train()
    net.load_state_dict(sd)
    net.drop_path_prob = 0.0
    with torch.enable_grad():
        o, f = net(x.cuda())
        assert o.shape == (1, 2, 1, 256, 256) and f.shape == (1, 2, 64, 1, 1)
        loss = (o * proj_out.cuda()).sum() + (f * proj_feat.cuda()).sum()
        loss.backward()
    assert rel(o, out) < 1e-5 and rel(f, feats) < 1e-4
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    worst, recurrent = {}, {}
    for k, p in net.named_parameters():
        if k not in ref:
            continue
        g, r, c = p.grad.detach().double().cpu(), ref[k], cut[k]
        if "up_path" in k and k.endswith("conv.conv.weight"):
            q = 3 * r.shape[0] // 4
            g, r, c = g[:q], r[:q], c[:q]
        worst[k] = rel(g, r)
        assert torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item() > 0.9995, k
        if rel(c, r) > 2e-2:     # tensors where the recurrence carries a measurable share of the gradient
            recurrent[k] = rel(g, r) / rel(c, r)
    assert max(worst.values()) < 2e-2, {k: v for k, v in worst.items() if v > 2e-2}
    assert len(recurrent) >= 4, recurrent
    assert max(recurrent.values()) < 0.25, recurrent
