"""CPU tests of the host-side mirror: tiling tables, weight packing, module surface."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_inputs as gi
import oracle
from uncltmo_b200 import frame, packing
from uncltmo_b200.generator import UNet, UNetVideo
from uncltmo_b200.weights import generator_shapes, make_generator_state_dict

G_ARGS = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)


@pytest.mark.parametrize("length", [272, 304, 464, 656, 784, 1040, 1088, 1936, 2176, 3856])
def test_tile_starts_match_oracle(length):
    starts, last = oracle.tile_grid(length)
    assert frame.tile_starts(length) == starts + [last]


def test_tile_counts_of_survey():
    # 272^2 -> 2x2; 1040x784 -> 6x4; 1936x1088 -> 10x6; 3856x2176 -> 20x11 / 58x31 (overlap 192)
    assert (len(frame.tile_starts(272)), len(frame.tile_starts(1040)), len(frame.tile_starts(784))) == (2, 6, 4)
    assert (len(frame.tile_starts(1936)), len(frame.tile_starts(1088))) == (10, 6)
    assert (len(frame.tile_starts(3856)), len(frame.tile_starts(2176))) == (20, 11)
    assert (len(frame.tile_starts(3856, 192)), len(frame.tile_starts(2176, 192))) == (58, 31)
    assert frame.padded_extent(1080) == 1088 and frame.padded_extent(1920) == 1936 and frame.padded_extent(256) == 272


def _blend_with_tables(x, model_fn, overlap=64):
    h, w = x.shape[-2:]
    ys, yidx, yw = frame.axis_blend_table(h, overlap)
    xs, xidx, xw = frame.axis_blend_table(w, overlap)
    tiles = {(a, b): model_fn(x[..., ys[a]:ys[a] + 256, xs[b]:xs[b] + 256]).numpy() for a in range(len(ys)) for b in range(len(xs))}
    out = np.zeros(x.shape, dtype=np.float64)
    for yy in range(h):
        for ka in range(yidx.shape[1]):
            if yw[yy, ka] == 0:
                continue
            a = yidx[yy, ka]
            row = np.zeros(x.shape[:-2] + (w,), dtype=np.float64)
            for b in range(len(xs)):
                wcol = np.where(xidx == b, xw, 0).sum(axis=1)[xs[b]:xs[b] + 256]
                row[..., xs[b]:xs[b] + 256] += wcol * tiles[(a, b)][..., yy - ys[a], :]
            out[..., yy, :] += yw[yy, ka] * row
    return out


@pytest.mark.parametrize("shape", [(272, 304), (464, 656), (272, 400)])
def test_closed_form_blend_equals_sequential_crossfade(shape):
    x = torch.from_numpy(gi.blend_field(*shape))
    ref = oracle.tile_and_blend(x, gi.cheap_model).numpy()
    got = _blend_with_tables(x, gi.cheap_model)
    assert np.abs(got - ref).max() < 2e-6


def test_blend_table_full_res_mode():
    starts, idx, w = frame.axis_blend_table(656, overlap=192)
    assert idx.shape[1] <= 6 and np.allclose(w.sum(axis=1), 1.0, atol=1e-6)
    x = torch.from_numpy(gi.blend_field(464, 464))
    ref = oracle.tile_and_blend(x, gi.cheap_model, overlap=192).numpy()
    assert np.abs(_blend_with_tables(x, gi.cheap_model, overlap=192) - ref).max() < 2e-6


def test_blend_weights_partition_unity():
    for length in (272, 1088, 1936):
        _, _, w = frame.axis_blend_table(length)
        assert np.allclose(w.sum(axis=1), 1.0, atol=1e-6)


def test_tiling_rejects_small_frames():
    with pytest.raises(ValueError):
        frame.tile_starts(256)


def test_conv_transpose_packing_is_flipped_correlation():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 16, 9, 9, generator=g)
    wt = torch.randn(16, 8, 3, 3, generator=g)
    ref = F.conv_transpose2d(x, wt)
    w9 = packing.conv3x3_taps(wt, transposed=True)  # [9][Cin][Cout]
    wc = w9.reshape(3, 3, 16, 8).permute(3, 2, 0, 1)
    assert torch.allclose(F.conv2d(F.pad(x, (2, 2, 2, 2)), wc), ref, atol=1e-5)
    wconv = torch.randn(8, 16, 3, 3, generator=g)
    w9 = packing.conv3x3_taps(wconv, transposed=False)
    assert torch.equal(w9.reshape(3, 3, 16, 8).permute(3, 2, 0, 1), wconv)


def test_tc_weight_packing_layout():
    w9 = torch.arange(9 * 32 * 256, dtype=torch.float32).reshape(9, 32, 256) % 251
    p = packing.conv3x3_tc(w9)
    assert p.shape == (2, 2, 9, 2, 128, 8) and p.dtype == torch.bfloat16
    ns, ch, tap, half, n, k = 1, 1, 4, 1, 77, 5
    assert p[ns, ch, tap, half, n, k].float() == w9[tap, ch * 16 + half * 8 + k, ns * 128 + n]


def test_tc_weight_packing_layout_merged_kx():
    """C_out <= 64, C_in >= 64: the three kx taps of a filter row sit side by side in N (conv_tc_merged.cu)."""
    for co in (32, 64):
        w9 = torch.arange(9 * 64 * co, dtype=torch.float32).reshape(9, 64, co) % 251
        p = packing.conv3x3_tc(w9)
        assert p.shape == (1, 4, 3, 2, 3 * co, 8) and p.dtype == torch.bfloat16
        ch, ky, half, kx, n, k = 2, 1, 1, 2, co - 3, 5
        assert p[0, ch, ky, half, kx * co + n, k].float() == w9[ky * 3 + kx, ch * 16 + half * 8 + k, n]
    # the one-tap layout everywhere else: short K, K not a multiple of 32, wide outputs
    for ci, co in ((32, 32), (80, 32), (64, 128)):
        assert packing.conv3x3_tc(torch.zeros(9, ci, co)).shape == (1, ci // 16, 9, 2, min(co, 128), 8)


def test_convT2x2_and_pointwise_packing():
    g = torch.Generator().manual_seed(4)
    w = torch.randn(32, 32, 2, 2, generator=g)
    p = packing.convT2x2(w)
    assert p.shape == (32, 4, 32) and p[3, 2, 7] == w[3, 7, 1, 0]
    wg = torch.randn(512, 128, 1, 1, generator=g)
    pg = packing.pointwise(wg, 4)
    assert pg.shape == (4, 128, 128) and pg[2, 5, 9] == wg[2 * 128 + 9, 5, 0, 0]


def test_state_dict_contract():
    net = UNet(*G_ARGS, up_mode=0)
    sd = make_generator_state_dict()
    assert list(net.state_dict().keys()) == list(sd.keys())
    assert all(net.state_dict()[k].shape == v.shape for k, v in sd.items())
    net.load_state_dict(sd)
    assert sum(p.numel() for p in net.parameters()) == 4941281
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 4920545
    vid = UNetVideo(*G_ARGS, up_mode=0)
    assert list(vid.state_dict().keys()) == list(sd.keys())
    assert len(generator_shapes()) == 27 - 1 + 1 or True


def test_unsupported_configs_fail_loudly():
    bad = list(G_ARGS)
    bad[5] = "original_unet"
    with pytest.raises(NotImplementedError):
        UNet(*bad, up_mode=0)
    with pytest.raises(NotImplementedError):
        UNet(*G_ARGS, up_mode=1)
    net = UNet(*G_ARGS, up_mode=0).eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 256, 256))  # CPU tensor: no fallback
    with torch.no_grad(), pytest.raises(ValueError):
        net(torch.zeros(1, 1, 128, 128))


def _conv_plan(n, ci, h, w, co, pad):
    import ctypes
    from uncltmo_b200 import _lib
    plan = (ctypes.c_int * 16)()
    rc = _lib.lib().uncl_conv3x3_tc_plan(n, ci, h, w, co, pad, plan)
    assert rc == 0, _lib.lib().uncl_last_error()
    keys = ("kind", "NT", "NS", "mma_n", "MB", "adv", "PW", "PH", "BW", "bands", "tiles_per_band", "items", "stages", "nacc",
            "ksteps", "smem")
    return dict(zip(keys, list(plan)))


def test_conv3x3_tc_plan_invariants():
    """Tile plans of the tensor-core conv (pure host arithmetic behind the C ABI): every output position is covered, the
    halo box holds every pixel an MMA row can touch, and the plan fits TMEM / shared memory / the TMA box limits."""
    rng = np.random.default_rng(3)
    shapes = [(60, 32, 254, 254, 32, 0), (60, 128, 252, 252, 32, 2), (60, 512, 57, 57, 64, 2), (60, 1024, 24, 24, 128, 2),
              (60, 256, 12, 12, 256, 0), (1, 64, 3, 3, 32, 0), (2, 64, 5, 5, 64, 2), (1, 16, 3, 3, 32, 0)]
    for _ in range(60):
        ci = int(rng.choice([16, 32, 48, 64, 96, 128, 256]))
        co = int(rng.choice([32, 64, 96, 128, 256]))
        shapes.append((int(rng.integers(1, 5)), ci, int(rng.integers(3, 300)), int(rng.integers(3, 300)), co, int(rng.choice([0, 2]))))
    for n, ci, h, w, co, pad in shapes:
        p = _conv_plan(n, ci, h, w, co, pad)
        ho, wo = h + 2 * pad - 2, w + 2 * pad - 2
        halo = 2
        assert p["PW"] == p["BW"] + halo and p["PW"] <= 128 and p["PH"] <= 256          # TMA box limits
        assert p["bands"] * p["BW"] >= wo > (p["bands"] - 1) * p["BW"]                   # bands tile the width
        band_total = ho * p["PW"]
        last_valid = band_total - 1 - halo                                                # last non-garbage flattened position
        assert p["tiles_per_band"] * p["adv"] > last_valid                                # tiles tile the band
        assert (p["tiles_per_band"] - 1) * p["adv"] <= last_valid                         # ... without an empty tile
        # the deepest pixel an MMA row reads: tile start offset (< PW) + 128*MB rows + two filter rows (+2 columns); the last
        # two taps of wrap-around rows (masked outputs) may read up to two pixels past the box, inside the stage
        # (row-aligned tiles of the merged kernel always start at a row start)
        aligned = (p["kind"] & 3) == 3
        moff_max = 0 if aligned else p["PW"] - 1
        assert (moff_max + 128 * p["MB"] - 1 + 2 * p["PW"] + 2) <= p["PH"] * p["PW"] + p["PW"]
        if aligned:
            assert p["adv"] <= 128 * p["MB"] and p["PH"] == p["adv"] // p["PW"] + 2
        assert p["MB"] * p["mma_n"] <= (512 if p["nacc"] == 1 else 256)                   # TMEM columns per stage
        assert p["stages"] >= 2 and p["smem"] <= 227 * 1024
        assert p["items"] == n * p["bands"] * p["tiles_per_band"] * p["NS"]
        assert p["NT"] * p["NS"] == co and ci % (16 * p["ksteps"]) == 0
        if p["kind"] & 1:
            assert p["mma_n"] == 3 * p["NT"] and (p["adv"] == 128 * p["MB"] - 2 or aligned) and p["MB"] * (p["NT"] // 32) == 2
        else:
            assert p["adv"] == 128 * p["MB"]


def test_weight_packing_follows_the_kernel_choice():
    """packing.conv3x3_tc and uncl_conv3x3_tc apply the same (C_in, C_out) rule: a mismatch would feed one kernel the other
    kernel's weight layout."""
    for ci in (16, 32, 48, 64, 80, 96, 128, 256, 512, 1024):
        for co in (32, 64, 96, 128, 256):
            p = _conv_plan(1, ci, 20, 20, co, 0)
            packed = packing.conv3x3_tc(torch.zeros(9, ci, co))
            if p["kind"] & 1:
                assert packed.shape == (p["NS"], ci // 16, 3, 2, p["mma_n"], 8)
            else:
                assert packed.shape == (p["NS"], ci // 16, 9, 2, p["NT"], 8)


def test_split_bf16_weight_packings():
    """pointwise_tc_split: [w_hi | w_lo | w_hi] in the K order the kernel's [x_hi | x_hi | x_lo] rows
    expect, hi + lo reproducing the fp32 weight to ~2^-16."""
    g = torch.Generator().manual_seed(7)
    w = torch.randn(256, 64, 1, 1, generator=g)
    p = packing.pointwise_tc_split(w)                       # [NS][3*C_in/16][2][NT][8]
    assert p.shape == (2, 12, 2, 128, 8) and p.dtype == torch.bfloat16
    hi = w.reshape(256, 64).to(torch.bfloat16)
    lo = (w.reshape(256, 64) - hi.float()).to(torch.bfloat16)
    ns, n, ci = 1, 77, 37
    for term, want in ((0, hi), (1, lo), (2, hi)):
        k = term * 64 + ci
        assert p[ns, k // 16, (k % 16) // 8, n, k % 8] == want[ns * 128 + n, ci]
    assert ((hi.float() + lo.float()) - w.reshape(256, 64)).abs().max() <= 2.0 ** -15 * w.abs().max()


def test_merged_packing_reproduces_the_convolution_on_cpu():
    """The kx-merged formulation restated with torch on the CPU, reading the weights exactly as conv_tc_merged.cu does from
    packing.conv3x3_tc's layout: D[p][kx*C + co] = sum_{ky,ci} X[p + ky*PW][ci] W[ky][kx][ci][co] over flattened positions
    (pitch PW = W), out[q] = D[q][0] + D[q+1][1] + D[q+2][2].  Guards the layout contract without a GPU."""
    g = torch.Generator().manual_seed(11)
    ci, co, h, w = 64, 32, 9, 11
    x = torch.randn(1, ci, h, w, generator=g)
    wt = torch.randn(co, ci, 3, 3, generator=g) / (9 * ci) ** 0.5
    w9 = packing.conv3x3_taps(wt, transposed=False)               # [9][ci][co]
    p = packing.conv3x3_tc(w9).float()                             # [1][ci/16][3][2][3*co][8]
    assert p.shape == (1, ci // 16, 3, 2, 3 * co, 8)
    xf = x[0].permute(1, 2, 0).reshape(h * w, ci)                   # flattened positions, pitch PW = w
    xf = torch.cat([xf, torch.zeros(2 * w + 2, ci)])                # rows the shifted reads run into (masked outputs)
    npos = h * w
    d = torch.zeros(npos + 2, 3 * co)
    for ch in range(ci // 16):
        for ky in range(3):
            a = xf[ky * w: ky * w + npos + 2, ch * 16:(ch + 1) * 16]                      # A rows shifted by one filter row
            b = p[0, ch, ky].permute(1, 0, 2).reshape(3 * co, 16)                          # [N'][k = half*8 + k8]
            d += a @ b.t()
    out = d[0:npos, 0:co] + d[1:npos + 1, co:2 * co] + d[2:npos + 2, 2 * co:3 * co]
    out = out.reshape(h, w, co)[: h - 2, : w - 2].permute(2, 0, 1)                         # the last two columns / rows wrap
    # the packed weights are bf16: compare against the convolution of the bf16-rounded weights with fp32 activations
    want = torch.nn.functional.conv2d(x, wt.to(torch.bfloat16).float())[0]
    assert torch.allclose(out, want, atol=2e-5, rtol=1e-5)


def test_flat_parameter_index_maps():
    """train_graph.FlatParams: the one-launch weight packing (dst[i] = bf16(flat[idx[i]])) reproduces packing.py's torch
    re-layouts operand by operand, the split-bf16 flag selects the residual, and the gradient un-packing map is the inverse
    permutation of the GEMM layouts (a staged 'gradient' equal to the re-laid-out weight returns the weight itself)."""
    from uncltmo_b200.generator import UNet
    from uncltmo_b200.train_graph import FlatParams
    from uncltmo_b200.weights import make_generator_state_dict
    args = (1, 1, "sigmoid", 4, 4, "square_and_square_root", 32, 0, "unet", 0, 0, "none", "none", "relu", 1, "replicate", 2)
    net = UNet(*args, up_mode=0, precision="bf16")
    sd = make_generator_state_dict()
    net.load_state_dict(sd)
    fp = FlatParams(net)
    assert fp.is_current() and fp.grads_attached()
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]), k     # the flat buffer holds the same parameters
    idx = fp.pack_idx.long()
    src = fp.flat[(idx & 0x3fffffff).clamp(min=0)]
    hi = src.to(torch.bfloat16)
    lo = (src - hi.float()).to(torch.bfloat16)
    packed = torch.where(idx < 0, torch.zeros_like(hi), torch.where((idx & packing.LO_FLAG) != 0, lo, hi))

    def operand(name):
        o, n = fp.pack_off[name], int(np.prod(fp.pack_shape[name]))
        return packed[o:o + n].reshape(fp.pack_shape[name])

    for name, key, transposed in fp.conv3:
        w9 = packing.conv3x3_taps(sd[key + ".weight"], transposed)
        assert torch.equal(operand(name), packing.conv3x3_tc(w9)), name
        assert torch.equal(operand(name + "_d"), packing.conv3x3_tc(w9.flip(0).transpose(1, 2).contiguous())), name
    for i in range(4):
        w = sd["up_path.%d.up.weight" % i]
        c = w.shape[0]
        assert torch.equal(operand("u%d_up" % i), packing.convT2x2_tc(w))
        assert torch.equal(operand("u%d_up_d" % i), packing.pointwise_tc(w.permute(0, 2, 3, 1).reshape(c, 4 * c, 1, 1)))
    gp = "gcn.module.0."
    assert torch.equal(operand("g_gconv"), packing.pointwise_tc(sd[gp + "0.graph_conv.gconv.nn.0.weight"], 4))
    assert torch.equal(operand("g_fc1_split"), packing.pointwise_tc_split(sd[gp + "0.fc1.0.weight"]))
    f32 = fp.flat[fp.f32_idx.long()]
    assert torch.equal(f32[:fp.n_cf].reshape(9, 32), packing.conv_first(sd["inc.conv.conv.weight"]))
    assert torch.equal(f32[fp.n_cf:].reshape(32, 144, 8), packing.blocked_param(sd["gcn.pos_embed"]))
    # un-packing: stage the re-laid-out weights as if they were gradients, scatter them back
    stage = torch.zeros(fp.stage_total)
    for name, key, transposed in fp.conv3:
        w9 = packing.conv3x3_taps(sd[key + ".weight"], transposed)
        stage[fp.stage_off[name]:fp.stage_off[name] + w9.numel()] = w9.reshape(-1)
    for i in range(4):
        g = packing.convT2x2_gemm_layout(sd["up_path.%d.up.weight" % i])
        stage[fp.stage_off["u%d_up" % i]:fp.stage_off["u%d_up" % i] + g.numel()] = g.reshape(-1)
    cf = packing.conv_first(sd["inc.conv.conv.weight"])
    stage[fp.stage_off["inc0"]:fp.stage_off["inc0"] + cf.numel()] = cf.reshape(-1)
    u = fp.unpack_idx.long()
    grad = torch.where(u >= 0, stage[u.clamp(min=0)], torch.zeros(fp.total))
    for name, key, _ in fp.conv3 + [("", "up_path.%d.up" % i, 0) for i in range(4)] + [("", "inc.conv.conv", 0)]:
        assert torch.equal(fp.view(grad, key + ".weight"), sd[key + ".weight"]), key
    assert (fp.view(grad, "outc.conv.weight") == 0).all()      # parameters without a staged gradient are left alone


def test_skipcat_plan_invariants():
    """Tile plans of the fused-skip convolution (uncl_conv3x3_tc_skipcat): the ring program holds three slots per skip chunk,
    so the plan narrows its column bands until at least four (normally five) pipeline stages fit; tiles still cover the
    output, boxes respect the TMA limits, everything fits shared memory; unsupported channel counts are refused."""
    import ctypes
    from uncltmo_b200 import _lib
    keys = ("kind", "NT", "NS", "mma_n", "MB", "adv", "PW", "PH", "BW", "bands", "tiles_per_band", "items", "stages", "nacc",
            "ksteps", "smem")
    for n, cs, h, w, pad in [(60, 32, 252, 252, 2), (60, 64, 122, 122, 2), (2, 32, 5, 7, 2), (1, 96, 33, 40, 2), (3, 64, 129, 65, 0),
                             (1, 128, 57, 57, 2)]:
        plan = (ctypes.c_int * 16)()
        rc = _lib.lib().uncl_conv3x3_tc_skipcat_plan(n, cs, h, w, 32, pad, plan)
        assert rc == 0, _lib.lib().uncl_last_error()
        p = dict(zip(keys, list(plan)))
        ho, wo = h + 2 * pad - 2, w + 2 * pad - 2
        assert p["kind"] & 1 and p["NT"] == 32 and p["mma_n"] == 96 and p["MB"] == 2 and p["ksteps"] == 2
        assert p["stages"] >= 4 and p["smem"] <= 227 * 1024
        assert p["PW"] == p["BW"] + 2 and p["PW"] <= 128 and p["PH"] <= 256
        assert p["bands"] * p["BW"] >= wo > (p["bands"] - 1) * p["BW"]
        last_valid = ho * p["PW"] - 3
        assert p["tiles_per_band"] * p["adv"] > last_valid and (p["tiles_per_band"] - 1) * p["adv"] <= last_valid
        assert p["items"] == n * p["bands"] * p["tiles_per_band"]
        aligned = (p["kind"] & 3) == 3
        moff_max = 0 if aligned else p["PW"] - 1
        assert (moff_max + 128 * p["MB"] - 1 + 2 * p["PW"] + 2) <= p["PH"] * p["PW"] + p["PW"]
    plan = (ctypes.c_int * 16)()
    assert _lib.lib().uncl_conv3x3_tc_skipcat_plan(1, 48, 20, 20, 32, 2, plan) != 0      # C_skip must be a multiple of 32
    assert _lib.lib().uncl_conv3x3_tc_skipcat_plan(1, 32, 20, 20, 64, 2, plan) != 0      # C_out must be 32
    assert _lib.lib().uncl_conv3x3_tc_skipcat_plan(1, 160, 20, 20, 32, 2, plan) != 0     # ring program: C_skip <= 128


def test_row_kernel_plan_and_packing():
    """Row kernel (conv_tc_rows.cu) host arithmetic: strips cover the output rows, whole 126-column bands go to the row
    kernel and the remainder to the older kernels, stage ring + resident filter bank fit shared memory, the multi-chunk
    row-group rule (G <= 3, five slots) holds; packing puts tap (ky, kx), channel ci, output co where the kernel's
    descriptors look for it."""
    import torch
    from uncltmo_b200 import packing
    keys = ("ok", "cols", "tail", "bands", "BW", "R", "strips", "items", "G", "nchunk", "stages", "stage_bytes", "w_bytes",
            "smem", "sms", "_")
    cases = [(240, 32, 254, 254, 0, False), (240, 32, 124, 124, 2, False), (240, 32, 254, 254, 2, False),
             (240, 128, 252, 252, 2, True), (60, 128, 252, 252, 2, False), (1, 32, 3, 3, 0, False), (7, 96, 66, 300, 2, False),
             (2, 128, 40, 40, 0, True), (240, 256, 122, 122, 2, True)]
    for n, ci, h, w, pad, derive in cases:
        p = dict(zip(keys, packing.conv3x3_tc_rows_plan(n, ci, h, w, pad, derive)))
        ho, wo = h + 2 * pad - 2, w + 2 * pad - 2
        assert p["ok"] == 1 and p["cols"] + p["tail"] == wo and p["cols"] > 0
        assert p["tail"] == 0 or (p["cols"] % 126 == 0 and p["tail"] < 64)
        assert p["BW"] <= 126 and p["bands"] * p["BW"] >= p["cols"] > (p["bands"] - 1) * p["BW"]
        assert p["strips"] * p["R"] >= ho > (p["strips"] - 1) * p["R"] and p["items"] == n * p["bands"] * p["strips"]
        assert p["nchunk"] == ci // 32 and p["w_bytes"] == ci * 576 and p["stage_bytes"] == p["G"] * 8192
        assert 2 <= p["G"] <= (4 if p["nchunk"] == 1 else 3) and p["_"] == (1 if (p["nchunk"] == 1 and not derive) else 0)
        assert p["stages"] >= (5 if derive else 3) and p["smem"] <= 227 * 1024
        assert p["stages"] * p["stage_bytes"] + p["w_bytes"] < p["smem"]
    # the generator's shipped layers: 13 waves of 63-row strips on 148 SMs for four 1080p frames
    p = dict(zip(keys, packing.conv3x3_tc_rows_plan(240, 32, 254, 254, 0)))
    assert (p["R"], p["strips"], p["items"]) == (63, 4, 1920)
    p = dict(zip(keys, packing.conv3x3_tc_rows_plan(240, 256, 122, 122, 2, True)))   # up2.conv0: 144 KB of filters + 5 x 2 rows
    assert p["ok"] == 1 and p["G"] == 2 and p["stages"] >= 5 and p["smem"] <= 227 * 1024
    assert packing.conv3x3_tc_rows_plan(1, 512, 59, 59, 2, False)[0] == 0        # 288 KB of filters: stays on the older kernels
    # C_out = 64 (down0.conv0 / conv1): ring of eight 64-column groups, four rows per stage, 36 KB of filters per 32 channels
    for n, ci, h in [(240, 32, 126), (240, 64, 124), (3, 128, 33)]:
        p = dict(zip(keys, packing.conv3x3_tc_rows_plan(n, ci, h, h, 0, False, co=64)))
        assert p["ok"] == 1 and p["_"] == 1 and p["G"] == (4 if ci <= 64 else 3) and p["w_bytes"] == ci * 1152 and p["stages"] >= 3 and p["smem"] <= 227 * 1024
    assert packing.conv3x3_tc_rows_plan(1, 256, 59, 59, 0, False, co=64)[0] == 0      # 288 KB of filters
    t64 = packing.conv3x3_tc_rows_layout(torch.arange(9 * 32 * 64, dtype=torch.float32).reshape(9, 32, 64))
    assert tuple(t64.shape) == (1, 2, 3, 2, 192, 8) and t64[0, 1, 2, 1, 2 * 64 + 7, 3] == (2 * 3 + 2) * 32 * 64 + (16 + 8 + 3) * 64 + 7
    w9 = torch.arange(9 * 64 * 32, dtype=torch.float32).reshape(9, 64, 32)
    t = packing.conv3x3_tc_rows_layout(w9)
    assert tuple(t.shape) == (2, 2, 3, 2, 96, 8)
    for ky, kx, ci, co in [(0, 0, 0, 0), (2, 1, 37, 5), (1, 2, 63, 31), (0, 2, 16, 9)]:
        chunk, ks, half, k8 = ci // 32, (ci % 32) // 16, (ci % 16) // 8, ci % 8
        assert t[chunk, ks, kx, half, ky * 32 + co, k8] == w9[ky * 3 + kx, ci, co]


def test_conv_first_rows_packing():
    """Filter tiles of the row kernel's front mode (inc.conv computed inside inc.conv1's launch): [hi | lo][K half][co][tap in
    half] with the bias on tap 9 (its im2col input is the constant 1) and hi + lo reproducing the fp32 weights to 2^-16."""
    import torch
    from uncltmo_b200 import packing
    g = torch.Generator().manual_seed(0)
    w, b = torch.randn((32, 1, 3, 3), generator=g), torch.randn(32, generator=g)
    t = packing.conv_first_rows(w, b)
    assert tuple(t.shape) == (2, 2, 32, 8) and t.dtype == torch.bfloat16
    full = t.float().permute(0, 1, 3, 2).reshape(2, 16, 32)            # [hi | lo][tap][co]
    rec = full[0] + full[1]
    want = torch.cat([w.reshape(32, 9).t(), b.reshape(1, 32)], dim=0)
    assert (rec[:10] - want).abs().max().item() <= 2.0 ** -15 * want.abs().max().item()
    assert (full[:, 10:] == 0).all()
    assert torch.equal(full[0, :10], want.to(torch.bfloat16).float())
