"""The generator's training pass on the inference design: bf16 C8-blocked activations end to end, one autograd node.

`GeneratorTrainFn` is a single torch.autograd.Function whose forward is the bf16 tensor-core network of
`generator._run_frame` (skip tensors emitted in place into their concat buffers, nothing converted or copied between
kernels) and whose backward is a hand-scheduled chain of library kernels:

  * every 3x3 data gradient runs on the forward's tcgen05 kernel with the ReLU backward of the producing layer fused into
    the epilogue (`uncl_conv3x3_tc_dgrad`), so the tensor it writes IS the next layer's bf16 gradient operand;
  * every 3x3 / up-conv weight gradient is a tensor-core GEMM over pixels (`uncl_conv3x3_wgrad_tc`, `uncl_pw_wgrad_tc`);
  * where gradient paths meet - a skip tensor fed by the decoder concat [x2 | up | x2^2 | sqrt(x2+eps)]
    (unet_parts.py:319-322) and by MaxPool2d (unet_parts.py:210-213) - one kernel combines, masks and reduces the bias
    gradient (`uncl_skip_pool_bwd`);
  * all weights are re-laid out for the forward and the data-gradient GEMMs by ONE gather launch per step from the flat
    fp32 parameter buffer (`uncl_pack_gather`, index maps built once from packing.py), and all GEMM-layout weight
    gradients return to the parameter layout by one more (`uncl_unpack_gather`);
  * parameter gradients are accumulated straight into one flat fp32 buffer that every `p.grad` is a view of - the same
    buffer the gradient all-reduce and the flat Adam kernel work on.

The graph block (12 x 12, 0.6 % of the FLOPs) keeps its per-operator autograd Functions (fp32 internals, KNN exact) as a
nested graph.  Reference: models/unet_multi_filters/Unet_singleFrame.py:177-213, unet_parts.py.
"""
import torch
from torch.autograd import Function

from . import _lib, packing
from . import autograd as A
from ._lib import ACT_RELU, BF16, F32, call

LEVELS = [(32, 252), (64, 122), (128, 57), (256, 24)]   # (channels, extent) of the four skip tensors


class FlatParams:
    """Flat fp32 storage for a generator's parameters and gradients + the packed operands derived from them."""

    def __init__(self, net):
        self.net = net
        self.named = [(k, p) for k, p in net.named_parameters()]
        dev = self.named[0][1].device
        self.device = dev
        self.offsets, off = {}, 0
        for k, p in self.named:
            self.offsets[k] = off
            off += (p.numel() + 7) // 8 * 8     # keep every tensor 32-byte aligned
        self.total = off
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(off, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for k, p in self.named:
                o = self.offsets[k]
                self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
        self._build_maps()
        self.attach_grads()

    # ------------------------------------------------------------------ bookkeeping
    def is_current(self):
        """False once something (a .to(), a load with assign=True ...) has moved a parameter out of the flat buffer."""
        for k, p in self.named:
            if p.data_ptr() != self.flat.data_ptr() + 4 * self.offsets[k]:
                return False
        return True

    def view(self, buf, k):
        p = dict(self.named)[k]
        o = self.offsets[k]
        return buf[o:o + p.numel()].view(p.shape)

    def attach_grads(self):
        """Make every trainable p.grad a view of the flat gradient buffer (zero_grad(set_to_none=True) detaches them)."""
        for k, p in self.named:
            if p.requires_grad:
                o = self.offsets[k]
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def grads_attached(self):
        k, p = next((k, p) for k, p in self.named if p.requires_grad)
        return p.grad is not None and p.grad.data_ptr() == self.grad.data_ptr() + 4 * self.offsets[k]

    # ------------------------------------------------------------------ index maps
    def _idx(self, k):
        p = dict(self.named)[k]
        o = self.offsets[k]
        return torch.arange(o, o + p.numel(), dtype=torch.int64).view(p.shape)

    def _build_maps(self):
        """pack map: bf16 operand buffer <- flat params; unpack map: flat grads <- GEMM-layout gradient staging."""
        net = self.net
        names3 = [("inc1", "inc.conv.conv1", False)]
        for i in range(4):
            names3 += [("d%d_0" % i, "down_path.%d.mpconv.1.conv" % i, False),
                       ("d%d_1" % i, "down_path.%d.mpconv.1.conv1" % i, i == 3)]
        for i in range(4):
            names3 += [("u%d_0" % i, "up_path.%d.conv.conv" % i, True), ("u%d_1" % i, "up_path.%d.conv.conv1" % i, True)]
        self.conv3 = names3
        pieces, self.pack_off, self.pack_shape = [], {}, {}
        stage_pieces, self.stage_off = [], {}
        pos = [0]
        spos = [0]

        def add(name, idx):
            self.pack_off[name] = pos[0]
            self.pack_shape[name] = tuple(idx.shape)
            pieces.append(idx.reshape(-1))
            pos[0] += (idx.numel() + 63) // 64 * 64      # 128-byte aligned operands (cp.async.bulk / TMA sources)
            if pieces[-1].numel() % 64:
                pieces.append(torch.full((64 - pieces[-1].numel() % 64,), -1, dtype=torch.int64))

        unpack = torch.full((self.total,), -1, dtype=torch.int64)

        def add_stage(name, layout_idx, numel):
            """layout_idx: for every element of the GEMM-layout gradient, the flat parameter index it belongs to."""
            self.stage_off[name] = spos[0]
            unpack[layout_idx.reshape(-1)] = spos[0] + torch.arange(numel, dtype=torch.int64)
            spos[0] += (numel + 7) // 8 * 8

        for name, key, transposed in names3:
            idx = self._idx(key + ".weight")
            w9 = packing.conv3x3_taps_layout(idx, transposed)            # [9][ci][co] of flat indices
            add(name, packing.conv3x3_tc_layout(w9))
            add(name + "_d", packing.conv3x3_tc_layout(packing.conv3x3_dgrad_taps_layout(w9)))
            add_stage(name, w9, w9.numel())
        for i in range(4):
            idx = self._idx("up_path.%d.up.weight" % i)
            add("u%d_up" % i, packing.convT2x2_tc_layout(idx))
            add("u%d_up_d" % i, packing.convT2x2_dgrad_layout(idx))
            g = packing.convT2x2_gemm_layout(idx)                         # [C][4C]
            add_stage("u%d_up" % i, g, g.numel())
        gp = "gcn.module.0."
        for name, key, groups in (("g_gconv", gp + "0.graph_conv.gconv.nn.0", 4), ("g_fc2", gp + "0.fc2.0", 1),
                                  ("f_fc1", gp + "1.fc1.0", 1), ("f_fc2", gp + "1.fc2.0", 1)):
            add(name, packing.pointwise_tc_layout(self._idx(key + ".weight"), groups))
        add("g_fc1_split", packing.pointwise_tc_split_index_layout(self._idx(gp + "0.fc1.0.weight")))
        # conv_first [9][32] gradient staging
        cf = packing.conv_first_layout(self._idx("inc.conv.conv.weight"))
        add_stage("inc0", cf, cf.numel())
        dev = self.device
        self.pack_idx = torch.cat(pieces).to(torch.int32).to(dev)
        self.packed = torch.zeros(self.pack_idx.numel(), device=dev, dtype=torch.bfloat16)
        self.unpack_idx = unpack.to(torch.int32).to(dev)
        self.stage = torch.zeros(spos[0], device=dev, dtype=torch.float32)
        self.stage_total = spos[0]
        # fp32 side operands: conv_first taps [9][32] and the blocked pos_embed
        f32_idx = torch.cat([cf.reshape(-1), packing.blocked_param_layout(self._idx("gcn.pos_embed")).reshape(-1)])
        self.f32_idx = f32_idx.to(torch.int32).to(dev)
        self.f32_buf = torch.zeros(f32_idx.numel(), device=dev, dtype=torch.float32)
        self.n_cf = cf.numel()
        self._packed_version = None

    # ------------------------------------------------------------------ packing (one gather launch per buffer)
    def version(self):
        return tuple(p._version for _, p in self.named)

    def pack(self, force=False):
        v = self.version()
        if force or v != self._packed_version:
            call("uncl_pack_gather", self.flat, self.pack_idx, self.packed, self.packed.numel())
            call("uncl_unpack_gather", self.flat, self.f32_idx, self.f32_buf, self.f32_buf.numel())
            self._packed_version = v
        return self

    def w(self, name):
        o = self.pack_off[name]
        n = 1
        for d in self.pack_shape[name]:
            n *= d
        return self.packed[o:o + n]

    def bias(self, key):
        return self.view(self.flat, key + ".bias")

    def P(self):
        """The operand dictionary `generator._run_frame` reads (same keys as _GeneratorBase._pack)."""
        P = {"inc0": (self.f32_buf[:self.n_cf], self.bias("inc.conv.conv"))}
        for name, key, _ in self.conv3:
            P[name] = (self.w(name), self.bias(key))
        for i in range(4):
            P["u%d_up" % i] = (self.w("u%d_up" % i), self.bias("up_path.%d.up" % i))
        gp = "gcn.module.0."
        P["pos"] = self.f32_buf[self.n_cf:]
        P["relpos"] = self.view(self.flat, gp + "0.relative_pos").reshape(144, 144)
        for name, key in (("g_gconv", gp + "0.graph_conv.gconv.nn.0"), ("g_fc2", gp + "0.fc2.0"), ("f_fc1", gp + "1.fc1.0"),
                          ("f_fc2", gp + "1.fc2.0")):
            P[name] = (self.w(name), self.bias(key))
        P["g_fc1_split"] = (self.w("g_fc1_split"), self.bias(gp + "0.fc1.0"))
        P["g_fc1"] = (None, self.bias(gp + "0.fc1.0"))
        P["outc"] = (self.view(self.flat, "outc.conv.weight").reshape(-1), self.bias("outc.conv"))
        return P

    def g(self, key):
        """gradient view of parameter `key` in the buffer the running backward accumulates into (`self.gbuf`: the flat
        p.grad buffer in side-effect mode, a scratch `delta` buffer handed to autograd in autograd mode)"""
        return self.view(self.gbuf, key)

    def begin_backward(self, through_autograd):
        if through_autograd:
            if getattr(self, "delta", None) is None:
                self.delta = torch.zeros_like(self.grad)
            else:
                self.delta.zero_()
            self.gbuf = self.delta
        else:
            if not self.grads_attached():     # first backward after zero_grad(set_to_none=True): a new accumulation cycle
                self.grad.zero_()
                self.attach_grads()
            self.gbuf = self.grad
        self.stage.zero_()

    def stage_view(self, name, numel):
        o = self.stage_off[name]
        return self.stage[o:o + numel]


def flat_params(net):
    fp = getattr(net, "_flat", None)
    if fp is None or not fp.is_current():
        fp = FlatParams(net)
        net._flat = fp
    return fp


def _bf(shape, dev):
    return torch.empty(shape, device=dev, dtype=torch.bfloat16)


class _Saved:
    pass


def forward_train(net, x, droppath_scale, prev=None):
    """bf16 forward that keeps what the backward needs.  Returns (out fp32 [N,1,256,256], up bf16 blocked, saved).
    prev: the saved state of the PREVIOUS frame of the clip (video generator, Unet.py:229-286): the first C/32 channels of
    the four max-pool inputs and of the four up-convolution inputs are taken from that frame's tensors."""
    fp = flat_params(net).pack()
    n, dev = x.shape[0], x.device
    S = _Saved()
    S.fp, S.n, S.x, S.prev = fp, n, x, prev
    f = 32

    def st(t):
        return t.stride(0)

    def conv3(name, key, src, ci, h, co, pad, dst, emit_skip=0):
        call("uncl_conv3x3_tc", src, st(src), fp.w(name), fp.bias(key), dst, st(dst), BF16, n, ci, h, h, co, pad, ACT_RELU,
             emit_skip, 0, None, None, None, None)

    S.cat = [_bf((n, 4 * c // 8, s, s, 8), dev) for c, s in LEVELS]
    S.a0 = _bf((n, f // 8, 254, 254, 8), dev)
    call("uncl_conv_first", x, fp.f32_buf[:fp.n_cf], fp.bias("inc.conv.conv"), S.a0, st(S.a0), n, 256, 256, f, ACT_RELU, BF16)
    conv3("inc1", "inc.conv.conv1", S.a0, f, 254, f, 0, S.cat[0], emit_skip=1)
    S.pooled, S.mid = [], []
    cur, cur_c, cur_s = S.cat[0], f, 252
    for i in range(4):
        ps = cur_s // 2
        pooled = _bf((n, cur_c // 8, ps, ps, 8), dev)
        pv = prev.cat[i] if prev is not None else None
        call("uncl_maxpool2", cur, st(cur), pv, st(pv) if pv is not None else 0, cur_c // 32 if pv is not None else 0, pooled,
             st(pooled), n, cur_c, cur_s, cur_s, BF16)
        co = cur_c * 2 if i < 3 else cur_c
        mid = _bf((n, co // 8, ps - 2, ps - 2, 8), dev)
        key = "down_path.%d.mpconv.1." % i
        conv3("d%d_0" % i, key + "conv", pooled, cur_c, ps, co, 0, mid)
        if i < 3:
            conv3("d%d_1" % i, key + "conv1", mid, co, ps - 2, co, 0, S.cat[i + 1], emit_skip=1)
            cur, cur_c, cur_s = S.cat[i + 1], co, ps - 4
        else:
            S.x4 = _bf((n, co // 8, ps, ps, 8), dev)
            conv3("d3_1", key + "conv1", mid, co, ps - 2, co, 2, S.x4)
            cur, cur_c, cur_s = S.x4, co, ps
        S.pooled.append(pooled)
        S.mid.append(mid)
    # ---- graph block: nested autograd graph over the per-operator Functions (fp32 internals)
    C = cur_c
    x4f = torch.empty((n, C // 8, 12, 12, 8), device=dev, dtype=torch.float32)
    call("uncl_convert", S.x4, BF16, x4f, F32, x4f.numel())
    S.x4f = x4f.requires_grad_(True)
    g, ffn = net.gcn.module[0][0], net.gcn.module[0][1]
    s0 = droppath_scale[0] if droppath_scale is not None else None
    s1 = droppath_scale[1] if droppath_scale is not None else None
    # The nested graph's leaves are FRESH detached views of the parameters (same storage): autograd ties a leaf's gradient
    # bookkeeping to the stream it first saw the leaf on, and a parameter first used in an eager step on the legacy stream
    # would make a later CUDA-graph capture of the same trainer fail ("legacy stream depends on a capturing stream").
    gkeys = [k for k, p in fp.named if k.startswith("gcn.") and p.requires_grad]
    leaf = {k: fp.view(fp.flat, k).detach().requires_grad_(True) for k in gkeys}
    gp = "gcn.module.0."
    with torch.enable_grad():
        x0 = A.AddPos.apply(S.x4f, leaf["gcn.pos_embed"])
        # fc1 feeds the discrete KNN selection: three-term bf16 split forward (~2^-16), bf16 tensor-core gradients
        y = A.PwConv.apply(x0, leaf[gp + "0.fc1.0.weight"], leaf[gp + "0.fc1.0.bias"], None, None, 1, False, "split")
        z = A.KnnAggregate.apply(y, g.relative_pos.detach().reshape(144, 144).float().contiguous())
        z2 = A.PwConv.apply(z, leaf[gp + "0.graph_conv.gconv.nn.0.weight"], leaf[gp + "0.graph_conv.gconv.nn.0.bias"], None, None,
                            4, True, True)
        x1 = A.PwConv.apply(z2, leaf[gp + "0.fc2.0.weight"], leaf[gp + "0.fc2.0.bias"], x0, s0, 1, False, True)
        f1 = A.PwConv.apply(x1, leaf[gp + "1.fc1.0.weight"], leaf[gp + "1.fc1.0.bias"], None, None, 1, True, True)
        S.gout_f = A.PwConv.apply(f1, leaf[gp + "1.fc2.0.weight"], leaf[gp + "1.fc2.0.bias"], x1, s1, 1, False, True)
    S.gcn_keys, S.gcn_leaves = gkeys, [leaf[k] for k in gkeys]
    gout = _bf((n, C // 8, 12, 12, 8), dev)
    call("uncl_convert", S.gout_f.detach(), F32, gout, BF16, gout.numel())
    S.gout = gout
    # ---- decoder
    S.ups_in, S.dmid, S.ups = [], [], []
    up, up_c, up_s = gout, C, 12
    for i in range(4):
        sk_c, sk_s = LEVELS[3 - i]
        cb = S.cat[3 - i]
        dst = cb[:, sk_c // 8:]
        if prev is not None:
            # the up-convolution reads cat(prev[:, :r], up[:, r:]); `up` itself stays intact (skip / mask / hand-over)
            pv = prev.gout if i == 0 else prev.ups[i - 1]
            up_in = up.clone()
            call("uncl_splice_channels", up_in, st(up_in), pv, st(pv), up_c // 32, n, up_s * up_s, BF16)
        else:
            up_in = up
        call("uncl_convT2x2_tc", up_in, st(up_in), fp.w("u%d_up" % i), fp.bias("up_path.%d.up" % i), dst, st(cb), BF16, n, up_c,
             up_s, up_s, sk_s, sk_s)
        S.ups_in.append(up_in)
        co = f if i >= 2 else up_c // 2
        mid = _bf((n, co // 8, sk_s + 2, sk_s + 2, 8), dev)
        key = "up_path.%d.conv." % i
        conv3("u%d_0" % i, key + "conv", cb, 4 * sk_c, sk_s, co, 2, mid)
        nxt = _bf((n, co // 8, sk_s + 4, sk_s + 4, 8), dev)
        conv3("u%d_1" % i, key + "conv1", mid, co, sk_s + 2, co, 2, nxt)
        S.dmid.append(mid)
        S.ups.append(nxt)
        up, up_c, up_s = nxt, co, sk_s + 4
    out = torch.empty((n, 1, 256, 256), device=dev, dtype=torch.float32)
    call("uncl_outc_sigmoid", up, st(up), fp.view(fp.flat, "outc.conv.weight").reshape(-1), fp.bias("outc.conv"), out, None, n,
         f, 256 * 256, BF16)
    S.out = out
    return out, up, S


SIDE_STREAM_WGRAD = True
_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def backward_train(S, d_out, d_feat, through_autograd=False):
    """One image batch: begin, the frame's backward, the final gradient un-packing."""
    S.fp.begin_backward(through_autograd)
    backward_frame(S, d_out, d_feat, None)
    backward_end(S.fp)


def backward_end(fp):
    # GEMM-layout weight gradients -> parameter layout, accumulated into the gradient buffer of this backward
    call("uncl_unpack_add", fp.stage, fp.unpack_idx, fp.gbuf, fp.total)


def backward_frame(S, d_out, d_feat, d_state):
    """d_out: fp32 [N,1,256,256] or None; d_feat: bf16 blocked [N,4,256,256,8] or None.  Accumulates every parameter
    gradient into `fp.gbuf`: the flat gradient buffer every p.grad is a view of (side-effect mode, the fast trainer), or a
    scratch buffer whose per-parameter views are returned to autograd (hooks, DDP and optimizers see ordinary gradients).

    Video recurrence: d_state = the gradients the NEXT frame of the clip sent to this frame's handed-over channel slices
    (dict: "cat"[i] for the four pool inputs, "gout", "ups"[i] for the three decoder tensors; bf16 dense [N,1,H,W,8], or
    None for the last frame); returns the same structure for the PREVIOUS frame (None for the first frame of a clip)."""
    fp, n = S.fp, S.n
    dev = S.x.device
    prev = S.prev
    rec = prev is not None
    d_prev = {"cat": [None] * 4, "gout": None, "ups": [None] * 3} if rec else None
    ds = d_state or {"cat": [None] * 4, "gout": None, "ups": [None] * 3}

    def st(t):
        return t.stride(0)

    # Weight and bias gradients are leaves of the backward: nothing downstream reads them before the final un-packing.
    # They run on a second stream next to the data-gradient chain (which, at 16 images, leaves most SMs idle on the deeper
    # layers); inside a CUDA-graph capture this becomes a parallel branch of the graph.  Joined before this frame returns.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if SIDE_STREAM_WGRAD else None
    if side is not None:
        side.wait_stream(main)

    def aside(fn, *deps):
        if side is None:
            return fn()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        for t in deps:
            t.record_stream(side)
        with torch.cuda.stream(side):
            fn()

    def wgrad(name, x, ci, h, dz, co, pad):
        aside(lambda: call("uncl_conv3x3_wgrad_tc", x, st(x), dz, fp.stage_view(name, 9 * ci * co), n, ci, h, h, co, pad), x, dz)

    def dgrad(name, dz, ci_d, h_d, co_d, pad_d, mask, out_dtype=torch.bfloat16):
        """data gradient of layer `name`: dz [n, ci_d, h_d] -> [n, co_d, h_d + 2*pad_d - 2], masked by `mask` (> 0)"""
        ho = h_d + 2 * pad_d - 2
        o = torch.empty((n, co_d // 8, ho, ho, 8), device=dev, dtype=out_dtype)
        call("uncl_conv3x3_tc_dgrad", dz, st(dz), fp.w(name + "_d"), mask, st(mask) if mask is not None else 0, o, st(o),
             _lib.DTYPE_OF[out_dtype], n, ci_d, h_d, h_d, co_d, pad_d)
        return o

    def bias_grad(dz, key, c, hw):
        aside(lambda: call("uncl_bias_grad_bf16", dz, st(dz), fp.g(key + ".bias"), n, c, hw), dz)

    f = 32
    # ---- out conv + sigmoid + feature gradient + ReLU of up3.conv1
    up3 = S.ups[3]
    dz = _bf((n, f // 8, 256, 256, 8), dev)
    call("uncl_outc_feat_bwd", d_out, S.out, up3, st(up3), d_feat, fp.view(fp.flat, "outc.conv.weight").reshape(-1), dz,
         fp.g("outc.conv.weight").reshape(-1), fp.g("outc.conv.bias"), fp.g("up_path.3.conv.conv1.bias"), n, f, 256 * 256)
    # ---- decoder, last stage first
    dcats = [None] * 4
    d_gout = None
    for i in range(3, -1, -1):
        sk_c, sk_s = LEVELS[3 - i]
        co = f if i >= 2 else (128 if i == 0 else 64)
        key = "up_path.%d.conv." % i
        mid, cb = S.dmid[i], S.cat[3 - i]
        wgrad("u%d_1" % i, mid, co, sk_s + 2, dz, co, 2)
        dz_mid = dgrad("u%d_1" % i, dz, co, sk_s + 4, co, 0, mid)
        bias_grad(dz_mid, key + "conv", co, (sk_s + 2) ** 2)
        wgrad("u%d_0" % i, cb, 4 * sk_c, sk_s, dz_mid, co, 2)
        dcat = dgrad("u%d_0" % i, dz_mid, co, sk_s + 2, 4 * sk_c, 0, None)
        dcats[3 - i] = dcat
        # up-convolution (k2 s2): its output gradient is the second channel group of dcat
        up_in = S.ups_in[i]
        c_up, h_up = up_in.shape[1] * 8, up_in.shape[2]
        s2d = _bf((n, 4 * c_up // 8, h_up, h_up, 8), dev)
        call("uncl_convT2x2_s2d_bf16", dcat[:, sk_c // 8:], st(dcat), s2d, fp.g("up_path.%d.up.bias" % i), n, c_up, h_up, h_up,
             sk_s, sk_s)
        aside(lambda up_in=up_in, s2d=s2d, c_up=c_up, h_up=h_up, i=i: call(
            "uncl_pw_wgrad_tc", up_in, st(up_in), s2d, st(s2d), fp.stage_view("u%d_up" % i, 4 * c_up * c_up), n, c_up, 4 * c_up,
            h_up, h_up), up_in, s2d)
        r_up = c_up // 32
        if i > 0:
            dz = _bf((n, c_up // 8, h_up, h_up, 8), dev)
            # mask = the (spliced) tensor the up-conv read: own channels by this frame's ReLU, handed-over ones by prev's
            call("uncl_pw_conv_tc_dgrad", s2d, st(s2d), fp.w("u%d_up_d" % i), up_in, st(up_in), dz, st(dz), BF16, n, 4 * c_up,
                 c_up, 1, h_up, h_up)
            if rec or ds["ups"][i - 1] is not None:
                if rec:
                    d_prev["ups"][i - 1] = _bf((n, 1, h_up, h_up, 8), dev)
                own = S.ups[i - 1]
                call("uncl_splice_grad", dz, st(dz), own, st(own), r_up, d_prev["ups"][i - 1] if rec else None, ds["ups"][i - 1],
                     n, h_up * h_up)
            bias_grad(dz, "up_path.%d.conv.conv1" % (i - 1), c_up, h_up * h_up)
        else:
            d_gout = torch.empty((n, c_up // 8, h_up, h_up, 8), device=dev, dtype=torch.float32)
            call("uncl_pw_conv_tc_dgrad", s2d, st(s2d), fp.w("u0_up_d"), None, 0, d_gout, st(d_gout), F32, n, 4 * c_up, c_up,
                 1, h_up, h_up)
            if rec:    # the graph block's output has no ReLU: plain slicing (12 x 12 tensors)
                d_prev["gout"] = torch.zeros((n, 1, h_up, h_up, 8), device=dev, dtype=torch.bfloat16)
                d_prev["gout"][..., :r_up] = d_gout[:, :1, :, :, :r_up].to(torch.bfloat16)
                d_gout[:, :1, :, :, :r_up] = 0
            if ds["gout"] is not None:
                d_gout[:, :1, :, :, :r_up] += ds["gout"][..., :r_up].float()
    # ---- graph block (nested autograd graph)
    grads = torch.autograd.grad([S.gout_f], [S.x4f] + S.gcn_leaves, [d_gout.reshape(S.gout_f.shape)], retain_graph=True,
                                allow_unused=True)
    for k, g in zip(S.gcn_keys, grads[1:]):
        if g is not None:
            fp.g(k).add_(g)
    C = 256
    dz = _bf((n, C // 8, 12, 12, 8), dev)
    call("uncl_relu_bwd_bias_out", grads[0].contiguous(), S.x4f.detach(), S.x4f.stride(0), dz, BF16,
         fp.g("down_path.3.mpconv.1.conv1.bias"), n, C, 144, 1)
    # ---- encoder, deepest stage first
    for i in range(3, -1, -1):
        key = "down_path.%d.mpconv.1." % i
        mid, pooled = S.mid[i], S.pooled[i]
        co, ci = mid.shape[1] * 8, pooled.shape[1] * 8
        ps = pooled.shape[2]
        pad1 = 2 if i == 3 else 0
        wgrad("d%d_1" % i, mid, co, ps - 2, dz, co, pad1)
        dz_mid = dgrad("d%d_1" % i, dz, co, dz.shape[2], co, 2 - pad1, mid)
        bias_grad(dz_mid, key + "conv", co, (ps - 2) ** 2)
        wgrad("d%d_0" % i, pooled, ci, ps, dz_mid, co, 0)
        dpool = dgrad("d%d_0" % i, dz_mid, co, ps - 2, ci, 2, None)
        # skip level i: x2 = first channel group of cat[i], produced by d{i-1}_1 (inc1 for i = 0)
        c, s = LEVELS[i]
        dz = _bf((n, c // 8, s, s, 8), dev)
        prod_bias = ("down_path.%d.mpconv.1.conv1" % (i - 1)) if i > 0 else "inc.conv.conv1"
        if rec or ds["cat"][i] is not None:
            pv = prev.cat[i] if rec else None
            if rec:
                d_prev["cat"][i] = _bf((n, 1, s, s, 8), dev)
            call("uncl_skip_pool_bwd_rec", S.cat[i], st(S.cat[i]), dcats[i], dpool, dz, fp.g(prod_bias + ".bias"), n, c, s, s,
                 pv, st(pv) if rec else 0, c // 32, d_prev["cat"][i] if rec else None, ds["cat"][i])
        else:
            call("uncl_skip_pool_bwd", S.cat[i], st(S.cat[i]), dcats[i], dpool, dz, fp.g(prod_bias + ".bias"), n, c, s, s)
    # ---- inc: conv1 (32 -> 32) and the first conv (1 -> 32, fp32 CUDA cores)
    wgrad("inc1", S.a0, f, 254, dz, f, 0)
    dz_a0 = dgrad("inc1", dz, f, 252, f, 2, S.a0)
    call("uncl_conv_first_wgrad_bias", S.x, dz_a0, BF16, fp.stage_view("inc0", 9 * f), fp.g("inc.conv.conv.bias"), n, 256, 256, f)
    if side is not None:
        main.wait_stream(side)
    return d_prev


class GeneratorTrainFn(Function):
    """(x, anchor, net, droppath, *params) -> (out, up).

    Side-effect mode (no `params`; UNet.forward_blocked, the fast trainer): `anchor` is a 1-element tensor that requires
    grad so that autograd schedules the backward; parameter gradients are accumulated straight into the flat buffer that
    every p.grad is a view of - no per-parameter AccumulateGrad work, no hooks.
    Autograd mode (`params` = the trainable parameters; UNet.forward, the drop-in surface): the backward returns one
    gradient per parameter, so hooks, nn.parallel.DistributedDataParallel and uncltmo_b200.dist.GradientBuckets work as
    with any module."""

    @staticmethod
    def forward(ctx, x, anchor, net, droppath_scale, *params):
        out, up, S = forward_train(net, x.contiguous().float(), droppath_scale)
        ctx.S = S
        ctx.keys = [k for k, p in S.fp.named if p.requires_grad] if params else None
        ctx.set_materialize_grads(False)   # an unused output (features at epoch > 9) arrives as None, not as zeros
        # fresh tensor objects for the outputs: `S` keeps `out` / `up` for the backward, and the objects a Function returns
        # get grad_fn = this node - ctx.S -> S.out -> grad_fn -> node -> ctx would be a reference cycle through C++ that the
        # garbage collector cannot see (every step's activations, 0.8 GB at 16 images, would stay allocated for ever)
        return out.detach(), up.detach()

    @staticmethod
    def backward(ctx, d_out, d_up):
        S = ctx.S
        d_out = d_out.contiguous().float() if d_out is not None else None
        if d_up is not None:
            d_up = d_up.contiguous()
            if d_up.dtype != torch.bfloat16:
                d_up = d_up.to(torch.bfloat16)
        backward_train(S, d_out, d_up, through_autograd=ctx.keys is not None)
        if ctx.keys is None:
            return None, None, None, None
        # clones: a later backward (retain_graph) reuses the scratch buffer while autograd may still hold these
        return (None, None, None, None) + tuple(S.fp.view(S.fp.delta, k).clone() for k in ctx.keys)


class VideoTrainFn(Function):
    """The recurrent (video) generator over a clip as ONE autograd node: (x [N,T,1,256,256], anchor, net, droppath scales
    per frame, *params) -> (out_0, up_0, out_1, up_1, ...).  Frame k's forward takes the first C/32 channels of eight stage
    inputs from frame k-1 (Unet.py:229-286); the backward walks the frames in reverse and hands the gradients of those
    slices back (they are not detached in the reference).  Parameter gradients of all frames accumulate in one staging
    buffer and are un-packed once."""

    @staticmethod
    def forward(ctx, x, anchor, net, scales, *params):
        x = x.contiguous().float()
        saved, outs, prev = [], [], None
        for k in range(x.shape[1]):
            out, up, S = forward_train(net, x[:, k].contiguous(), scales[k] if scales is not None else None, prev)
            saved.append(S)
            outs += [out.detach(), up.detach()]    # fresh objects: see GeneratorTrainFn.forward
            prev = S
        ctx.saved = saved
        ctx.keys = [k for k, p in saved[0].fp.named if p.requires_grad] if params else None
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved
        fp = saved[0].fp
        fp.begin_backward(ctx.keys is not None)
        d_state = None
        for k in range(len(saved) - 1, -1, -1):
            d_out, d_up = grads[2 * k], grads[2 * k + 1]
            d_out = d_out.contiguous().float() if d_out is not None else None
            if d_up is not None:
                d_up = d_up.contiguous()
                if d_up.dtype != torch.bfloat16:
                    d_up = d_up.to(torch.bfloat16)
            d_state = backward_frame(saved[k], d_out, d_up, d_state)
        backward_end(fp)
        if ctx.keys is None:
            return None, None, None, None
        return (None, None, None, None) + tuple(fp.view(fp.delta, k).clone() for k in ctx.keys)
