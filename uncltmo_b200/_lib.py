"""ctypes binding of libuncltmo_b200.so, generated from the C-ABI header include/uncltmo_b200.h.

There is no CPU or PyTorch fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# UNCL_LIB selects another build of the same ABI (tools/ use the -DUNCL_PROBES variant); the default is the product library
LIB_PATH = os.environ.get("UNCL_LIB") or os.path.join(_HERE, "libuncltmo_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "uncltmo_b200.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
DTYPE_OF = {torch.float32: F32, torch.bfloat16: BF16}

_lib = None
_protos = None


def _ctype(decl):
    decl = decl.strip()
    if "*" in decl or decl.startswith("uncl_stream_t"):
        return ctypes.c_void_p
    base = decl.split()[0] if not decl.startswith("unsigned") else "unsigned"
    return {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "double": ctypes.c_double,
            "unsigned": ctypes.c_uint}[base]


def prototypes():
    """{name: (restype, [argtypes], takes_stream)} parsed from the header (single source of truth)."""
    global _protos
    if _protos is None:
        src = re.sub(r"/\*.*?\*/", "", open(HEADER_PATH).read(), flags=re.S)
        out = {}
        for m in re.finditer(r"(const char\*|int|long)\s+(uncl_\w+)\s*\(([^)]*)\)\s*;", src):
            ret, name, args = m.group(1), m.group(2), m.group(3).strip()
            arglist = [] if args in ("", "void") else [a for a in args.split(",")]
            out[name] = (ctypes.c_char_p if "char" in ret else (ctypes.c_long if ret == "long" else ctypes.c_int), [_ctype(a) for a in arglist],
                         bool(arglist) and arglist[-1].strip().startswith("uncl_stream_t"))
        _protos = out
    return _protos


def declared_symbols():
    return sorted(prototypes())


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("uncltmo_b200: %s is missing - run `python -m uncltmo_b200.build` (or "
                               "__graft_entry__.build()); there is no fallback path" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (ret, args, _) in prototypes().items():
            fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = ret
            fn.argtypes = args
        _lib = l
    return _lib


def ptr(t):
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("uncltmo_b200 kernels need CUDA tensors (got %s); there is no CPU path" % t.device)
        return t.data_ptr()
    return t


# kernels launched by one C-ABI call (1 unless listed); used for the launch count bench.py reports
KERNELS_PER_CALL = {"uncl_frame_fused_supported": 0, "uncl_frame_normalise_pad": 5, "uncl_percentile_pair": 3, "uncl_plane_mean_contrast": 2,
                    "uncl_disc_forward": 3, "uncl_nce_fwd": 2, "uncl_tv_loss": 2}
_launches = 0
_timing = None


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


def start_call_timing():
    """Record a CUDA-event pair around every subsequent call (diagnostic pass; adds host overhead)."""
    global _timing
    _timing = []


def stop_call_timing():
    """-> [(entry point, device ms)] since start_call_timing(); synchronises."""
    global _timing
    rec, _timing = _timing, None
    torch.cuda.synchronize()
    return [(name, e0.elapsed_time(e1)) for name, e0, e1 in rec]


def call(name, *args):
    """Invoke a C-ABI entry point on the current CUDA stream of the device the tensor arguments live on; raise
    RuntimeError on a non-zero return.  All tensors of a call must be on one device; when that is not the current device
    the call runs under torch.cuda.device(...) so that the kernels, the stream and the pointers agree."""
    global _launches
    l = lib()
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            if dev is None:
                dev = a.device
            elif a.device != dev:
                raise RuntimeError("%s: tensors on different devices (%s, %s)" % (name, dev, a.device))
    argv = [ptr(a) for a in args]
    if dev is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return call(name, *args)
    stream = torch.cuda.current_stream().cuda_stream
    if _timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(l, name)(*argv, stream)
        e1.record()
        _timing.append((name, e0, e1))
    else:
        rc = getattr(l, name)(*argv, stream)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, l.uncl_last_error().decode()))
    _launches += KERNELS_PER_CALL.get(name, 1)
    if name in ("uncl_conv3x3_tc_rows", "uncl_conv3x3_tc_rows_skipcat"):
        # the row kernel leaves the columns past the last whole 126-column band to a second launch of the older kernels
        # (conv_tc_rows.cu:rw_cols): args[10] = W, args[12] / args[11] = pad
        w, pad = args[10], (args[12] if name == "uncl_conv3x3_tc_rows" else args[11])
        wo = w + 2 * pad - 2
        if wo >= 126 and 0 < wo % 126 < 64:
            _launches += 1
