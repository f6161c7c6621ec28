"""Builds libuncltmo_b200.so in-tree with nvcc for sm_100a (no PTX, no other arch).

`python -m uncltmo_b200.build`            product library (no probes, no environment knobs, no mutable globals)
`python -m uncltmo_b200.build --probes`   libuncltmo_b200_probes.so for tools/ only: -DUNCL_PROBES compiles in the timing
                                          switches (UNCL_PROBE_FLAGS ...) and the per-role cycle counters of the
                                          tensor-core kernels; select it with UNCL_LIB=<path> (uncltmo_b200/_lib.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(os.path.dirname(HERE), "include", "uncltmo_b200.h")
LIB = os.path.join(HERE, "libuncltmo_b200.so")
LIB_PROBES = os.path.join(HERE, "libuncltmo_b200_probes.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    """Everything a translation unit may include: the csrc headers and the public C-ABI header (the ctypes prototypes
    are parsed from it, so a signature edit must rebuild the library it describes)."""
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in sources() + _deps())


def build(force=False, verbose=False, probes=False):
    lib = LIB_PROBES if probes else LIB
    if not force and not needs_build(lib):
        return lib
    objs = []
    objdir = os.path.join(HERE, "build", "probes" if probes else "")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    dep_time = max(os.path.getmtime(d) for d in _deps())
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), dep_time):
            continue
        cmd = [NVCC] + FLAGS + (["-DUNCL_PROBES"] if probes else []) + ["-I", os.path.dirname(HEADER), "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (os.path.basename(src), out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(objdir, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv))
