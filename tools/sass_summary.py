"""SASS summary of the shipped library: which kernels use the tensor cores / TMEM / TMA, and how (cuobjdump -sass).
python tools/sass_summary.py > profiles/<round>_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uncltmo_b200", "libuncltmo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, per, archs = None, collections.OrderedDict(), collections.Counter()
pat = {"UTCHMMA": r"\bUTCHMMA\b", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "UTCBAR": r"\bUTCBAR", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP",
       "LDTM": r"\bLDTM", "SYNCS": r"\bSYNCS", "MUFU.SQRT": r"MUFU\.SQRT", "RED/ATOM": r"\b(REDG|RED|ATOMG|ATOMS)\b"}
total = collections.Counter()
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        archs[m.group(1)] += 1
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[kern] = collections.Counter()
        continue
    if kern and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        per[kern]["insts"] += 1
        for k, p in pat.items():
            if re.search(p, line):
                per[kern][k] += 1
                total[k] += 1
print("# SASS summary of uncltmo_b200/libuncltmo_b200.so (cuobjdump -sass, CUDA 12.9)")
print("# code objects:", dict(archs))
print("# whole library:", ", ".join("%s %d" % kv for kv in total.items()))
print("# UTCHMMA = tcgen05.mma (kind::f16, cta_group::1 - tools/mma_probe2.cu holds the cta_group::2 measurement), UTCBAR = tcgen05.commit,")
print("# UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (weights), LDTM = tcgen05.ld, SYNCS = mbarrier ops")
print("%-110s %6s %8s %7s %8s %7s %5s" % ("kernel", "insts", "UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM"))
for k, c in per.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print("%-110s %6d %8d %7d %8d %7d %5d" % (k[:110], c["insts"], c["UTCHMMA"], c["UTCBAR"], c["UTMALDG"], c["UBLKCP"], c["LDTM"]))
