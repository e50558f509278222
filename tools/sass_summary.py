#!/usr/bin/env python
"""Counts of the SASS mnemonics that show what each kernel of libedmp_b200.so runs on (tcgen05 tensor cores, TMEM, the TMA
unit's bulk copies, mbarriers, packed fp32), per kernel: `cuobjdump -sass` here in the build container (no GPU needed).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edmp_b200", "libedmp_b200.so")
MNEMONICS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCQMMA", r"\bUTCQMMA"),
             ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("UTCATOMSWS", r"\bUTCATOMSWS"),
             ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"),
             ("FADD2/FMUL2", r"\bF(ADD|MUL)2"), ("MUFU", r"\bMUFU"), ("HMMA", r"\bHMMA"), ("STL/LDL", r"\b(STL|LDL)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("edmp::", "")
            kernels[name] = collections.Counter()
            kernels[name]["instructions"] = 0
            continue
        if name is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[name]["instructions"] += 1
            for label, pat in MNEMONICS:
                if re.search(pat, line):
                    kernels[name][label] += 1
    labels = ["instructions"] + [l for l, _ in MNEMONICS]
    print("# SASS mnemonic counts per kernel of edmp_b200/libedmp_b200.so (cuobjdump -sass, sm_100a; tools/sass_summary.py)")
    print("# UTCHMMA = tcgen05.mma kind::f16 / tf32 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
    print("# UBLKCP = cp.async.bulk (1-D bulk copies on the TMA unit; no tensor maps: UTMALDG = 0), SYNCS = mbarrier ops,")
    print("# FFMA2 = packed fma.rn.f32x2, STL/LDL = local-memory (stack / spill) traffic")
    w = max(len(k) for k in kernels) + 2
    print("%-*s" % (w, "kernel") + "".join("%14s" % l for l in labels))
    tot = collections.Counter()
    for k, c in kernels.items():
        if not any(c[l] for l in labels[1:8]) and "conv_fused" in k:
            tot["simt"] += 1
            continue
        print("%-*s" % (w, k) + "".join("%14d" % c[l] for l in labels))
        for l in labels:
            tot[l] += c[l]
    print("%-*s" % (w, "total (listed kernels)") + "".join("%14d" % tot[l] for l in labels))
    print("# + %d conv_fused_kernel<...> instantiations (fp32 CUDA-core parity mode, no tensor / TMA instructions)" % tot["simt"])


if __name__ == "__main__":
    main()
