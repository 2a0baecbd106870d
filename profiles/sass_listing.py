"""Per-kernel counts of the SASS instructions that identify the data path (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM,
cp.async.bulk = UBLKCP, packed fp32 = FFMA2 / FADD2, ...) from the built library; no GPU needed.

    python profiles/sass_listing.py > profiles/r02/sass_listing.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "neural-ficititious-self-play-in-imperfect-information-games_b200", "libnfsp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCMMA|LDTM|STTM|UTCBAR|UTCATOMSWS|UBLKCP|UTMALDG|SYNCS|FFMA2|FADD2|LDCU|ATOMS|ATOMG|REDG|RED|BAR|ELECT|"
                 r"CREDUX|LDS|STS|LDG|STG)\b")
cur, counts, samples = None, collections.OrderedDict(), collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur], samples[cur] = collections.Counter(), {}
        continue
    if cur and "/*" in line:
        m = pat.search(line)
        if m:
            op = m.group(1)
            counts[cur][op] += 1
            if op in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UTCATOMSWS") and op not in samples[cur]:
                samples[cur][op] = line.strip()[:110]
print("# cuobjdump -sass libnfsp_b200.so (sm_100a), per kernel: counts of the instructions that identify the data path.")
print("# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc,")
print("# UBLKCP = cp.async.bulk (TMA unit), SYNCS = mbarrier ops, FFMA2/FADD2 = packed fp32, LDCU = uniform constant load")
print("# regenerate: python profiles/sass_listing.py > profiles/r02/sass_listing.txt\n")
for fn, c in counts.items():
    short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print("%s\n    %s" % (short[:140], "  ".join("%s=%d" % (k, v) for k, v in sorted(c.items()))))
    for op, l in samples[fn].items():
        print("    first %-10s %s" % (op, l))
