#!/usr/bin/env python3
"""SASS opcode evidence per hot kernel of libavsr_b200.so (cuobjdump; no GPU needed): counts of the Blackwell-native
instructions (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, ELECT) and the top opcodes.
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "auto_avsr_b200", "csrc", "libavsr_b200.so")
WANT = [r"attention_f16_kernel", r"gemm_tc2_kernelILi0ELi384ELb1ELb0ELb1E", r"gemm_tc2_kernelILi0ELi256ELb0ELb0ELb1E",
        r"gemm_tc2_kernelILi0ELi128ELb0ELb1ELb1E", r"gemm_tc2_kernelILi1ELi256", r"gemm_tc2_kernelILi3ELi256",
        r"gemm_tc2_kernelILi5ELi512", r"gemm_tc_kernelILi4ELi256ELi2E6__half", r"ln_kernelILi0ELb0ELi2ELi4ELi3E",
        r"ln_kernelILi3ELb1ELi2ELi4ELi3E", r"dwconv_bn_silu_kernelILi31E"]
if len(sys.argv) > 1:          # patterns on the command line replace the default kernel list
    WANT = sys.argv[1:]
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "MUFU", "FSEL", "HMMA", "FFMA", "LDG", "STG"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
print("SASS opcode histogram of", os.path.relpath(LIB, ROOT), "(sm_100a)")
for pat in WANT:
    for name, cnt in funcs.items():
        if re.search(pat, name):
            tot = sum(cnt.values())
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
            print(f"\n{demangled}\n  {tot} instructions; " + "  ".join(f"{k} {cnt[k]}" for k in KEY if cnt[k]))
            print("  top: " + "  ".join(f"{k} {v}" for k, v in cnt.most_common(10)))
            break
