#!/usr/bin/env python
"""Digest of `cuobjdump -sass rmcv_b200/librmcv_b200.so`: per kernel the instruction count, the mnemonics that prove what the
kernel is built from (1-D bulk TMA = UBLKCP, mbarriers = SYNCS, dp4a = IDP.4A, cp.async = LDGSTS, cluster barrier = UCGABAR_*,
distributed shared memory = MAPA / ATOMS..., warp primitives) and the top of its instruction mix.  usage: sass_summary.py [so]"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "rmcv_b200/librmcv_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = None
mix = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*\)$", "", kern)
        mix[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        mix[kern][m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "IDP.4A", "LDGSTS", "UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "MATCH.ANY", "REDUX", "SHFL", "VOTE", "ATOMS", "ATOMG", "RED", "DFMA", "DMUL", "MUFU"]
print("# SASS digest of", so, "(sm_100a)\n")
print("| kernel | instructions | " + " | ".join(KEY) + " | top of the mix |")
print("|---|---:|" + "---:|" * len(KEY) + "---|")
for k, c in mix.items():
    if not c:
        continue
    tot = sum(c.values())
    def cnt(key):
        return sum(v for op, v in c.items() if op == key or op.startswith(key + ".") or op.startswith(key))
    top = ", ".join("%s %d" % (op.split(".")[0] if False else op, v) for op, v in c.most_common(6))
    print("| `%s` | %d | " % (k.replace("|", "\\|")[:110], tot) + " | ".join(str(cnt(x)) for x in KEY) + " | " + top + " |")
