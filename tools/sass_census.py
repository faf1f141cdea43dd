#!/usr/bin/env python
"""Dev tool: opcode census of the product library's SASS (no GPU needed): which kernels are on the tcgen05 / TMA path.

    python tools/sass_census.py [tag]   ->  profiles/<tag>_sass_census.md
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "thermompnn_b200", "libthermompnn_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FFMA", "MUFU", "LDGSTS", "UCGABAR", "ELECT"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
census, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        census[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        census[cur]["total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or op.startswith(o + "_"):
                census[cur][o] += 1
                break
out = [f"# Round {tag[1:]}: SASS opcode census of `thermompnn_b200/libthermompnn_b200.so`", "",
       "`cuobjdump -sass` of the product library (sm_100a), instructions per kernel (static counts). `UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = "
       "tcgen05.ld/st, `UTMALDG` = TMA tensor-map load, `UBLKCP` = cp.async.bulk, `SYNCS` = mbarrier ops, `HMMA` = warp-level mma.sync, "
       "`FFMA2` = packed fp32 FMA, `LDGSTS` = cp.async, `UCGABAR` = cluster barrier.", "",
       "| kernel | total | " + " | ".join(OPS) + " |", "|---|---|" + "---|" * len(OPS)]
for k, c in census.items():
    out.append(f"| `{k[:70]}` | {c['total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_census.md")
open(path, "w").write("\n".join(out) + "\n")
print(open(path).read())
