#!/usr/bin/env python
"""Dev tool: per-phase cycle breakdown of the tensor-core edge kernels (block 0), needs TMPNN_TC_TRACE=1."""
import os, sys
os.environ["TMPNN_TC_TRACE"] = "1"; os.environ["TMPNN_LIB"] = "dev"   # needs the -DTMPNN_DEV build (python __graft_entry__.py --dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.synth import config3
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
eng = model._get_engine()
p = featurize_batch(config3(32))
for _ in range(2): eng.run(p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
tr = eng.fetch_trace()
names_e = {0: "start(st1)", 1: "next input done", 2: "d1 seen", 3: "st1 epi done", 4: "d2 seen", 5: "st2 epi done", 6: "d3 seen", 7: "tile done", 8: "LN pass1 done"}
for kname, t in (("enc_msg (layer 1)", tr[0]), ("edge_upd (layer 1)", tr[1])):
    print("====", kname, " (cycles; ~1.9 cycles/ns)")
    for ti in range(2, 6):
        base = t[ti, 0]
        if base == 0: continue
        e0 = {k: int(t[ti, k] - base) for k in range(9) if t[ti, k]}
        e1 = {k: int(t[ti, 16 + k] - base) for k in range(8) if t[ti, 16 + k]}
        mm = {k: int(t[ti, 24 + k] - base) for k in range(8) if t[ti, 24 + k]}
        nxt = int(t[ti + 1, 0] - base) if t[ti + 1, 0] else None
        print(f"tile#{ti}: next tile starts at {nxt}")
        print("   set0:", {names_e[k]: v for k, v in e0.items()})
        print("   set1:", {names_e[k]: v for k, v in e1.items()})
        print("   mma :", {("A0 st%d" % (k // 2 + 1)) if k % 2 == 0 else ("commit st%d" % (k // 2 + 1)): v for k, v in mm.items()})
