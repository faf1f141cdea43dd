#!/usr/bin/env python
"""Dev tool: timeline of one tile of the tcgen05 node_update kernel (encoder layer 1, block 0), C4-sized batch."""
import os, sys
os.environ["TMPNN_TC_TRACE"] = "1"; os.environ["TMPNN_LIB"] = "dev"   # needs the -DTMPNN_DEV build (python __graft_entry__.py --dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.synth import synth
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
eng = model._get_engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
p = featurize_batch([synth(300, i) for i in range(n)])
epi = ["tile start", "x published", "D0(W3) ready", "h1 published"]
for q in range(4):
    epi += [f"D1(in{q}) ready", f"g{q} published", f"D0(out{q}) ready", f"h1 republished {q}"]
epi += ["LN2 done, h2 published"] + [x for p_ in range(4) for x in (f"proj{p_} ready", f"proj{p_} stored")]
for it in range(3):
    eng.run(p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
    t = eng.fetch_node_trace()
    t0 = t[0]
    print("run", it, "epilogue (warp 0):", " | ".join(f"{nm}={int(t[i] - t0)}" for i, nm in enumerate(epi) if t[i]))
    print("      mma warp: " + " ".join(f"g{g}:[{int(t[32 + 2 * g] - t0)},{int(t[33 + 2 * g] - t0)}]" for g in range(13) if t[32 + 2 * g]))
