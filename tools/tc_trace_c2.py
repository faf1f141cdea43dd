#!/usr/bin/env python
"""Dev tool: single-tile timeline of the tensor-core edge kernels on workload C2 (one 300-residue protein), block 0."""
import os, sys
os.environ["TMPNN_TC_TRACE"] = "1"; os.environ["TMPNN_LIB"] = "dev"   # needs the -DTMPNN_DEV build (python __graft_entry__.py --dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.synth import synth
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
eng = model._get_engine()
p = featurize_batch([synth(300, 0)])
for _ in range(3): eng.run(p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
tr = eng.fetch_trace()
names = {13: "entry", 14: "setup done", 15: "upstream complete", 0: "input converted (st1 wait starts)", 2: "d1 seen", 3: "st1 epi done",
         4: "d2 seen", 5: "st2 epi done", 6: "d3 seen", 8: "LN pass1 done", 7: "tile done"}
for kname, t in (("enc_msg (layer 1)", tr[0]), ("edge_upd (layer 1)", tr[1])):
    base = t[0, 13]
    ev = sorted((int(t[0, k] - base), names[k]) for k in names if t[0, k])
    print("====", kname, " cycles from kernel entry (~1.9 cycles/ns):")
    print("   ", "  ".join(f"{n}={c}" for c, n in ev))
    print("    mma:", {("A0 st%d" % (k // 2 + 1)) if k % 2 == 0 else ("commit st%d" % (k // 2 + 1)): int(t[0, 24 + k] - base) for k in range(6) if t[0, 24 + k]})
