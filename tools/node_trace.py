#!/usr/bin/env python
"""Dev tool: phase timeline of the node_update cluster kernel (encoder layer 1, cluster 0 / CTA 0), workload C2."""
import os, sys
os.environ["TMPNN_TC_TRACE"] = "1"; os.environ["TMPNN_LIB"] = "dev"   # needs the -DTMPNN_DEV build (python __graft_entry__.py --dev)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.synth import synth
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
eng = model._get_engine()
p = featurize_batch([synth(300, 0)])
names = ["start", "ring issued", "pdl_wait done", "inputs loaded", "W3 fma done", "allreduce 1 done", "LN1 done", "FFN-in done",
         "FFN-out fma done", "allreduce 2 done", "LN2 done", "proj done", "exit sync done"]
for it in range(3):
    eng.run(p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
    t = eng.fetch_node_trace()
    print("run", it, " ".join(f"{n}={int(t[i] - t[0])}" for i, n in enumerate(names)), f"| AR1: before cluster.sync={int(t[13] - t[0])} after={int(t[14] - t[0])}")
