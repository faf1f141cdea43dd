#!/usr/bin/env python
"""Dev tool: stage-by-stage difference between two kernel generations (TMPNN_TC masks) on one fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.transfer_model import get_trained_model
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case

ma, mb = (sys.argv[1:3] + ["7", "3"])[:2] if len(sys.argv) >= 3 else ("7", "3")
name = sys.argv[3] if len(sys.argv) > 3 else "2OCJ_A"
d, pdb = load_case(name)
p = featurize_batch([pdb])
W = os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz")
engs = []
for m in (ma, mb):
    os.environ["TMPNN_TC"] = m
    model = get_trained_model(W, override_custom=True).eval().cuda()
    engs.append(model._get_engine()); engs[-1]._keep = model
for stage in range(2, 9):
    out = []
    for e in engs:
        e.set_debug_stop(stage)
        e.run(p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
        out.append((e.fetch("h_E").copy(), e.fetch("h_V").copy()))
    dE = np.abs(out[0][0] - out[1][0]).reshape(-1, 128).max(axis=1)       # per edge row
    dV = np.abs(out[0][1] - out[1][1]).max(axis=1)
    bad = np.nonzero(dE > 1e-3)[0]
    print(f"stage {stage}: h_E max diff {dE.max():.2e} ({len(bad)} bad rows of {len(dE)})   h_V max diff {dV.max():.2e} (bad residues {np.nonzero(dV > 1e-3)[0][:12]})")
    if len(bad):
        print("   bad rows (row, tile, row-in-tile):", [(int(r), int(r) // 128, int(r) % 128) for r in bad[:24]])
        r = bad[0]
        dc = np.abs(out[0][0] - out[1][0]).reshape(-1, 128)[r]
        print("   first bad row, bad columns:", np.nonzero(dc > 1e-3)[0][:40])
os._exit(0)   # two engines in one process: skip interpreter teardown
