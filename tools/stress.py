#!/usr/bin/env python
"""Dev tool: repeat-run stress for races -- the same batch many times must give bit-identical tables, and batches of
random composition must equal single-protein runs bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from thermompnn_b200.synth import synth
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    n = int(rng.integers(1, 40))
    pdbs = [synth(int(rng.integers(1, 700)), int(rng.integers(0, 10**6))) for _ in range(n)]
    ref = model.head_tables(pdbs)
    for rep in range(3):
        again = model.head_tables(pdbs)
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, again))
    for i in rng.choice(n, size=min(n, 4), replace=False):
        bad += not np.array_equal(ref[i], model.head_tables([pdbs[i]])[0])
    assert all(np.isfinite(t).all() for t in ref)
    print(f"trial {trial}: {n} proteins, {sum(len(p['seq']) for p in pdbs)} residues, mismatches so far {bad}")
print("STRESS", "FAIL" if bad else "OK")
