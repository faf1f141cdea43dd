#!/usr/bin/env python
"""Dev tool: boundary and repeat-run stress.  (i) batches whose total size sits on the kernel-form boundaries (the
warp-level / tcgen05 switch at 593 residues, tile tails of the 128-row tcgen05 tiles and the 16-residue cluster tiles,
one edge tile per SM at 394 residues): every protein against the oracle; (ii) the same batch many times must give
bit-identical tables; (iii) batches of random composition against single-protein runs (fp32 rounding across regimes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from oracle import thermompnn_oracle as O
from thermompnn_b200.synth import synth
from thermompnn_b200.transfer_model import get_trained_model
ckpt = os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz")
model = get_trained_model(ckpt, override_custom=True).eval().cuda()
w = O.cast_weights(dict(np.load(ckpt)))
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad, worst = 0, 0.0
# (i) boundaries: total residues N built from a few proteins
for N in (15, 16, 17, 127, 128, 129, 255, 256, 257, 393, 394, 395, 591, 592, 593, 594, 607, 608, 609, 640, 1023, 1024, 1025):
    parts, left, k = [], N, 0
    while left > 0:
        L = min(left, int(rng.integers(1, 200)))
        parts.append(synth(L, 1000 * N + k)); left -= L; k += 1
    tables = model.ssm_tables(parts)
    err = 0.0
    for pdb, ddg in zip(parts[:3] + parts[-2:], tables[:3] + tables[-2:]):
        _, ref, _ = O.ssm_table(w, pdb)
        err = max(err, float(np.abs(ddg - ref.numpy()).max()))
    worst = max(worst, err)
    bad += err > 1e-4
    print(f"N={N:5d} ({len(parts)} proteins): max |ddG - oracle| = {err:.2e}")
# (ii), (iii)
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    n = int(rng.integers(1, 40))
    pdbs = [synth(int(rng.integers(1, 700)), int(rng.integers(0, 10**6))) for _ in range(n)]
    ref = model.head_tables(pdbs)
    for rep in range(3):
        again = model.head_tables(pdbs)
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, again))
    for i in rng.choice(n, size=min(n, 4), replace=False):
        d = float(np.abs(ref[i] - model.head_tables([pdbs[i]])[0]).max())
        worst = max(worst, d)
        bad += d > 5e-5
    assert all(np.isfinite(t).all() for t in ref)
    print(f"trial {trial}: {n} proteins, {sum(len(p['seq']) for p in pdbs)} residues, failures so far {bad}")
print("STRESS", "FAIL" if bad else "OK", f"worst deviation {worst:.2e}")
