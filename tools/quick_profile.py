#!/usr/bin/env python
"""Per-kernel CUDA-event timings for a workload (dev tool; bench.py is the graded harness)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from thermompnn_b200.featurize import featurize_batch
from thermompnn_b200.synth import synth, config3
from thermompnn_b200.transfer_model import get_trained_model

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
model = get_trained_model(os.path.join(ROOT, "tests/golden/thermompnn_default_slim.npz"), override_custom=True).eval().cuda()
eng = model._get_engine()
if a.workload == "c2":
    pdbs = [synth(300, 0)]
elif a.workload.startswith("c3"):
    n = int(a.workload[3:] or 256) if len(a.workload) > 2 else 256
    pdbs = config3(n)
elif a.workload.startswith("L"):
    pdbs = [synth(int(a.workload[1:]), 1)]
p = featurize_batch(pdbs)
args = (p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
for _ in range(3): eng.run(*args)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.iters): eng.run(*args)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.iters
nm = 20 * p.n_residues
print(f"{a.workload}: N={p.n_residues} B={p.n_proteins}  e2e {dt*1e3:.3f} ms/step  {nm/dt:.3e} mut/s")
eng.set_profiling(True); eng.profile_reset()
for _ in range(a.iters): eng.run(*args)
prof = eng.profile()
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:12s} {ms/a.iters*1e3:9.1f} us/step  {n//a.iters:3d} launches  {100*ms/tot:5.1f}%")
print(f"  sum kernels {tot/a.iters*1e3:.1f} us/step")
