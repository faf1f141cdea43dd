#!/usr/bin/env python
"""Print max |ddG_gpu - ddG_reference| for every golden case, and vs the fp64 oracle (dev tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import CASES, CKPT, load_case
from oracle import thermompnn_oracle as O
from thermompnn_b200.transfer_model import get_trained_model
model = get_trained_model(CKPT, override_custom=True).eval().cuda()
w64 = O.cast_weights(dict(np.load(CKPT)), torch.float64)
worst = 0
for name in CASES:
    d, pdb = load_case(name)
    ddg = model.ssm_tables([pdb])[0]
    ref = d["ddG"]; ok = ~np.isnan(ref)
    t64 = O.ssm_table(w64, pdb, dtype=torch.float64)[1].numpy()
    e_ref = np.abs(ddg - ref)[ok].max(); e64 = np.abs(ddg - t64)[ok].max(); r64 = np.abs(ref - t64)[ok].max()
    worst = max(worst, e_ref)
    print(f"{name:18s} gpu-vs-ref {e_ref:.2e}   gpu-vs-fp64 {e64:.2e}   ref-vs-fp64 {r64:.2e}   rms gpu-vs-ref {np.sqrt(np.mean((ddg-ref)[ok]**2)):.2e}")
print(f"worst gpu-vs-ref {worst:.2e}")
