import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
CKPT = os.path.join(GOLDEN, "thermompnn_default_slim.npz")
CASES = ["2OCJ_A", "2OCJ_AB", "synth300", "synth40", "synth120_gaps", "synth_multichain", "synth49"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def case_pdb(d) -> dict:
    """Rebuild the alt_parse_PDB-style dict a golden case was generated from."""
    pdb = {"name": str(d["name"]), "seq": str(d["seq"]), "num_of_chains": len(d["chains"])}
    for ch in d["chains"]:
        ch = str(ch)
        pdb["seq_chain_" + ch] = str(d["seq_" + ch])
        x = d["xyz_" + ch]
        pdb["coords_chain_" + ch] = {f"{a}_chain_{ch}": x[:, i].tolist() for i, a in enumerate(("N", "CA", "C", "O"))}
    return pdb


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    return d, case_pdb(d)


@pytest.fixture(scope="session")
def oracle_weights():
    from oracle import thermompnn_oracle as O
    return O.cast_weights(dict(np.load(CKPT)))


@pytest.fixture(scope="session")
def gpu_model():
    import torch
    from thermompnn_b200.transfer_model import get_trained_model
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return get_trained_model(CKPT, override_custom=True).eval().cuda()
