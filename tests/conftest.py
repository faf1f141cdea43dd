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


def c4_batch():
    """The 72 Megascale-shaped proteins of tests/golden/case_c4batch.npz (coordinates as stored) and the fixture itself."""
    d = np.load(os.path.join(GOLDEN, "case_c4batch.npz"))
    pdbs, o = [], 0
    for name, seq, L in zip(d["names"], d["seqs"], d["lengths"]):
        x = d["xyz"][o:o + L]
        pdbs.append({"name": str(name), "seq": str(seq), "num_of_chains": 1, "seq_chain_A": str(seq),
                     "coords_chain_A": {f"{a}_chain_A": x[:, i].tolist() for i, a in enumerate(("N", "CA", "C", "O"))}})
        o += int(L)
    return d, pdbs


def cfg_variant(tag, base_weights=None):
    """(cfg dict, weights, {protein: reference ddG}) of one non-default head in tests/golden/case_cfg_variants.npz.
    weights = the vanilla prot_mpnn.* tensors of the slim checkpoint + the variant's head tensors (numpy, or torch when
    ``base_weights`` -- the oracle's weight dict -- is given)."""
    d = np.load(os.path.join(GOLDEN, "case_cfg_variants.npz"))
    c = d[f"{tag}.cfg"]
    cfg = {"num_final_layers": int(c[0]), "lightattn": bool(c[1]), "hidden_dims": [int(x) for x in c[2:]]}
    head = {k[len(tag) + 3:]: d[k] for k in d.files if k.startswith(f"{tag}.w.")}
    if base_weights is not None:
        import torch
        w = {k: v for k, v in base_weights.items() if k.startswith("prot_mpnn.")}
        w.update({k: torch.from_numpy(v) for k, v in head.items()})
    else:
        w = {k: v for k, v in np.load(CKPT).items() if k.startswith("prot_mpnn.")}
        w.update(head)
    ddgs = {k[len(tag) + 5:]: d[k] for k in d.files if k.startswith(f"{tag}.ddG.")}
    return cfg, w, ddgs


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


@pytest.fixture(scope="session")
def gpu_baseline():
    """ProteinMPNNBaseline on the vanilla v_48_020 weights (= the prot_mpnn.* tensors of the slim checkpoint)."""
    import torch
    from thermompnn_b200.transfer_model import ProteinMPNNBaseline
    state = {k: v for k, v in np.load(CKPT).items() if k.startswith("prot_mpnn.")}
    return ProteinMPNNBaseline(state=state).eval().cuda()
