"""CPU: host-side logic and the C-ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import CKPT, GOLDEN, ROOT, load_case
from thermompnn_b200 import engine
from thermompnn_b200.datasets import ALPHABET, Mutation
from thermompnn_b200.featurize import featurize_batch, featurize_one
from thermompnn_b200.ssm import get_ssm_mutations, to_mutation_objects
from thermompnn_b200.synth import synth, synth_multichain
from thermompnn_b200.weights import load_state


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "thermompnn_b200.h")).read()
    declared = set(re.findall(r"\b(tmpnn_[a-z_0-9]+)\s*\(", hdr))
    declared.discard("tmpnn_ctx")
    lib = engine.load_library()
    assert declared == set(engine.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.tmpnn_abi_version() == 1


def test_weight_contract_matches_reference_checkpoint():
    state = load_state(CKPT)
    table = engine.weight_table()
    names = {n for n, *_ in table}
    assert names == set(state.keys())
    off = 0
    for n, o, r, c in table:
        assert o == off and state[n].size == r * c
        off += r * c
    assert off == engine.load_library().tmpnn_weight_blob_floats() == 1835740
    blob = engine.pack_weight_blob(state)
    n, o, r, c = table[5]
    assert np.array_equal(blob[o:o + r * c].reshape(r, c), state[n].reshape(r, c))


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError, match="no CUDA device|CPU path"):
        engine.Engine(0)
    from thermompnn_b200.transfer_model import get_trained_model
    m = get_trained_model(CKPT, override_custom=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m([synth(20, 1)], [Mutation(0, "A", "C")])


def test_featurize_matches_oracle_and_reference_layout():
    from oracle import thermompnn_oracle as O
    for pdb in (synth(30, 3, gaps=[2, 9]), synth_multichain([12, 9, 5], seed=2), load_case("2OCJ_AB")[1]):
        X, S, r, c, m = featurize_one(pdb)
        Xo, So, mo, ro, co = O.featurize(pdb)
        assert np.array_equal(X, Xo) and np.array_equal(S, So) and np.array_equal(r, ro)
        assert np.array_equal(c, co) and np.array_equal(m, mo.astype(np.uint8))
    pdb = load_case("2OCJ_AB")[1]
    _, _, r, c, _ = featurize_one(pdb)
    assert r[193] == 193 and r[194] == 294 and c[193] == 1 and c[194] == 2   # +100 per chain break


def test_featurize_batch_is_concatenation():
    pdbs = [synth(17, 1), synth(33, 2, gaps=[4]), synth_multichain([8, 7], 3)]
    p = featurize_batch(pdbs)
    assert p.offsets.tolist() == [0, 17, 50, 65] and p.n_proteins == 3 and p.n_residues == 65
    X1, S1, r1, c1, m1 = featurize_one(pdbs[1])
    assert np.array_equal(p.X[17:50], X1) and np.array_equal(p.S[17:50], S1) and np.array_equal(p.mask[17:50], m1)
    assert p.mask[17 + 4] == 0 and p.S[17 + 4] == 20
    with pytest.raises(ValueError, match="not in list"):
        bad = synth(5, 1)
        bad["seq_chain_A"] = "AB" + bad["seq_chain_A"][2:]
        featurize_batch([bad])


def test_ssm_enumeration():
    pdb = synth(6, 1, gaps=[2])
    muts = get_ssm_mutations(pdb)
    assert len(muts) == 5 * 20 + 1 and muts[40] is None
    assert muts[0] == pdb["seq"][0] + "0A" and muts[19] == pdb["seq"][0] + "0Y" and muts[41].startswith(pdb["seq"][3] + "3")
    objs = to_mutation_objects(muts, "x")
    assert objs[40] is None and objs[41].position == 3 and objs[41].mutation == "A" and objs[41].wildtype == pdb["seq"][3]
    assert len(ALPHABET) == 21


def test_synth_is_seeded_and_pdb_like():
    a, b = synth(50, 9), synth(50, 9)
    assert a == b and a != synth(50, 10)
    ca = np.asarray(a["coords_chain_A"]["CA_chain_A"])
    assert np.allclose(np.linalg.norm(np.diff(ca, axis=0), axis=1), 3.8, atol=2e-3)
    assert np.array_equal(ca, np.round(ca, 3))


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract: exactly one line on stdout, a JSON object with the driver's keys (the CPU arm runs anywhere)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
              "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
