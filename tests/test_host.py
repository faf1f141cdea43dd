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
    assert lib.tmpnn_abi_version() == 2


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


REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models", "thermoMPNN_default.pt")), reason="needs the reference checkout")


def test_cfg_weight_contracts():
    """cfg.model variants (transfer_model.py:58-73): the weight table follows hid_sizes = [128 F + 128] + hidden_dims + [21]."""
    lib = engine.load_library()
    t = engine.weight_table(engine.ModelCfg.make(3, False, (48,), True))
    names = [n for n, *_ in t]
    assert "light_attention.feature_convolution.weight_center" not in names
    shapes = {n: (r, c) for n, _, r, c in t}
    assert shapes["both_out.1.weight"] == (48, 512) and shapes["both_out.3.weight"] == (21, 48) and "both_out.5.weight" not in shapes
    t2 = {n: (r, c) for n, _, r, c in engine.weight_table(engine.ModelCfg.make(1, True, (32, 16, 8), True))}
    assert t2["light_attention.feature_convolution.weight_center"] == (256, 256) and t2["both_out.7.weight"] == (21, 8)
    t3 = [n for n, *_ in engine.weight_table(engine.ModelCfg.make(has_head=False))]
    assert t3[-1] == "prot_mpnn.W_out.bias" and all(n.startswith("prot_mpnn.") for n in t3)
    bad = engine.ModelCfg.make(4, True, (64,), True)
    assert lib.tmpnn_cfg_weight_count(ctypes.byref(bad)) < 0
    with pytest.raises(ValueError):
        engine.weight_table(bad)
    # the default-cfg forms are the cfg forms at the default cfg
    d = engine.ModelCfg()
    lib.tmpnn_default_cfg(ctypes.byref(d))
    assert d.key() == engine.ModelCfg.make().key() == (2, 1, (64, 32), 1)
    assert lib.tmpnn_cfg_weight_blob_floats(ctypes.byref(d)) == lib.tmpnn_weight_blob_floats()


def test_cfg_variant_fixture_packs():
    from conftest import cfg_variant
    from thermompnn_b200.transfer_model import AttrDict, TransferModel, default_config
    for tag in ("v1", "v2", "v3"):
        cfgd, w, _ = cfg_variant(tag)
        cfg = default_config()
        cfg.model.update(cfgd)
        m = TransferModel(cfg, state=w)
        assert m.weight_blob.numel() == engine.load_library().tmpnn_cfg_weight_blob_floats(ctypes.byref(m._mcfg))


@needs_ref
def test_get_trained_model_paths_like_upstream():
    """thermompnn_benchmarking.py:78-84: <thermompnn_dir>/<checkpt_dir>/<model_name> unless override_custom; the stock
    Lightning .pt loads and equals the slim fixture tensor for tensor."""
    from thermompnn_b200.transfer_model import default_config, get_trained_model
    cfg = default_config(REF)
    a = get_trained_model("thermoMPNN_default.pt", cfg)                                  # upstream's default call
    b = get_trained_model(os.path.join(REF, "models", "thermoMPNN_default.pt"), cfg, override_custom=True)
    c = get_trained_model(CKPT, override_custom=True)
    assert (a.weight_blob == b.weight_blob).all() and (a.weight_blob == c.weight_blob).all()
    with pytest.raises(FileNotFoundError):
        get_trained_model("thermoMPNN_default.pt", default_config("/nonexistent"))
    state = load_state(os.path.join(REF, "models", "thermoMPNN_default.pt"))
    slim = load_state(CKPT)
    assert set(state) == set(slim) and all(np.array_equal(state[k].reshape(-1), slim[k].reshape(-1)) for k in slim)


@needs_ref
def test_constructors_follow_get_protein_mpnn():
    """transfer_model.py:17-37: the vanilla checkpoint under <thermompnn_dir>/vanilla_model_weights is required
    (no silent random weights), any v_48_* version loads into ProteinMPNNBaseline."""
    import torch
    from thermompnn_b200.transfer_model import ProteinMPNNBaseline, TransferModel, default_config
    with pytest.raises(FileNotFoundError):
        TransferModel(default_config("/nonexistent"))
    m = TransferModel(default_config(REF))
    van = torch.load(os.path.join(REF, "vanilla_model_weights", "v_48_020.pt"), map_location="cpu", weights_only=True)["model_state_dict"]
    tab = {n: (o, r * c) for n, o, r, c in engine.weight_table()}
    o, n = tab["prot_mpnn.W_e.weight"]
    assert np.array_equal(m.weight_blob.numpy()[o:o + n], van["W_e.weight"].numpy().reshape(-1))
    for version in ("v_48_002.pt", "v_48_010.pt", "v_48_030.pt"):
        b = ProteinMPNNBaseline(default_config(REF), version=version)
        van = torch.load(os.path.join(REF, "vanilla_model_weights", version), map_location="cpu", weights_only=True)["model_state_dict"]
        tab = {n: (o, r * c) for n, o, r, c in engine.weight_table(b._mcfg)}
        o, n = tab["prot_mpnn.W_out.weight"]
        assert np.array_equal(b.weight_blob.numpy()[o:o + n], van["W_out.weight"].numpy().reshape(-1))


def test_positions_follow_tensor_indexing():
    from thermompnn_b200.transfer_model import TransferModel
    muts = [Mutation(-1, "A", "C"), None, Mutation(3, "A", "C"), Mutation(-10, "A", "C")]
    idx, pos = TransferModel._positions(muts, 10)
    assert idx == [0, 2, 3] and pos.tolist() == [9, 3, 0]
    for bad in (10, -11):
        with pytest.raises(IndexError, match="out of bounds"):
            TransferModel._positions([Mutation(bad, "A", "C")], 10)


def test_engine_run_validates_host_buffers():
    """Arguments that reach C as raw pointers are checked first (no GPU needed: the checks come before the call)."""
    class Dummy(engine.Engine):
        def __init__(self): self._lib = None; self._ctx = None
        def __del__(self): pass
    e = Dummy()
    p = featurize_batch([synth(10, 1)])
    args = (p.offsets, p.X, p.S, p.residue_idx, p.chain_enc, p.mask)
    with pytest.raises(ValueError, match="float32"):
        e.run(*args, out=np.empty((10, 21), np.float64))
    with pytest.raises(ValueError, match="10 x 21"):
        e.run(*args, out=np.empty((9, 21), np.float32))
    with pytest.raises(ValueError, match="C-contiguous"):
        e.run(*args, out=np.empty((21, 10), np.float32).T)
    with pytest.raises(ValueError, match="log_probs"):
        e.run(*args, log_probs=np.empty((10, 20), np.float32))
    with pytest.raises(ValueError, match="raw device pointers"):
        e.run(*args, out=123, flags=engine.IN_DEVICE | engine.OUT_DEVICE)
    with pytest.raises(ValueError, match="nothing to compute"):
        e.run(*args, want_table=False)
    with pytest.raises(ValueError, match="sizes do not match"):
        e.run(p.offsets, p.X[:5], p.S, p.residue_idx, p.chain_enc, p.mask)


def test_predict_batch_lookup_logic_on_a_fake_table(monkeypatch):
    """forward / predict_batch around the device call: list alignment, None entries, wt subtraction, negative positions,
    one tensor shared by all results (the device pass itself is replaced by a fake table; CPU only)."""
    import torch
    from thermompnn_b200.transfer_model import TransferModel, get_trained_model
    m = get_trained_model(CKPT, override_custom=True)
    pdbs = [synth(12, 1), synth(9, 2)]
    tabs = [np.arange(12 * 21, dtype=np.float32).reshape(12, 21), -np.arange(9 * 21, dtype=np.float32).reshape(9, 21)]
    monkeypatch.setattr(TransferModel, "head_tables", lambda self, p: tabs[:len(p)])
    muts0 = [Mutation(0, pdbs[0]["seq"][0], "C"), None, Mutation(-1, pdbs[0]["seq"][-1], "Y")]
    muts1 = [Mutation(4, pdbs[1]["seq"][4], "A")]
    outs = m.predict_batch([([pdbs[0]], muts0), (pdbs[1], muts1)])
    assert outs[0][1] is None and len(outs[0]) == 3 and len(outs[1]) == 1
    A = ALPHABET.index
    assert outs[0][0]["ddG"].item() == tabs[0][0, A("C")] - tabs[0][0, A(pdbs[0]["seq"][0])]
    assert outs[0][2]["ddG"].item() == tabs[0][11, A("Y")] - tabs[0][11, A(pdbs[0]["seq"][-1])]
    assert outs[1][0]["ddG"].item() == tabs[1][4, A("A")] - tabs[1][4, A(pdbs[1]["seq"][4])]
    assert outs[0][0]["ddG"].shape == (1,) and outs[0][0]["ddG"].dtype == torch.float32
    assert outs[0][0]["ddG"].untyped_storage().data_ptr() == outs[1][0]["ddG"].untyped_storage().data_ptr()
    pred, extra = m([pdbs[0]], muts0)
    assert extra is None and pred[1] is None and torch.equal(pred[0]["ddG"], outs[0][0]["ddG"])
    assert m.predict_batch([]) == [] and m([pdbs[0]], [None])[0] == [None]
    with pytest.raises(ValueError):
        m([pdbs[0]], [Mutation(0, "A", "B")])


def test_array_coordinates_pack_identically():
    """``arrays=True`` (numpy coordinates instead of upstream's nested lists) changes the container, not a single value."""
    from thermompnn_b200.pdb import alt_parse_PDB
    a, b = [synth(50, 3, gaps=[4, 9]), synth(33, 4)], [synth(50, 3, gaps=[4, 9], arrays=True), synth(33, 4, arrays=True)]
    pa, pb = featurize_batch(a), featurize_batch(b)
    for k in ("offsets", "X", "S", "residue_idx", "chain_enc", "mask"):
        assert np.array_equal(getattr(pa, k), getattr(pb, k)), k
    pdb_path = os.path.join(GOLDEN, "2OCJ.pdb")
    x, y = alt_parse_PDB(pdb_path, "AB"), alt_parse_PDB(pdb_path, "AB", arrays=True)
    assert isinstance(y[0]["coords_chain_B"]["CA_chain_B"], np.ndarray) and x[0]["seq"] == y[0]["seq"]
    fx, fy = featurize_batch(x), featurize_batch(y)
    assert np.array_equal(fx.X, fy.X) and np.array_equal(fx.mask, fy.mask)
    with pytest.raises(ValueError, match="do not match sequence length"):
        bad = synth(10, 1)
        bad["coords_chain_A"]["CA_chain_A"] = bad["coords_chain_A"]["CA_chain_A"][:9]
        featurize_batch([bad])
