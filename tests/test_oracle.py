"""CPU: pin the oracle against outputs of the UNMODIFIED reference (tests/golden, made by
tools/make_golden.py) and against the reference's only shipped known-answer file."""
import csv
import os

import numpy as np
import pytest
import torch

from conftest import CASES, GOLDEN, load_case
from oracle import thermompnn_oracle as O


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name, oracle_weights):
    d, pdb = load_case(name)
    tr = {}
    T, ddg, S = O.ssm_table(oracle_weights, pdb, trace=tr)
    ref = d["ddG"]
    ok = ~np.isnan(ref)
    assert ok.sum() == 20 * sum(a != "-" for a in pdb["seq"])
    # neighbour lists: identical, order included
    assert np.array_equal(tr["E_idx"].numpy(), d["E_idx"])
    # hidden states: same ATen ops in the same order -> tight
    for l in range(3):
        assert np.abs(tr[f"enc{l}_h_V"].numpy() - d[f"enc{l}_h_V"]).max() <= 1e-6
        assert np.abs(tr[f"dec{l}_h_V"].numpy() - d[f"dec{l}_h_V"]).max() <= 1e-6
        assert np.abs(tr[f"enc{l}_h_E"].numpy()[:6] - d[f"enc{l}_h_E_head"]).max() <= 1e-5
    # ddG: the vectorised head vs the reference's per-mutation conv loop
    assert np.abs(ddg.numpy() - ref)[ok].max() <= 5e-6
    # wt -> wt is exactly zero (transfer_model.py:113-114)
    wt = S
    assert np.all(ddg.numpy()[np.arange(len(wt))[wt < 20], wt[wt < 20]] == 0.0)


def test_oracle_vs_published_csv(oracle_weights):
    """examples/ThermoMPNN_inference_2OCJ.csv was produced on the authors' GPU: it pins results to ~1e-3."""
    d, pdb = load_case("2OCJ_A")
    _, ddg, _ = O.ssm_table(oracle_weights, pdb)
    ddg = ddg.numpy()
    rows = list(csv.DictReader(open(os.path.join(GOLDEN, "ref_2OCJ_A_published.csv"))))
    assert len(rows) == 194 * 20
    err = 0.0
    for r in rows:
        pos, mut = int(r["position"]), r["mutation"]
        assert pdb["seq"][pos] == r["wildtype"]
        err = max(err, abs(ddg[pos, O.ALPHABET.index(mut)] - float(r["ddG_pred"])))
    assert err <= 1e-3


def test_per_mutation_head_equals_table(oracle_weights):
    """The reference's per-mutation LightAttention loop (timed by bench.py --impl reference) collapses to the
    centre-tap table: off-centre taps and the attention branch cannot change the result."""
    d, pdb = load_case("synth40")
    w = dict(oracle_weights)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(384, 384, 9, generator=g)
    full[:, :, 4] = w["light_attention.feature_convolution.weight_center"]
    w["light_attention.feature_convolution.weight"] = full
    w["light_attention.attention_convolution.weight"] = torch.randn(384, 384, 9, generator=g)
    w["light_attention.attention_convolution.bias"] = torch.randn(384, generator=g)

    class M:
        def __init__(s, p, a, b): s.position, s.wildtype, s.mutation = p, a, b
    muts = [M(3, pdb["seq"][3], "W"), None, M(17, pdb["seq"][17], "A")]
    out = O.forward_as_reference(w, pdb, muts)
    ref = d["ddG"]
    assert out[1] is None
    assert abs(out[0] - ref[3, O.ALPHABET.index("W")]) <= 5e-6
    assert abs(out[2] - ref[17, O.ALPHABET.index("A")]) <= 5e-6


def test_fp64_oracle_brackets_fp32_noise(oracle_weights):
    """The reference itself is ~2e-5 away from fp64 truth; this is the noise floor the 1e-4 bar sits on."""
    d, pdb = load_case("synth300")
    w64 = O.cast_weights(oracle_weights, torch.float64)
    _, ddg64, _ = O.ssm_table(w64, pdb, dtype=torch.float64)
    assert np.nanmax(np.abs(ddg64.numpy() - d["ddG"])) <= 5e-5


# ---- round-2 fixtures (tools/make_golden_r2.py, also made by the unmodified reference) -----------------------------
@pytest.mark.parametrize("name", ["synth600", "synth2000"])
def test_oracle_long_proteins(name, oracle_weights):
    """L > 512 (the CUDA k-NN kernel's shared-memory row path) and BASELINE configs[4]'s protein length."""
    d, pdb = load_case(name)
    tr = {}
    _, ddg, _ = O.ssm_table(oracle_weights, pdb, trace=tr)
    assert np.array_equal(tr["E_idx"].numpy(), d["E_idx"])
    assert np.abs(ddg.numpy() - d["ddG"]).max() <= 5e-6
    assert np.abs(tr["log_probs"].numpy() - d["log_probs"]).max() <= 2e-5


def test_oracle_log_probs(oracle_weights):
    """ProteinMPNN's own output layer (protein_mpnn_utils.py:1275-1277) on the round-1 cases."""
    lp = np.load(os.path.join(GOLDEN, "extra_logprobs.npz"))
    for name in lp.files:
        _, pdb = load_case(name)
        tr = {}
        O.ssm_table(oracle_weights, pdb, trace=tr)
        got = tr["log_probs"].numpy()
        assert np.abs(got - lp[name]).max() <= 2e-5, name
        assert np.abs(np.exp(got).sum(-1) - 1).max() <= 1e-5


def test_oracle_c4_batch(oracle_weights):
    """Megascale-shaped proteins (L 32..72, half of them K = L < 48; BASELINE configs[3])."""
    from conftest import c4_batch
    d, pdbs = c4_batch()
    o = 0
    for pdb in pdbs[::6]:
        i = [p["name"] for p in pdbs].index(pdb["name"])
        o = int(d["lengths"][:i].sum())
        _, ddg, _ = O.ssm_table(oracle_weights, pdb)
        assert np.abs(ddg.numpy() - d["ddG"][o:o + len(pdb["seq"])]).max() <= 5e-6, pdb["name"]


def test_oracle_cfg_variants(oracle_weights):
    """Other cfg.model heads (transfer_model.py:45-73): 3 final layers without LightAttention, 1 final layer with it, none."""
    from conftest import cfg_variant
    for tag in ("v1", "v2", "v3"):
        cfg, w, ddgs = cfg_variant(tag, oracle_weights)
        for pn, ref in ddgs.items():
            pdb = synth_by_name(pn)
            _, ddg, _ = O.ssm_table(w, pdb, num_final_layers=cfg["num_final_layers"], lightattn=cfg["lightattn"])
            assert np.abs(ddg.numpy() - ref).max() <= 5e-6, (tag, pn)


def synth_by_name(pn):
    from thermompnn_b200.synth import synth
    return {"synth90": synth(90, seed=90), "synth40": synth(40, seed=7)}[pn]


def test_oracle_unconditional_probs(oracle_weights):
    """ProteinMPNN.unconditional_probs (protein_mpnn_utils.py:1557-1587): fixtures from the live reference."""
    ref = np.load(os.path.join(GOLDEN, "extra_unconditional.npz"))
    for name in ref.files:
        _, pdb = load_case(name)
        X, S, mask, ridx, cenc = O.featurize(pdb)
        with torch.no_grad():
            got = O.unconditional_log_probs(oracle_weights, torch.from_numpy(X), torch.from_numpy(mask), torch.from_numpy(ridx),
                                            torch.from_numpy(cenc)).numpy()
        assert np.abs(got - ref[name]).max() <= 2e-5, name
