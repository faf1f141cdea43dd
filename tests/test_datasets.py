"""CPU: dataset adapters (datasets.py upstream) and the dataset-level metric loops (thermompnn_benchmarking.py:68-187)."""
import os
import pickle

import numpy as np
import pytest
import torch

from thermompnn_b200 import benchmarking as B
from thermompnn_b200.datasets import (FireProtDataset, MegaScaleDataset, Mutation, ddgBenchDataset, global_align_xx,
                                      seq1_index_to_seq2_index)
from thermompnn_b200.synth import synth, write_pdb
from thermompnn_b200.transfer_model import AttrDict

REF = "/root/reference"


class FakeModel:
    """Stands in for the CUDA model: the forward / predict_batch contract with a deterministic 'prediction'."""

    def predict_batch(self, batch):
        return [[None if m is None else {"ddG": torch.tensor([0.1 * m.position - 0.05 * "ACDEFGHIKLMNPQRSTVWY".index(m.mutation)])}
                 for m in muts] for _, muts in batch]


def _mega_fixture(tmp_path, with_pdbs):
    import pandas as pd
    rows, names = [], ["p1.pdb", "p2|x.pdb"]
    seqs = {"p1.pdb": synth(12, 1)["seq"], "p2|x.pdb": synth(9, 2)["seq"]}
    for n in names:
        s = seqs[n]
        rows.append(dict(ddG_ML="0.0", mut_type="wt", WT_name=n, aa_seq=s, dG_ML="1.0"))
        for i, (pos, mut) in enumerate([(0, "W"), (3, "A"), (5, "G")]):
            mut = mut if s[pos] != mut else "C"
            rows.append(dict(ddG_ML="-" if i == 2 and n == "p1.pdb" else str(0.5 * i - 0.3), mut_type=f"{s[pos]}{pos + 1}{mut}",
                             WT_name=n, aa_seq=s[:pos] + mut + s[pos + 1:], dG_ML="1.0"))
        rows.append(dict(ddG_ML="0.2", mut_type="ins2", WT_name=n, aa_seq=s + "A", dG_ML="1.0"))
        rows.append(dict(ddG_ML="0.2", mut_type=f"{s[0]}1A:{s[1]}2C", WT_name=n, aa_seq=s, dG_ML="1.0"))
    csv = tmp_path / "mega.csv"
    pd.DataFrame(rows).to_csv(csv)
    spl = tmp_path / "splits.pkl"
    pickle.dump({"train": [names[0]], "val": [names[1]], "test": []}, open(spl, "wb"))
    pdb_dir = None
    if with_pdbs:
        pdb_dir = tmp_path / "pdbs"
        pdb_dir.mkdir()
        write_pdb(synth(12, 1), str(pdb_dir / "p1.pdb"))
        write_pdb(synth(9, 2), str(pdb_dir / "p2:x.pdb"))
        pdb_dir = str(pdb_dir)
    return AttrDict.wrap({"data_loc": {"megascale_csv": str(csv), "megascale_splits": str(spl), "megascale_pdbs": pdb_dir}}), seqs


@pytest.mark.parametrize("with_pdbs", [True, False])
def test_megascale_dataset(tmp_path, with_pdbs):
    cfg, seqs = _mega_fixture(tmp_path, with_pdbs)
    ds = MegaScaleDataset(cfg, "all")
    assert len(ds) == 2 and ds.wt_seqs["p2|x.pdb"] == seqs["p2|x.pdb"]
    pdb, muts = ds[0]
    assert isinstance(pdb, list) and pdb[0]["seq"] == seqs["p1.pdb"] and len(pdb[0]["coords_chain_A"]["CA_chain_A"]) == 12
    assert [m.position for m in muts] == [0, 3]                       # '-' row, ins row and double mutant dropped (datasets.py:70-72)
    assert muts[1].wildtype == seqs["p1.pdb"][3] and muts[1].pdb == "p1"
    assert torch.equal(muts[0].ddG, torch.tensor([0.3])) and muts[0].ddG.dtype == torch.float32     # sign flipped (:161)
    pdb2, muts2 = ds[1]
    assert muts2[0].pdb == "p2:x" and len(muts2) == 3
    if with_pdbs:                                                    # parsed coordinates == the synthetic ones written
        assert np.allclose(pdb[0]["coords_chain_A"]["CA_chain_A"], synth(12, 1)["coords_chain_A"]["CA_chain_A"], atol=1e-3)
    assert len(MegaScaleDataset(cfg, "val")) == 1


def test_ddgbench_dataset(tmp_path):
    import pandas as pd
    p = synth(15, 3, gaps=[4])
    write_pdb(p, str(tmp_path / "1abc.pdb"), first_resnum=10)
    seq = p["seq"]
    rows = [dict(PDB="1abcA", MUT=f"{seq[0]}10W", DDG=1.5, SEQ=seq.replace("-", "")),
            dict(PDB="1abcA", MUT=f"{seq[7]}17G", DDG=float("nan"), SEQ=seq.replace("-", "")),
            dict(PDB="1abcA", MUT=f"{seq[2]}12AA", DDG=0.5, SEQ=seq.replace("-", ""))]          # '12A' not a residue label: skipped
    csv = tmp_path / "bench.csv"
    pd.DataFrame(rows).to_csv(csv)
    ds = ddgBenchDataset(AttrDict(), str(tmp_path), str(csv))
    pdb, muts = ds[0]
    assert pdb[0]["seq"] == seq and len(muts) == 2
    assert muts[0].position == 0 and torch.equal(muts[0].ddG, torch.tensor([-1.5])) and muts[0].pdb == "1abc"
    assert muts[1].position == 7 and muts[1].ddG is None           # residue 17 sits after the gap: index 7 of the gapped sequence


def test_fireprot_dataset_and_alignment(tmp_path):
    import pandas as pd
    p = synth(14, 5)
    seq = p["seq"]
    write_pdb(p, str(tmp_path / "2xyz.pdb"))
    longer = "MK" + seq                                              # the CSV's sequence has two extra leading residues
    rows = [dict(pdb_id_corrected="2xyz", pdb_sequence=seq, pdb_position=3, wild_type=seq[3], mutation="W", ddG=1.0),
            dict(pdb_id_corrected="2xyz", pdb_sequence=seq, pdb_position=6, wild_type=seq[6], mutation="A", ddG=None)]
    csv = tmp_path / "fp.csv"
    pd.DataFrame(rows).to_csv(csv)
    pickle.dump({"test": ["2xyz"]}, open(tmp_path / "fp.pkl", "wb"))
    cfg = AttrDict.wrap({"data_loc": {"fireprot_csv": str(csv), "fireprot_splits": str(tmp_path / "fp.pkl"), "fireprot_pdbs": str(tmp_path)}})
    pdb, muts = FireProtDataset(cfg, "test")[0]
    assert len(muts) == 1 and muts[0].position == 3 and torch.equal(muts[0].ddG, torch.tensor([1.0]))   # NaN ddG rows dropped (:176)
    a, b = global_align_xx(longer, seq)
    assert a.replace("-", "") == longer and b.replace("-", "") == seq and len(a) == len(b)
    assert sum(x == y for x, y in zip(a, b)) == len(seq)             # optimal score
    assert seq1_index_to_seq2_index(a, b, 5) == 3 and seq1_index_to_seq2_index(a, b, 0) is None


def test_metrics_and_loops(tmp_path):
    from scipy import stats
    rng = np.random.default_rng(0)
    p, t = rng.normal(size=200), rng.normal(size=200)
    p[:10] = p[10:20]                                                # ties
    m = B.regression_metrics(p, t)
    assert abs(m["pearson"] - stats.pearsonr(p, t)[0]) < 1e-12 and abs(m["spearman"] - stats.spearmanr(p, t)[0]) < 1e-12
    assert abs(m["mse"] - ((p - t) ** 2).mean()) < 1e-12 and abs(m["rmse"] ** 2 - m["mse"]) < 1e-12
    assert abs(m["r2"] - (1 - ((p - t) ** 2).sum() / ((t - t.mean()) ** 2).sum())) < 1e-12
    cfg, seqs = _mega_fixture(tmp_path, False)
    ds = MegaScaleDataset(cfg, "all")
    res = B.run_prediction_default("M", FakeModel(), "Megascale-x", ds, [], proteins_per_launch=1)
    res2 = B.run_prediction_default("M", FakeModel(), "Megascale-x", ds, [], proteins_per_launch=64)
    assert res == res2 and set(res[0]) == {"Model", "Dataset", "ddG r2", "ddG mse", "ddG rmse", "ddG spearman", "ddG pearson"}
    out = tmp_path / "raw.csv"
    B.run_prediction_keep_preds("M", FakeModel(), "Megascale-x", ds, [], centrality=True, out_csv=str(out))
    import pandas as pd
    df = pd.read_csv(out, index_col=0)
    assert list(df.columns) == ["WT Seq", "Model", "Dataset", "ddG_true", "ddG_pred", "position", "wildtype", "mutation", "neighbors",
                                "best_AA", "pdb"] and len(df) == 5
    assert df["WT Seq"][0] == seqs["p1.pdb"] and df["pdb"][4] == "2:x" and df["neighbors"].notna().all()   # (sic) upstream strips CHARACTERS '.pdb' (:149)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "data_all", "testing", "mega_test.csv")), reason="needs the reference checkout")
def test_megascale_on_the_reference_csv():
    """The real mega_test.csv / mega_splits.pkl through the adapter (synthetic coordinates: the PDB files are not shipped):
    28 proteins, 28 172 listed mutations -- the numbers SURVEY.md 8(d) quotes for this split."""
    cfg = AttrDict.wrap({"data_loc": {"megascale_csv": os.path.join(REF, "data_all/testing/mega_test.csv"),
                                      "megascale_splits": os.path.join(REF, "dataset_splits/mega_splits.pkl"), "megascale_pdbs": None}})
    ds = MegaScaleDataset(cfg, "test")
    assert len(ds) == 28
    n = 0
    for pdb, muts in ds:
        assert all(pdb[0]["seq"][m.position] == m.wildtype for m in muts)
        n += len(muts)
    assert n == 28172
    shape = np.load(os.path.join(os.path.dirname(__file__), "golden", "megascale_shape.npz"))
    assert len(shape["names"]) == 298 and len(shape["protein"]) == 271231


def test_run_ssm_frame(tmp_path):
    """Dataset-wide site saturation (analysis/SSM.py:93-164): frame layout, row order, pick_best / include_cys filtering."""
    from thermompnn_b200.ssm import run_ssm

    class FakeSSM:
        def ssm_tables(self, pdbs):
            out = []
            for p in pdbs:
                L = len(p["seq"])
                t = np.add.outer(np.arange(L, dtype=np.float32), -0.01 * np.arange(20, dtype=np.float32))
                t[[i for i, a in enumerate(p["seq"]) if a == "-"]] = np.nan
                out.append(t)
            return out
    cfg, seqs = _mega_fixture(tmp_path, False)
    ds = MegaScaleDataset(cfg, "all")
    df = run_ssm("M", FakeSSM(), "Megascale-x", ds, centrality=True, include_cys=True)
    assert list(df.columns) == ["WT Seq", "Model", "Dataset", "ddG_pred", "position", "wildtype", "mutation", "neighbors", "best_AA", "pdb"]
    assert len(df) == 20 * (12 + 9) and df["mutation"][:20].tolist() == list("ACDEFGHIKLMNPQRSTVWY")
    assert df["position"][20] == 1 and df["wildtype"][0] == seqs["p1.pdb"][0] and df["WT Seq"][0] == seqs["p1.pdb"]
    assert abs(df["ddG_pred"][25] - (1 - 0.05)) < 1e-6 and df["neighbors"].notna().all()
    no_cys = run_ssm("M", FakeSSM(), "Megascale-x", ds)
    assert len(no_cys) == 19 * 21 and (no_cys["mutation"] != "C").all()
    best = run_ssm("M", FakeSSM(), "Megascale-x", ds, pick_best=True, include_cys=True)
    assert len(best) == 21 and (best["best_AA"] == "Y").all()      # column 19 (Y) has the lowest value in every row of the fake table
    best_nc = run_ssm("M", FakeSSM(), "Megascale-x", ds, pick_best=True, include_cys=False)
    assert (best_nc["best_AA"] == "Y").all()
