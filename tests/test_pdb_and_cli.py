"""CPU: PDB parser against the reference's own parse of its example structure; CSV writer shape."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from thermompnn_b200.pdb import alt_parse_PDB


@pytest.mark.parametrize("chains", ["A", "AB"])
def test_parser_matches_reference(chains):
    ref = np.load(os.path.join(GOLDEN, f"parse_2OCJ_{chains}.npz"))
    got = alt_parse_PDB(os.path.join(GOLDEN, "2OCJ.pdb"), chains)[0]
    assert got["seq"] == str(ref["seq"]) and got["name"] == str(ref["name"]) == "2OCJ"
    assert got["num_of_chains"] == len(chains)
    assert [k[-1] for k in got if k.startswith("seq_chain_")] == list(chains)
    for ch in chains:
        assert got["seq_chain_" + ch] == str(ref["seq_" + ch])
        xyz = np.stack([np.asarray(got["coords_chain_" + ch][f"{a}_chain_{ch}"]) for a in ("N", "CA", "C", "O")], 1)
        assert np.array_equal(xyz, ref["xyz_" + ch], equal_nan=True)       # float64, bit-identical


def test_parser_edge_cases(tmp_path):
    def atom(rec, serial, name, res, ch, num, ins, x):
        return f"{rec:<6}{serial:>5} {name:<4} {res:>3} {ch}{num:>4}{ins:1}   {x:8.3f}{x+1:8.3f}{x+2:8.3f}  1.00  0.00"
    lines = []
    s = 1
    for num, ins, res in [(5, " ", "ALA"), (5, "A", "GLY"), (7, " ", "MSE"), (8, " ", "UNK")]:
        rec = "HETATM" if res == "MSE" else "ATOM"
        for name in ("N", "CA", "C", "O"):
            if num == 8 and name == "O":
                continue                                            # missing atom -> NaN
            lines.append(atom(rec, s, name, res, "B", num, ins, float(s))); s += 1
    lines.append(atom("ATOM", s, "CA", "ALA", "B", 5, " ", 99.0))   # duplicate atom: first occurrence wins
    p = tmp_path / "toy_.pdb"
    p.write_text("\n".join(lines) + "\n")
    d = alt_parse_PDB(str(p), "B")[0]
    assert d["name"] == "toy_" and d["seq"] == "AG-M-"             # 5, 5A, gap 6, MSE->M, UNK -> '-'
    ca = np.asarray(d["coords_chain_B"]["CA_chain_B"])
    assert ca[0, 0] == 2.0 and np.isnan(ca[2]).all()
    assert np.isnan(np.asarray(d["coords_chain_B"]["O_chain_B"])[4]).all() and not np.isnan(ca[4]).any()
    assert alt_parse_PDB(str(p), "Z")[0]["num_of_chains"] == 0


def test_csv_frame_matches_published_layout():
    import pandas as pd
    from thermompnn_b200 import custom_inference as ci

    class Fake:                                                     # stands in for the CUDA model: forward contract only
        def __call__(self, pdb, muts):
            import torch
            return [None if m is None else {"ddG": torch.tensor([float(m.position)])} for m in muts], None
    df, pdb_id = ci.run(os.path.join(GOLDEN, "2OCJ.pdb"), "A", Fake())
    ref = pd.read_csv(os.path.join(GOLDEN, "ref_2OCJ_A_published.csv"), index_col=0)
    assert pdb_id == "2OCJ" and list(df.columns) == list(ref.columns) and len(df) == len(ref) == 3880
    for col in ("Model", "Dataset", "position", "wildtype", "mutation", "pdb", "chain"):
        assert (df[col].astype(str).values == ref[col].astype(str).values).all(), col
