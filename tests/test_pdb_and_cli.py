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


REF = "/root/reference"


def _random_pdb(path, rng):
    """PDB text with the quirks the reference parser has rules for: several chains, numbering gaps, insertion codes,
    negative residue numbers, MSE HETATMs, unknown residues, missing atoms, duplicated (alt-loc) atoms, side chains."""
    aa3 = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER",
           "THR", "TRP", "TYR", "VAL", "UNK", "MSE"]
    lines, serial = ["HEADER    TEST", "REMARK 1"], 1
    for ch in rng.permutation(list("ABcd1"))[:int(rng.integers(1, 5))]:
        num = int(rng.integers(-3, 20))
        for _ in range(int(rng.integers(3, 60))):
            num += 1 if rng.random() > 0.1 else int(rng.integers(2, 5))
            for ins in ([" "] + (["A", "B"][:int(rng.integers(1, 3))] if rng.random() < 0.08 else [])):
                res = aa3[int(rng.integers(0, len(aa3)))]
                rec = "HETATM" if res == "MSE" else "ATOM"
                for name in ("N", "CA", "C", "O", "CB"):
                    if name != "CB" and rng.random() < 0.03:
                        continue
                    for _rep in range(2 if rng.random() < 0.05 else 1):
                        x, y, z = rng.normal(scale=30, size=3)
                        lines.append(f"{rec:<6}{serial:>5} {name:<4} {res:>3} {ch}{num:>4}{ins:1}   {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00")
                        serial += 1
        lines.append("TER")
    open(path, "w").write("\n".join(lines) + "\nEND\n")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "protein_mpnn_utils.py")), reason="needs the reference checkout")
def test_parser_differential_against_live_reference(tmp_path):
    """Random PDB text through the reference's own parsers and ours: every key, every float64, resn_list included
    (also with the default chain list, where upstream ends on list('no_chain'))."""
    import sys
    sys.path.insert(0, REF)
    import protein_mpnn_utils as R
    from thermompnn_b200.pdb import parse_PDB
    rng = np.random.default_rng(5)
    for it in range(12):
        p = str(tmp_path / f"rand{it}.pdb")
        _random_pdb(p, rng)
        for chains in (None, "A", "AB", ["c", "1", "Z"], "ZA"):
            for ours, theirs in ((alt_parse_PDB(p, chains), R.alt_parse_PDB(p, chains)), (parse_PDB(p, chains), R.parse_PDB(p, chains))):
                assert len(ours) == len(theirs) == 1
                a, b = ours[0], theirs[0]
                assert list(a.keys()) == list(b.keys()), (chains, list(a.keys()), list(b.keys()))
                for k in b:
                    if k.startswith("coords_chain_"):
                        for kk in b[k]:
                            assert np.array_equal(np.asarray(a[k][kk], dtype=np.float64), np.asarray(b[k][kk], dtype=np.float64), equal_nan=True), (k, kk)
                    else:
                        assert a[k] == b[k], (chains, k, a[k], b[k])


def test_parser_resn_list_semantics():
    """resn_list is what the LAST chain letter iterated produced (protein_mpnn_utils.py:346)."""
    p = os.path.join(GOLDEN, "2OCJ.pdb")
    assert alt_parse_PDB(p)[0]["resn_list"] == list("no_chain")           # default chain list ends on letters not in the file
    assert alt_parse_PDB(p, "BZ")[0]["resn_list"] == list("no_chain")
    a = alt_parse_PDB(p, "A")[0]["resn_list"]
    assert alt_parse_PDB(p, "BA")[0]["resn_list"] == a and a[0] == "96" and len(a) == 194
