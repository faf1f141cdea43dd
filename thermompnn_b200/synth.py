"""Seeded synthetic proteins in the ``alt_parse_PDB`` dict schema.

The same dict feeds the CUDA path, the oracle and (in the build container) the
unmodified reference, so benchmark inputs and parity inputs are one object.
Coordinates are rounded to 1e-3 A like PDB files, which keeps the 48th/49th
neighbour gap far above fp32 rounding so that neighbour *sets* are reproducible.
"""
from __future__ import annotations

import zlib
from typing import List, Optional

import numpy as np

AA20 = "ACDEFGHIKLMNPQRSTVWY"


def synth_backbone(L: int, rng: np.random.Generator) -> np.ndarray:
    """[L,4,3] float64 backbone (N, CA, C, O): confined self-avoiding CA walk."""
    radius = 3.3 * L ** (1.0 / 3.0) + 4.0
    ca = np.zeros((L, 3))
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    for i in range(1, L):
        cand, nd = None, d
        for _ in range(50):
            nd = d + 0.9 * rng.normal(size=3)
            nd /= np.linalg.norm(nd)
            cand = ca[i - 1] + 3.8 * nd
            if np.linalg.norm(cand) > radius:
                continue
            if i > 1 and np.min(np.linalg.norm(ca[: i - 1] - cand, axis=1)) < 3.5:
                continue
            break
        else:
            # boxed in: step back towards the centre, ignoring the clash rule
            nd = -ca[i - 1] / max(np.linalg.norm(ca[i - 1]), 1e-6) + 0.3 * rng.normal(size=3)
            nd /= np.linalg.norm(nd)
            cand = ca[i - 1] + 3.8 * nd
        ca[i], d = cand, nd
    n = ca + np.array([-1.2, 0.5, 0.3]) + rng.normal(scale=0.1, size=(L, 3))
    c = ca + np.array([1.2, 0.6, -0.3]) + rng.normal(scale=0.1, size=(L, 3))
    o = c + np.array([0.3, 1.1, 0.2])
    return np.round(np.stack([n, ca, c, o], 1), 3)


def synth(L: int, seed: int, seq: Optional[str] = None, name: Optional[str] = None,
          chain: str = "A", gaps: Optional[List[int]] = None, arrays: bool = False) -> dict:
    """One synthetic single-chain protein as an ``alt_parse_PDB``-style dict.

    ``gaps`` lists positions to turn into missing residues ('-' in ``seq`` and
    NaN coordinates), the way ``alt_parse_PDB`` reports unresolved residues.  ``arrays=True`` keeps the coordinates as
    numpy arrays instead of nested lists (same values; cheaper to pack).
    """
    rng = np.random.default_rng(seed)
    X = synth_backbone(L, rng)
    if seq is None:
        seq = "".join(AA20[i] for i in rng.integers(0, 20, size=L))
    assert len(seq) == L
    if gaps:
        s = list(seq)
        for g in gaps:
            s[g] = "-"
            X[g] = np.nan
        seq = "".join(s)
    return {
        "name": name or f"synth_L{L}_s{seed}",
        "num_of_chains": 1,
        "seq": seq,
        f"seq_chain_{chain}": seq,
        f"coords_chain_{chain}": {
            f"N_chain_{chain}": np.ascontiguousarray(X[:, 0]) if arrays else X[:, 0].tolist(),
            f"CA_chain_{chain}": np.ascontiguousarray(X[:, 1]) if arrays else X[:, 1].tolist(),
            f"C_chain_{chain}": np.ascontiguousarray(X[:, 2]) if arrays else X[:, 2].tolist(),
            f"O_chain_{chain}": np.ascontiguousarray(X[:, 3]) if arrays else X[:, 3].tolist(),
        },
    }


def synth_multichain(lengths: List[int], seed: int, name: Optional[str] = None) -> dict:
    """Several chains (A, B, ...) placed side by side; exercises the +100
    residue_idx offset and the cross-chain positional class."""
    rng = np.random.default_rng(seed)
    out = {"name": name or f"synth_multi_s{seed}", "num_of_chains": len(lengths), "seq": ""}
    shift = np.zeros(3)
    for ci, L in enumerate(lengths):
        ch = chr(ord("A") + ci)
        X = synth_backbone(L, rng) + shift
        X = np.round(X, 3)
        shift = shift + np.array([3.3 * L ** (1 / 3.0) + 6.0, 0.0, 0.0])
        seq = "".join(AA20[i] for i in rng.integers(0, 20, size=L))
        out["seq"] += seq
        out[f"seq_chain_{ch}"] = seq
        out[f"coords_chain_{ch}"] = {
            f"N_chain_{ch}": X[:, 0].tolist(), f"CA_chain_{ch}": X[:, 1].tolist(),
            f"C_chain_{ch}": X[:, 2].tolist(), f"O_chain_{ch}": X[:, 3].tolist()}
    return out


_ONE_TO_THREE = dict(zip("ARNDCQEGHILKMFPSTWYV", ['ALA', 'ARG', 'ASN', 'ASP', 'CYS', 'GLN', 'GLU', 'GLY', 'HIS', 'ILE', 'LEU', 'LYS',
                                                     'MET', 'PHE', 'PRO', 'SER', 'THR', 'TRP', 'TYR', 'VAL']))


def write_pdb(pdb: dict, path: str, first_resnum: int = 1) -> None:
    """Write a protein dict as PDB text (ATOM records of the backbone atoms; gap residues are skipped, so a parser
    sees them as numbering gaps).  Round-trips through ``pdb.alt_parse_PDB`` / ``parse_PDB``."""
    lines, serial = [], 1
    for key in pdb:
        if key[:10] != "seq_chain_":
            continue
        ch = key[-1:]
        co = pdb[f"coords_chain_{ch}"]
        for i, a in enumerate(pdb[key]):
            if a == "-":
                continue
            for name in ("N", "CA", "C", "O"):
                x, y, z = co[f"{name}_chain_{ch}"][i]
                lines.append(f"ATOM  {serial:>5} {name:<4} {_ONE_TO_THREE[a]:>3} {ch}{first_resnum + i:>4}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00")
                serial += 1
        lines.append("TER")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\nEND\n")


# ---- the BASELINE.json workloads ------------------------------------------
def config2() -> List[dict]:
    """C2: one 300-residue protein."""
    return [synth(300, seed=0)]


def config3(n: int = 256) -> List[dict]:
    """C3: 256 proteins, lengths U[150,500] from seed 1234."""
    lengths = np.random.default_rng(1234).integers(150, 501, size=256)[:n]
    return [synth(int(L), seed=1000 + i) for i, L in enumerate(lengths)]


def config5(n: int = 64, L: int = 2000) -> List[dict]:
    """C5: 64 proteins of 2000 residues."""
    return [synth(L, seed=2000 + i) for i in range(n)]


def megascale_like(names_and_seqs) -> List[dict]:
    """C4 shape: given (WT_name, aa_seq) pairs, synthetic coords seeded by crc32(name)."""
    return [synth(len(s), seed=zlib.crc32(n.encode()), seq=s, name=n) for n, s in names_and_seqs]


def config4(shape_npz: str):
    """C4: the Megascale-shaped batch (298 proteins, L 32..72) from the committed shape table (names, wild-type
    sequences and the 271 231 listed single substitutions of the reference's mega_{train,val,test}.csv; made by
    tools/make_golden_r2.py).  Returns (proteins, protein_index[int], position[int], mutant_letter_index[int])."""
    d = np.load(shape_npz)
    pdbs = megascale_like(zip([str(n) for n in d["names"]], [str(s) for s in d["seqs"]]))
    return pdbs, d["protein"].astype(np.int64), d["position"].astype(np.int64), d["mutant"].astype(np.int64)
