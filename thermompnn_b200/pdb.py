"""PDB text -> the dict schema the model consumes.

Behaviour-compatible with the reference's ``alt_parse_PDB`` / ``alt_parse_PDB_biounits``
(protein_mpnn_utils.py:183-350), re-implemented as one vectorisable pass over the file:

* only ``ATOM`` records, plus ``HETATM`` MSE rewritten to MET (:218-220);
* chain = column 22; atom name = cols 13-16 stripped; residue field = cols 23-27 (number + optional insertion
  letter, :227-235); x, y, z = cols 31-38 / 39-46 / 47-54 as float64;
* first occurrence of an atom wins (alt-locs beyond the first are ignored, :249-250);
* every integer residue number between the chain's min and max is emitted; missing ones become '-' with NaN
  coordinates (:256-276); insertion codes are emitted in sorted order after their residue number;
* unknown 3-letter codes map to '-' in ``seq`` while their coordinates are kept (:261);
* ``name`` = file basename minus its last 4 characters (:340-341); ``resn_list`` keeps the residue labels of the
  last chain parsed (:346).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_AA3 = ['ALA', 'ARG', 'ASN', 'ASP', 'CYS', 'GLN', 'GLU', 'GLY', 'HIS', 'ILE',
        'LEU', 'LYS', 'MET', 'PHE', 'PRO', 'SER', 'THR', 'TRP', 'TYR', 'VAL']
_AA1 = "ARNDCQEGHILKMFPSTWYV"
_THREE_TO_ONE = dict(zip(_AA3, _AA1))
_BACKBONE = ("N", "CA", "C", "O")
_DEFAULT_CHAINS = [chr(c) for c in range(ord('A'), ord('Z') + 1)] + [chr(c) for c in range(ord('a'), ord('z') + 1)] + \
                  [str(i) for i in range(300)]


def _parse_chain(lines: List[str], chain: str):
    """One chain -> (xyz [L,4,3] float64 with NaN for missing, seq str, residue labels) or None if absent."""
    xyz: Dict[int, Dict[str, Dict[str, np.ndarray]]] = {}
    seq: Dict[int, Dict[str, str]] = {}
    labels: List[str] = []
    seen = set()
    lo, hi = None, None
    for line in lines:
        if line[:6] == "HETATM" and line[17:20] == "MSE":
            line = line.replace("HETATM", "ATOM  ").replace("MSE", "MET")
        if line[:4] != "ATOM" or line[21:22] != chain:
            continue
        label = line[22:27].strip()
        if label not in seen:
            seen.add(label)
            labels.append(label)
        if label[-1].isalpha():
            ins, num = label[-1], int(label[:-1]) - 1
        else:
            ins, num = "", int(label) - 1
        lo = num if lo is None or num < lo else lo
        hi = num if hi is None or num > hi else hi
        seq.setdefault(num, {}).setdefault(ins, line[17:20])
        atoms = xyz.setdefault(num, {}).setdefault(ins, {})
        name = line[12:16].strip()
        if name not in atoms:
            atoms[name] = np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    if lo is None:
        return None
    out_seq, out_xyz = [], []
    nan3 = np.full(3, np.nan)
    for num in range(lo, hi + 1):
        if num in seq:
            for ins in sorted(seq[num]):
                out_seq.append(_THREE_TO_ONE.get(seq[num][ins], "-"))
        else:
            out_seq.append("-")
        if num in xyz:
            for ins in sorted(xyz[num]):
                out_xyz.append([xyz[num][ins].get(a, nan3) for a in _BACKBONE])
        else:
            out_xyz.append([nan3] * 4)
    return np.asarray(out_xyz, dtype=np.float64).reshape(-1, 4, 3), "".join(out_seq), labels


def alt_parse_PDB(path_to_pdb: str, input_chain_list: Optional[Sequence[str]] = None) -> List[dict]:
    """Same call and return shape as the reference's ``alt_parse_PDB(path, chains)`` (backbone atoms only).
    ``input_chain_list`` is iterated element-wise, so a string like ``"AB"`` means chains A and B (as upstream)."""
    with open(path_to_pdb, "rb") as f:
        lines = [ln.decode("utf-8", "ignore").rstrip() for ln in f]
    chains = input_chain_list if input_chain_list else _DEFAULT_CHAINS
    d: dict = {}
    concat, n_chains, labels = "", 0, []
    for ch in chains:
        parsed = _parse_chain(lines, ch)
        if parsed is None:
            continue
        xyz, seq, labels = parsed
        concat += seq
        d["seq_chain_" + ch] = seq
        d["coords_chain_" + ch] = {f"{a}_chain_{ch}": xyz[:, i, :].tolist() for i, a in enumerate(_BACKBONE)}
        n_chains += 1
    d["name"] = os.path.basename(path_to_pdb)[:-4]
    d["num_of_chains"] = n_chains
    d["seq"] = concat
    d["resn_list"] = list(labels)
    return [d]
