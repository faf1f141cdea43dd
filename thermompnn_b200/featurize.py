"""Host-side packing of parsed proteins into the flat arrays the C ABI takes.

Covers the part of ``tied_featurize`` (protein_mpnn_utils.py:353-605) that
``TransferModel.forward`` consumes (transfer_model.py:78-81), generalised from the
reference's B=1 to a ragged batch: proteins are concatenated and ``offsets`` marks the cuts.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .datasets import ALPHABET

_LUT = np.full(256, -1, dtype=np.int32)
for _i, _a in enumerate(ALPHABET):
    _LUT[ord(_a)] = _i
_LUT[ord("-")] = 20  # '-' -> 'X' (protein_mpnn_utils.py:454)


class Packed:
    """offsets [B+1] i32, X [N,4,3] f32, S/residue_idx/chain_enc [N] i32, mask [N] u8."""
    __slots__ = ("offsets", "X", "S", "residue_idx", "chain_enc", "mask")

    def __init__(self, offsets, X, S, residue_idx, chain_enc, mask):
        self.offsets, self.X, self.S, self.residue_idx, self.chain_enc, self.mask = offsets, X, S, residue_idx, chain_enc, mask

    @property
    def n_proteins(self): return self.offsets.size - 1

    @property
    def n_residues(self): return int(self.offsets[-1])


def _chains(pdb: dict):
    """Per chain of one protein, in dict order: (x [n,4,3] float64 with NaN for missing, letters [n] int32, n)."""
    letters = [k[-1:] for k in pdb if k[:10] == "seq_chain_"]   # dict order, as protein_mpnn_utils.py:387
    if not letters:
        raise ValueError("protein dict has no 'seq_chain_*' entries")
    out = []
    for ch in letters:
        seq = pdb[f"seq_chain_{ch}"]
        co = pdb[f"coords_chain_{ch}"]
        n = len(seq)
        x = np.empty((n, 4, 3), dtype=np.float64)
        for i, a in enumerate(("N", "CA", "C", "O")):           # nested lists (upstream's schema) or [n,3] arrays
            v = np.asarray(co[f"{a}_chain_{ch}"], dtype=np.float64)
            if v.shape != (n, 3):
                raise ValueError(f"chain {ch}: coordinates {(v.shape[0], 4, 3) if v.ndim == 2 else v.shape} do not match sequence length {n}")
            x[:, i, :] = v
        s = _LUT[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
        if (s < 0).any():
            bad = seq[int(np.argmax(s < 0))]
            raise ValueError(f"'{bad}' is not in list")       # what ALPHABET.index raises upstream (:568)
        out.append((x, s, n))
    return out


def _pack(protein_chains):
    """[[(x, s, n) per chain] per protein] -> Packed: label arrays per chain, NaN handling once on the whole batch."""
    xs, ss, cs, rs, lens = [], [], [], [], []
    for chains in protein_chains:
        pos = 0
        for c, (x, s, n) in enumerate(chains, start=1):
            xs.append(x); ss.append(s)
            cs.append(np.full(n, c, dtype=np.int32))                               # 1-based chain label (:471)
            rs.append(np.arange(100 * (c - 1) + pos, 100 * (c - 1) + pos + n, dtype=np.int32))   # (:473)
            pos += n
        lens.append(pos)
    X = np.concatenate(xs, 0)
    mask = np.isfinite(X.reshape(X.shape[0], 12).sum(1))                           # (:576)
    X = np.nan_to_num(X, nan=0.0, posinf=np.inf, neginf=-np.inf).astype(np.float32)   # NaN -> 0 (:577), cast (:595)
    offsets = np.zeros(len(lens) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    return Packed(offsets, np.ascontiguousarray(X), np.ascontiguousarray(np.concatenate(ss)), np.concatenate(rs), np.concatenate(cs),
                  mask.astype(np.uint8))


def featurize_one(pdb: dict):
    p = _pack([_chains(pdb)])
    return p.X, p.S, p.residue_idx, p.chain_enc, p.mask


def featurize_batch(pdbs: Sequence[dict]) -> Packed:
    return _pack([_chains(p) for p in pdbs])
