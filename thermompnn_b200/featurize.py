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


def featurize_one(pdb: dict):
    letters = [k[-1:] for k in pdb if k[:10] == "seq_chain_"]   # dict order, as protein_mpnn_utils.py:387
    if not letters:
        raise ValueError("protein dict has no 'seq_chain_*' entries")
    xs, ss, cs, rs = [], [], [], []
    pos = 0
    for c, ch in enumerate(letters, start=1):
        seq = pdb[f"seq_chain_{ch}"]
        co = pdb[f"coords_chain_{ch}"]
        x = np.stack([np.asarray(co[f"{a}_chain_{ch}"], dtype=np.float64) for a in ("N", "CA", "C", "O")], 1)
        n = len(seq)
        if x.shape != (n, 4, 3):
            raise ValueError(f"chain {ch}: coordinates {x.shape} do not match sequence length {n}")
        s = _LUT[np.frombuffer(seq.encode("ascii"), dtype=np.uint8)]
        if (s < 0).any():
            bad = seq[int(np.argmax(s < 0))]
            raise ValueError(f"'{bad}' is not in list")       # what ALPHABET.index raises upstream (:568)
        xs.append(x); ss.append(s)
        cs.append(np.full(n, c, dtype=np.int32))                       # 1-based chain label (:471)
        rs.append((100 * (c - 1) + np.arange(pos, pos + n)).astype(np.int32))   # (:473)
        pos += n
    X = np.concatenate(xs, 0)
    mask = np.isfinite(X.sum((1, 2)))                                   # (:576)
    X = np.where(np.isnan(X), 0.0, X).astype(np.float32)               # (:577, :595)
    return X, np.concatenate(ss), np.concatenate(rs), np.concatenate(cs), mask.astype(np.uint8)


def featurize_batch(pdbs: Sequence[dict]) -> Packed:
    parts = [featurize_one(p) for p in pdbs]
    lens = np.asarray([p[0].shape[0] for p in parts], dtype=np.int64)
    offsets = np.zeros(len(parts) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    cat = lambda i: np.ascontiguousarray(np.concatenate([p[i] for p in parts], 0))
    return Packed(offsets, cat(0), cat(1), cat(2), cat(3), cat(4))
