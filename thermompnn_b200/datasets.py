"""API types and dataset adapters shared with the reference's callers (``datasets.py`` upstream).

``Mutation`` is the reference's dataclass (datasets.py:25-31).  The three dataset classes yield exactly what the
reference's yield -- ``(pdb, [Mutation, ...])`` per protein, ``pdb`` being the one-element list the parsers return --
so they feed ``TransferModel.forward`` one protein at a time like upstream, or ``TransferModel.predict_batch`` /
``benchmarking.run_prediction_*`` many proteins per launch.  Differences, all deliberate:

* construction is vectorised (one ``groupby`` instead of one ``DataFrame.query`` per protein, datasets.py:117-123);
* ``cfg.data_loc.megascale_pdbs`` may be ``None``: coordinates are then synthetic (``synth.synth`` seeded by
  ``crc32(WT_name)``), which is how BASELINE.json's Megascale-shaped workload is built where the PDB files are not shipped;
* FireProt's mis-alignment contingency (datasets.py:243-248) uses a built-in global alignment (match 1, mismatch 0,
  gaps 0 = Biopython's ``globalxx``); Biopython is not needed.  Among equal-score alignments Biopython's pick is an
  implementation detail that is NOT reproduced (the traceback here prefers diagonal, then gap-in-seq2, then gap-in-seq1).
"""
from __future__ import annotations

import os
import pickle
import zlib
from dataclasses import dataclass
from math import isnan
from typing import Dict, List, Optional

import numpy as np

ALPHABET = "ACDEFGHIKLMNPQRSTVWYX"  # transfer_model.py:11  (datasets.py's own ALPHABET ends in '-' and is unused there)


@dataclass
class Mutation:
    """Same fields as the reference's ``datasets.Mutation`` (datasets.py:25-31);
    ``position`` is 0-based into the concatenated chain sequence."""
    position: int
    wildtype: str
    mutation: str
    ddG: Optional[float] = None
    pdb: Optional[str] = ""


def _tensor1(x: float):
    import torch
    return torch.tensor([x], dtype=torch.float32)


def seq1_index_to_seq2_index(seqA: str, seqB: str, index: int) -> Optional[int]:
    """datasets.py:34-58 on an alignment given as its two gapped strings."""
    cur, aln_idx = 0, 0
    for aln_idx, char in enumerate(seqA):
        if char != "-":
            cur += 1
        if cur > index:
            break
    if seqB[aln_idx] == "-":
        return None
    seq2_idx = aln_idx - seqB[:aln_idx + 1].count("-")
    return None if seq2_idx < 0 else seq2_idx


def global_align_xx(a: str, b: str):
    """Needleman-Wunsch with match = 1, mismatch = 0, gap = 0 (Biopython ``pairwise2.align.globalxx``): returns the
    two gapped strings of ONE optimal alignment (see the module docstring for the tie rule)."""
    n, m = len(a), len(b)
    A = np.frombuffer(a.encode(), dtype=np.uint8)
    B = np.frombuffer(b.encode(), dtype=np.uint8)
    S = np.zeros((n + 1, m + 1), dtype=np.int32)
    for i in range(1, n + 1):
        match = (A[i - 1] == B).astype(np.int32)
        diag = S[i - 1, :-1] + match
        row = np.maximum(diag, S[i - 1, 1:])
        S[i, 1:] = np.maximum.accumulate(row)      # a left step costs nothing, so the row is a running maximum
    i, j, ra, rb = n, m, [], []
    while i > 0 or j > 0:
        if i > 0 and j > 0 and S[i, j] == S[i - 1, j - 1] + (1 if a[i - 1] == b[j - 1] else 0):
            ra.append(a[i - 1]); rb.append(b[j - 1]); i -= 1; j -= 1
        elif i > 0 and S[i, j] == S[i - 1, j]:
            ra.append(a[i - 1]); rb.append("-"); i -= 1
        else:
            ra.append("-"); rb.append(b[j - 1]); j -= 1
    return "".join(reversed(ra)), "".join(reversed(rb))


class _ListDataset:
    def __len__(self):
        return len(self.wt_names)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class MegaScaleDataset(_ListDataset):
    """datasets.py:61-164.  ``cfg.data_loc``: ``megascale_csv``, ``megascale_splits`` (pickle of {train,val,test,...}
    -> WT names), ``megascale_pdbs`` (directory, or None for synthetic coordinates)."""

    def __init__(self, cfg, split):
        import pandas as pd
        self.cfg, self.split = cfg, split
        df = pd.read_csv(cfg.data_loc.megascale_csv, usecols=["ddG_ML", "mut_type", "WT_name", "aa_seq", "dG_ML"])
        df = df.loc[df.ddG_ML != "-", :].reset_index(drop=True)                                   # :70
        df = df.loc[~df.mut_type.str.contains("ins") & ~df.mut_type.str.contains("del") &
                    ~df.mut_type.str.contains(":"), :].reset_index(drop=True)                     # :72
        self.df = df
        with open(cfg.data_loc.megascale_splits, "rb") as f:
            splits = pickle.load(f)
        if "reduce" not in cfg:
            cfg.reduce = ""
        if split == "all":
            names = list(splits["train"]) + list(splits["val"]) + list(splits["test"])
        elif cfg.reduce == "prot" and split == "train":
            names = list(np.random.choice(splits["train"], 58))                                   # :108-110
        else:
            names = list(splits[split])
        self.wt_names = names
        groups = {k: g for k, g in df.groupby("WT_name", sort=False)}
        self.wt_seqs: Dict[str, str] = {}
        self.mut_rows = {}
        for wt_name in names:
            g = groups[wt_name]
            muts = g.loc[g.mut_type != "wt"].reset_index(drop=True)
            if type(cfg.reduce) is float and split == "train":
                muts = muts.sample(frac=float(cfg.reduce), replace=False)                         # :120-121
            self.mut_rows[wt_name] = muts
            self.wt_seqs[wt_name] = g.loc[g.mut_type == "wt", "aa_seq"].iloc[0]

    def __getitem__(self, index):
        wt_key = self.wt_names[index]
        mut_data, wt_seq = self.mut_rows[wt_key], self.wt_seqs[wt_key]
        wt_name = wt_key.split(".pdb")[0].replace("|", ":")
        pdb_dir = self.cfg.data_loc.megascale_pdbs if "megascale_pdbs" in self.cfg.data_loc else None
        if pdb_dir:
            from .pdb import parse_PDB
            pdb = parse_PDB(os.path.join(pdb_dir, f"{wt_name}.pdb"))
            assert len(pdb[0]["seq"]) == len(wt_seq)
            pdb[0]["seq"] = wt_seq                                                                 # :140
        else:
            from .synth import synth
            pdb = [synth(len(wt_seq), seed=zlib.crc32(wt_key.encode()), seq=wt_seq, name=wt_name)]
        mutations = []
        for mut_type, aa_seq, ddG_ML in zip(mut_data.mut_type, mut_data.aa_seq, mut_data.ddG_ML):
            assert len(aa_seq) == len(wt_seq)
            wt, mut, idx = mut_type[0], mut_type[-1], int(mut_type[1:-1]) - 1
            assert wt_seq[idx] == wt and aa_seq[idx] == mut
            mutations.append(Mutation(idx, wt, mut, -_tensor1(float(ddG_ML)), wt_name))           # :161-162
        return pdb, mutations


class FireProtDataset(_ListDataset):
    """datasets.py:167-256.  ``cfg.data_loc``: ``fireprot_csv``, ``fireprot_splits``, ``fireprot_pdbs``."""

    def __init__(self, cfg, split):
        import pandas as pd
        self.cfg, self.split = cfg, split
        df = pd.read_csv(cfg.data_loc.fireprot_csv).dropna(subset=["ddG"])
        df = df.where(pd.notnull(df), None)
        self.df = df
        self.seq_to_data = {s: g.reset_index(drop=True) for s, g in df.groupby("pdb_sequence", sort=False)}
        with open(cfg.data_loc.fireprot_splits, "rb") as f:
            splits = pickle.load(f)
        self.wt_names = [j for sub in splits.values() for j in sub] if split == "all" else list(splits[split])
        by_id = {k: g.reset_index(drop=True) for k, g in df.groupby("pdb_id_corrected", sort=False)}
        self.mut_rows = {n: by_id[n] for n in self.wt_names}
        self.wt_seqs = {n: self.mut_rows[n].pdb_sequence[0] for n in self.wt_names}

    def __getitem__(self, index):
        from .pdb import parse_PDB
        wt_name = self.wt_names[index]
        seq = self.wt_seqs[wt_name]
        data = self.seq_to_data[seq]
        pdb = parse_PDB(os.path.join(self.cfg.data_loc.fireprot_pdbs, f"{data.pdb_id_corrected[0]}.pdb"))
        pseq = pdb[0]["seq"]
        mutations = []
        for row in data.itertuples(index=False):
            pdb_idx = row.pdb_position
            if not (pseq[pdb_idx] == row.wild_type == row.pdb_sequence[row.pdb_position]):
                sa, sb = global_align_xx(seq, pseq.replace("-", "X"))                              # :244
                pdb_idx = seq1_index_to_seq2_index(sa, sb, row.pdb_position)
                if pdb_idx is None:
                    continue
                assert pseq[pdb_idx] == row.wild_type == row.pdb_sequence[row.pdb_position]
            ddG = None if row.ddG is None or isnan(row.ddG) else _tensor1(row.ddG)
            mutations.append(Mutation(pdb_idx, pseq[pdb_idx], row.mutation, ddG, wt_name))
        return pdb, mutations


class ddgBenchDataset(_ListDataset):
    """datasets.py:259-322 (S669, SSym, P53, Myoglobin CSVs: columns PDB, MUT, DDG, SEQ)."""

    def __init__(self, cfg, pdb_dir, csv_fname):
        import pandas as pd
        self.cfg, self.pdb_dir = cfg, pdb_dir
        df = pd.read_csv(csv_fname)
        self.df = df
        self.wt_names = df.PDB.unique()
        self.wt_seqs, self.mut_rows = {}, {}
        for wt_name_query, g in df.groupby("PDB", sort=False):
            wt_name = wt_name_query[:-1]
            self.mut_rows[wt_name] = g.reset_index(drop=True)
            if "S669" not in self.pdb_dir:
                self.wt_seqs[wt_name] = self.mut_rows[wt_name].SEQ[0]

    def __getitem__(self, index):
        from .pdb import alt_parse_PDB
        wt_name = self.wt_names[index]
        chain = [wt_name[-1]]
        wt_name = wt_name.split(".pdb")[0][:-1]
        mut_data = self.mut_rows[wt_name]
        pdb = alt_parse_PDB(os.path.join(self.pdb_dir, wt_name + ".pdb"), chain)
        resn_list, pseq = pdb[0]["resn_list"], pdb[0]["seq"]
        mutations = []
        for mut_info, DDG in zip(mut_data.MUT, mut_data.DDG):
            wtAA, mutAA = mut_info[0], mut_info[-1]
            try:
                pdb_idx = resn_list.index(mut_info[1:-1])
            except ValueError:           # positions with insertion codes are skipped upstream too (:297-298)
                continue
            if pseq[pdb_idx] != wtAA:    # contingency for mis-alignments (:301-313)
                gaps = [g for g in (pseq if "S669" in self.pdb_dir else pseq[:pdb_idx + 10]) if g == "-"]
                pdb_idx += len(gaps) if len(gaps) > 0 else 1
                assert pseq[pdb_idx] == wtAA
            ddG = None if DDG is None or isnan(DDG) else _tensor1(DDG * -1.0)
            mutations.append(Mutation(pdb_idx, pseq[pdb_idx], mutAA, ddG, wt_name))
        return pdb, mutations
