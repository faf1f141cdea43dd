"""Site-saturation mutation lists, as ``analysis/SSM.py:16-29`` builds them."""
from typing import List, Optional

import numpy as np

from .datasets import ALPHABET, Mutation


def get_ssm_mutations(pdb: dict) -> List[Optional[str]]:
    """``wt + str(pos) + mut`` for the 20 letters (identity included) at every residue;
    a single ``None`` for each gap ('-') residue."""
    out: List[Optional[str]] = []
    for pos, wt in enumerate(pdb["seq"]):
        if wt != "-":
            out.extend(wt + str(pos) + m for m in ALPHABET[:-1])
        else:
            out.append(None)
    return out


def to_mutation_objects(mutation_list, pdb_name: str = "") -> List[Optional[Mutation]]:
    """Parse 'A12G'-style strings into ``Mutation`` objects (custom_inference.py:79-90)."""
    out: List[Optional[Mutation]] = []
    for m in mutation_list:
        if m is None:
            out.append(None)
            continue
        m = m.strip()
        wt, pos, mut = m[0], int(m[1:-1]), m[-1]
        assert wt in ALPHABET, f"Wild type residue {wt} invalid, please try again with one of the following options: {ALPHABET}"
        assert mut in ALPHABET, f"Mutant residue {mut} invalid, please try again with one of the following options: {ALPHABET}"
        out.append(Mutation(position=pos, wildtype=wt, mutation=mut, ddG=None, pdb=pdb_name))
    return out


def retrieve_best_mutants(df_slice, allow_cys: bool = True) -> list:
    """Most stabilising substitution at each position of one protein's rows (analysis/SSM.py:32-43)."""
    import pandas as pd
    best = []
    for p in df_slice.position.unique():
        p_slice = df_slice.loc[df_slice["position"] == p].reset_index(drop=True)
        if not allow_cys:
            p_slice = p_slice.loc[p_slice["mutation"] != "C"].reset_index(drop=True)
        best.append(p_slice.iloc[pd.to_numeric(p_slice["ddG_pred"]).idxmin()]["mutation"])
    return best


def run_ssm(name: str, model, dataset_name: str, dataset, centrality: bool = False, pick_best: bool = False,
            include_cys: bool = False, proteins_per_launch: int = 64):
    """Site-saturation mutagenesis over a whole dataset: the loop of ``analysis/SSM.py:93-164`` with the same output
    frame (columns ``WT Seq, Model, Dataset, ddG_pred, position, wildtype, mutation, neighbors, best_AA, pdb``, one row
    per (position, substitution) in upstream's order, the same ``pick_best`` / ``include_cys`` filtering), but many
    proteins per GPU launch (``model.ssm_tables``; for ``ProteinMPNNBaseline`` the ddG proxy is -log_prob) and one frame
    construction per protein instead of one ``DataFrame.loc`` write per cell."""
    import pandas as pd
    from .benchmarking import compute_centrality
    items = [(pdb, muts) for pdb, muts in dataset]
    frames = []
    for c0 in range(0, len(items), proteins_per_launch):
        chunk = [pdb[0] if isinstance(pdb, (list, tuple)) else pdb for pdb, _ in items[c0:c0 + proteins_per_launch]]
        if hasattr(model, "ssm_tables"):
            tables = model.ssm_tables(chunk)                                  # [L, 20] ddG, NaN rows at gaps
        else:                                                                 # ProteinMPNNBaseline: ddG proxy = -log_prob (:62)
            tables = [-lp[:, :20] for lp in model.log_prob_tables(chunk)]
        for pdb, ddg in zip(chunk, tables):
            seq = pdb["seq"]
            pos = [i for i, a in enumerate(seq) if a != "-"]
            pid = pdb["name"].strip(".pdb")                                   # (sic) character strip, as upstream (:141)
            key = pdb["name"] if "Megascale" not in dataset_name else pdb["name"] + ".pdb"
            ws = getattr(dataset, "wt_seqs", {})
            wt_seq = ws.get(key, ws.get(key.replace(":", "|")))
            df = pd.DataFrame({
                "WT Seq": wt_seq, "Model": name, "Dataset": dataset_name,
                "ddG_pred": np.asarray(ddg)[pos].reshape(-1).astype(object),
                "position": np.repeat(np.asarray(pos, dtype=object), 20), "wildtype": np.repeat([seq[i] for i in pos], 20),
                "mutation": list(ALPHABET[:-1]) * len(pos), "neighbors": None, "best_AA": None, "pdb": pid,
            }, columns=["WT Seq", "Model", "Dataset", "ddG_pred", "position", "wildtype", "mutation", "neighbors", "best_AA", "pdb"])
            if centrality:
                coord_chain = [c for c in pdb.keys() if "coords" in c][0]
                nb = compute_centrality(pdb[coord_chain], basis_atom="CA", chain=coord_chain[-1], radius=10.0)
                df["neighbors"] = np.repeat(np.asarray(nb)[pos], 20).astype(object)
            if pick_best:
                best = retrieve_best_mutants(df, allow_cys=include_cys)
                df["best_AA"] = np.repeat(best, 20)
                df = df.drop_duplicates(subset=["position"], keep="first")
            elif not include_cys:
                df = df.loc[df["mutation"] != "C"]
            frames.append(df)
    return pd.concat(frames).reset_index(drop=True) if frames else pd.DataFrame()
