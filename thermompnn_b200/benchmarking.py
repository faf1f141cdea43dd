"""Dataset-level inference loops of ``analysis/thermompnn_benchmarking.py`` (:68-187) on the CUDA path.

Upstream runs one protein per forward call and updates five torchmetrics objects once per MUTATION
(``metric.update(out["ddG"], mut.ddG)`` on CUDA scalars) and writes its raw-prediction frame one CELL at a time.
Here the proteins of a dataset go through ``model.predict_batch`` many per launch (ragged batch), the predictions
come back in one device->host copy per batch, and metrics / the raw-prediction frame are computed once, vectorised.
Results carry the same names: the ``results`` rows ("Model", "Dataset", "ddG r2|mse|rmse|spearman|pearson") and the
``<name>_<dataset>_raw_preds.csv`` columns.

Metric definitions are torchmetrics' (R2Score, MeanSquaredError(squared=True/False), SpearmanCorrCoef with average
ranks for ties, PearsonCorrCoef); torchmetrics itself is not needed.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .datasets import Mutation


def compute_centrality(xyz: dict, basis_atom: str = "CA", radius: float = 10.0, chain: str = "A", **_unused) -> np.ndarray:
    """Number of residues within ``radius`` of each residue (thermompnn_benchmarking.py:20-35); NaN coordinates count
    as far away (``nan_to_num(nan=2*radius)`` upstream)."""
    c = np.asarray(xyz[basis_atom + f"_chain_{chain}"], dtype=np.float64)
    d = np.sqrt(((c[:, None, :] - c[None, :, :]) ** 2).sum(-1))
    d = np.where(np.isnan(d), 2 * radius, d)
    return (d < radius).sum(-1) - 1


def _rank_average(x: np.ndarray) -> np.ndarray:
    order = np.argsort(x, kind="stable")
    ranks = np.empty(len(x), dtype=np.float64)
    xs = x[order]
    bounds = np.concatenate([[0], np.nonzero(xs[1:] != xs[:-1])[0] + 1, [len(x)]])
    for a, b in zip(bounds[:-1], bounds[1:]):
        ranks[order[a:b]] = 0.5 * (a + b - 1) + 1.0
    return ranks


def regression_metrics(pred: np.ndarray, target: np.ndarray) -> Dict[str, float]:
    """The five metrics of ``get_metrics()`` (thermompnn_benchmarking.py:68-75) on all (prediction, target) pairs."""
    p, t = np.asarray(pred, dtype=np.float64), np.asarray(target, dtype=np.float64)
    out: Dict[str, float] = {}
    if len(p) < 2:
        return out
    err = p - t
    mse = float((err ** 2).mean())
    ss_tot = float(((t - t.mean()) ** 2).sum())
    out["r2"] = 1.0 - float((err ** 2).sum()) / ss_tot if ss_tot > 0 else float("nan")
    out["mse"], out["rmse"] = mse, float(np.sqrt(mse))

    def pearson(a, b):
        a, b = a - a.mean(), b - b.mean()
        den = np.sqrt((a ** 2).sum() * (b ** 2).sum())
        return float((a * b).sum() / den) if den > 0 else float("nan")
    out["spearman"] = pearson(_rank_average(p), _rank_average(t))
    out["pearson"] = pearson(p, t)
    return out


def _predict_dataset(model, dataset, proteins_per_launch: int, max_batches: Optional[int] = None):
    """[(pdb, mutations, preds)] for every protein of the dataset, ``proteins_per_launch`` proteins per GPU call."""
    import torch
    items, chunk = [], []

    def flush():
        if not chunk:
            return
        with torch.no_grad():
            if hasattr(model, "predict_batch"):
                outs = model.predict_batch(chunk)
            else:                                   # any object with the reference's forward contract
                outs = [model(pdb, muts)[0] for pdb, muts in chunk]
        items.extend((pdb, muts, out) for (pdb, muts), out in zip(chunk, outs))
        chunk.clear()
    for i, (pdb, mutations) in enumerate(dataset):
        chunk.append((pdb, mutations))
        if len(chunk) >= proteins_per_launch:
            flush()
        if max_batches is not None and i >= max_batches:
            break
    flush()
    return items


def _pairs(items):
    """(pred, target) arrays over the mutations that carry an experimental ddG -- ONE device->host copy."""
    import torch
    pred_t, targ = [], []
    for _, muts, outs in items:
        for m, o in zip(muts, outs):
            if m is not None and o is not None and m.ddG is not None:
                pred_t.append(o["ddG"])
                targ.append(float(m.ddG))
    pred = torch.cat(pred_t).cpu().numpy() if pred_t else np.zeros(0, np.float32)
    return pred, np.asarray(targ, dtype=np.float32)


def run_prediction_default(name, model, dataset_name, dataset, results: List[dict], proteins_per_launch: int = 64):
    """thermompnn_benchmarking.py:87-121: appends {"Model", "Dataset", "ddG r2", ...} to ``results``."""
    items = _predict_dataset(model, dataset, proteins_per_launch)
    pred, targ = _pairs(items)
    column = {"Model": name, "Dataset": dataset_name}
    for met, val in regression_metrics(pred, targ).items():
        column[f"ddG {met}"] = val
    results.append(column)
    return results


def run_prediction_keep_preds(name, model, dataset_name, dataset, results: List[dict], centrality: bool = False,
                              proteins_per_launch: int = 64, out_csv: Optional[str] = None):
    """thermompnn_benchmarking.py:124-187: the same plus ``<name>_<dataset>_raw_preds.csv`` (same columns)."""
    import pandas as pd
    items = _predict_dataset(model, dataset, proteins_per_launch)
    pred, targ = _pairs(items)
    rows, k = [], 0
    for pdb, muts, outs in items:
        nb = None
        if centrality:
            coord_chain = [c for c in pdb[0].keys() if "coords" in c][0]
            nb = compute_centrality(pdb[0][coord_chain], basis_atom="CA", chain=coord_chain[-1], radius=10.0)
        for m, o in zip(muts, outs):
            row = {"WT Seq": None, "Model": name, "Dataset": dataset_name, "ddG_true": None, "ddG_pred": None, "position": None,
                   "wildtype": None, "mutation": None, "neighbors": None, "best_AA": None, "pdb": None}
            if m.ddG is not None:
                row.update(ddG_true=float(targ[k]), ddG_pred=float(pred[k]), position=m.position, wildtype=m.wildtype,
                           mutation=m.mutation, pdb=m.pdb.strip(".pdb"))
                k += 1
                if nb is not None:
                    row["neighbors"] = int(nb[m.position])
            key = m.pdb if "Megascale" not in dataset_name else m.pdb + ".pdb"           # :163-166
            if "S669" not in dataset_name:
                ws = dataset.wt_seqs            # (names containing '|' are stored as ':' in Mutation.pdb, datasets.py:134)
                row["WT Seq"] = ws[key] if key in ws else ws[key.replace(":", "|")]
            rows.append(row)
    column = {"Model": name, "Dataset": dataset_name}
    for met, val in regression_metrics(pred, targ).items():
        column[f"ddG {met}"] = val
    results.append(column)
    df = pd.DataFrame(rows, columns=["WT Seq", "Model", "Dataset", "ddG_true", "ddG_pred", "position", "wildtype", "mutation",
                                     "neighbors", "best_AA", "pdb"])
    df.to_csv(out_csv if out_csv is not None else name + "_" + dataset_name + "_raw_preds.csv")
    return results
