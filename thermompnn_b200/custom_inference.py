"""Single-PDB site-saturation scan: the entry ``analysis/custom_inference.py`` provides upstream (:28-126), on the
CUDA path.  Same flags (``--pdb --chain --model_path --out_dir``), same CSV name and columns
(``Model, Dataset, ddG_pred, position, wildtype, mutation, pdb, chain``), same row order; the table is written
in one shot instead of upstream's per-cell ``DataFrame.loc`` writes (6.8 s for 3 880 rows there).

    python -m thermompnn_b200.custom_inference --pdb examples/2OCJ.pdb --chain A --model_path models/thermoMPNN_default.pt
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from .datasets import ALPHABET
from .pdb import alt_parse_PDB
from .ssm import get_ssm_mutations, to_mutation_objects
from .transfer_model import default_config, get_trained_model


def first_chain(pdb_path: str) -> str:
    with open(pdb_path, "rb") as f:
        for ln in f:
            s = ln.decode("utf-8", "ignore")
            if s[:4] == "ATOM":
                return s[21:22]
    raise ValueError(f"{pdb_path}: no ATOM records")


def run(pdb_path: str, chain: str, model, model_name: str = "ThermoMPNN"):
    """Returns a pandas DataFrame identical in shape/order to upstream's ``raw_pred_df``."""
    import pandas as pd
    pdb_id = os.path.basename(pdb_path).rstrip(".pdb")          # (sic) upstream uses rstrip, custom_inference.py:56
    mut_pdb = alt_parse_PDB(pdb_path, chain)
    muts = to_mutation_objects(get_ssm_mutations(mut_pdb[0]), mut_pdb[0]["name"])
    with torch.no_grad():
        pred, _ = model(mut_pdb, muts)
    keep = [(m, p) for m, p in zip(muts, pred) if m is not None]
    ddg = torch.cat([p["ddG"] for _, p in keep]).cpu().numpy() if keep else np.zeros(0, np.float32)   # ONE device->host copy
    df = pd.DataFrame({
        "Model": model_name, "Dataset": pdb_id, "ddG_pred": ddg.astype(object),
        "position": np.asarray([m.position for m, _ in keep], dtype=object),
        "wildtype": [m.wildtype for m, _ in keep], "mutation": [m.mutation for m, _ in keep],
        "pdb": [m.pdb.strip(".pdb") for m, _ in keep], "chain": chain,
    }, columns=["Model", "Dataset", "ddG_pred", "position", "wildtype", "mutation", "pdb", "chain"])
    return df, pdb_id


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdb", type=str, default="", help="Input PDB to use for custom inference")
    ap.add_argument("--chain", type=str, default="A", help="Chain in input PDB to use.")
    ap.add_argument("--model_path", type=str, default="", help="filepath to model to use for inference")
    ap.add_argument("--out_dir", type=str, default="./", help="Output directory in which to save predictions.")
    args = ap.parse_args(argv)
    cfg = default_config(os.path.dirname(os.path.dirname(os.path.abspath(args.model_path))))
    model = get_trained_model(args.model_path, cfg, override_custom=True).eval().cuda()
    chain = args.chain if len(args.chain) >= 1 else first_chain(args.pdb)
    df, pdb_id = run(args.pdb, chain, model)
    print(df)
    out_dir = os.getcwd() if args.out_dir == "./" else args.out_dir
    assert os.path.isdir(out_dir), f"{out_dir} is not a valid directory."
    csv_file = os.path.join(out_dir, "ThermoMPNN_inference_%s.csv" % pdb_id)
    df.to_csv(csv_file)
    print(f"Saved ThermoMPNN output to {csv_file}")
    return csv_file


if __name__ == "__main__":
    main()
