#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tools/make_golden.py

Writes
  tests/golden/thermompnn_default_slim.npz   reference checkpoint tensors (fp32), LightAttention
                                             feature conv reduced to its centre tap, attention conv dropped
  tests/golden/2OCJ.pdb                      the reference's example structure (data, not source)
  tests/golden/ref_2OCJ_A_published.csv      the reference's shipped known-answer file
  tests/golden/parse_2OCJ_{A,AB}.npz         alt_parse_PDB output (seq, coords) for the parser test
  tests/golden/case_<name>.npz               inputs + reference outputs (ddG [L,20], per-layer traces)
"""
import os
import shutil
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from thermompnn_b200.synth import synth, synth_multichain  # noqa: E402


class Cfg(dict):
    """OmegaConf stand-in: attribute access, `in`, attribute assignment."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
    __setattr__ = dict.__setitem__


class Mut:
    def __init__(self, position, wildtype, mutation):
        self.position, self.wildtype, self.mutation = position, wildtype, mutation


def load_reference():
    from transfer_model import TransferModel
    cfg = Cfg(model=Cfg(hidden_dims=[64, 32], subtract_mut=True, num_final_layers=2, freeze_weights=True,
                        load_pretrained=True, lightattn=True),
              platform=Cfg(thermompnn_dir=REF))
    m = TransferModel(cfg)
    sd = torch.load(os.path.join(REF, "models/thermoMPNN_default.pt"), map_location="cpu")["state_dict"]
    m.load_state_dict({k[6:]: v for k, v in sd.items()}, strict=True)
    m.eval()
    return m, sd


def dump_weights(sd):
    out = {}
    for k, v in sd.items():
        k = k[6:]
        if k == "light_attention.feature_convolution.weight":
            out["light_attention.feature_convolution.weight_center"] = v[:, :, 4].contiguous().numpy()
        elif k.startswith("light_attention.attention_convolution"):
            continue
        else:
            out[k] = v.numpy()
    np.savez(os.path.join(OUT, "thermompnn_default_slim.npz"), **out)
    print("weights:", sum(v.size for v in out.values()), "params")


def pdb_to_arrays(pdb):
    """Flatten an alt_parse_PDB dict into npz-able arrays."""
    d = {"seq": np.array(pdb["seq"]), "name": np.array(pdb["name"]),
         "chains": np.array([k[-1:] for k in pdb if k[:10] == "seq_chain_"])}
    for k in pdb:
        if k[:10] == "seq_chain_":
            ch = k[-1:]
            d[f"seq_{ch}"] = np.array(pdb[k])
            co = pdb[f"coords_chain_{ch}"]
            d[f"xyz_{ch}"] = np.stack([np.asarray(co[f"{a}_chain_{ch}"], dtype=np.float64)
                                       for a in ("N", "CA", "C", "O")], 1)
    return d


def run_case(model, name, pdb):
    trace = {}
    hooks = []
    pm = model.prot_mpnn

    def grab(key):
        def fn(mod, inp, out):
            trace[key] = out
        return fn
    hooks.append(pm.features.register_forward_hook(grab("features")))
    hooks.append(pm.W_e.register_forward_hook(grab("h_E0")))
    for l in range(3):
        hooks.append(pm.encoder_layers[l].register_forward_hook(grab(f"enc{l}")))
        hooks.append(pm.decoder_layers[l].register_forward_hook(grab(f"dec{l}")))
    seq = pdb["seq"]
    muts = []
    for i, wt in enumerate(seq):
        if wt == "-":
            muts.append(None)
        else:
            muts.extend(Mut(i, wt, a) for a in "ACDEFGHIKLMNPQRSTVWY")
    with torch.no_grad():
        pred, _ = model([pdb], muts)
    for h in hooks:
        h.remove()
    L = len(seq)
    ddg = np.full((L, 20), np.nan, dtype=np.float32)
    for m, p in zip(muts, pred):
        if m is not None:
            ddg[m.position, "ACDEFGHIKLMNPQRSTVWY".index(m.mutation)] = float(p["ddG"][0])
    E, E_idx = trace["features"]
    d = pdb_to_arrays(pdb)
    d.update(ddG=ddg, E_idx=E_idx[0].numpy().astype(np.int32),
             E_head=E[0, :6].numpy(), h_E0_head=trace["h_E0"][0, :6].numpy())
    for l in range(3):
        hv, he = trace[f"enc{l}"]
        d[f"enc{l}_h_V"] = hv[0].numpy()
        d[f"enc{l}_h_E_head"] = he[0, :6].numpy()
        d[f"dec{l}_h_V"] = trace[f"dec{l}"][0].numpy()
    np.savez_compressed(os.path.join(OUT, f"case_{name}.npz"), **d)
    print(f"case {name}: L={L} K={E_idx.shape[-1]} ddG range [{np.nanmin(ddg):.3f}, {np.nanmax(ddg):.3f}]")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    model, sd = load_reference()
    dump_weights(sd)
    shutil.copyfile(os.path.join(REF, "examples/2OCJ.pdb"), os.path.join(OUT, "2OCJ.pdb"))
    shutil.copyfile(os.path.join(REF, "examples/ThermoMPNN_inference_2OCJ.csv"),
                    os.path.join(OUT, "ref_2OCJ_A_published.csv"))
    from protein_mpnn_utils import alt_parse_PDB
    pdb_a = alt_parse_PDB(os.path.join(REF, "examples/2OCJ.pdb"), "A")[0]
    pdb_ab = alt_parse_PDB(os.path.join(REF, "examples/2OCJ.pdb"), "AB")[0]
    np.savez_compressed(os.path.join(OUT, "parse_2OCJ_A.npz"), **pdb_to_arrays(pdb_a))
    np.savez_compressed(os.path.join(OUT, "parse_2OCJ_AB.npz"), **pdb_to_arrays(pdb_ab))
    run_case(model, "2OCJ_A", pdb_a)
    run_case(model, "2OCJ_AB", pdb_ab)
    run_case(model, "synth300", synth(300, seed=0))                       # BASELINE config 2
    run_case(model, "synth40", synth(40, seed=7))                         # L < 48  =>  K = L
    run_case(model, "synth120_gaps", synth(120, seed=11, gaps=[0, 17, 18, 19, 64, 119]))
    run_case(model, "synth_multichain", synth_multichain([60, 50], seed=5))
    run_case(model, "synth49", synth(49, seed=3))                         # K = 48 of 49


if __name__ == "__main__":
    main()
