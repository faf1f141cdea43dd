"""Checkpoint loader: reads the reference's own files, no Lightning needed.

Mirrors what ``get_trained_model`` -> ``TransferModelPL.load_from_checkpoint`` ->
``TransferModel.__init__`` -> ``get_protein_mpnn`` do upstream
(thermompnn_benchmarking.py:78-84, train_thermompnn.py:30-32, transfer_model.py:17-73):
the Lightning checkpoint's ``state_dict`` (keys prefixed ``model.``) overrides the vanilla
ProteinMPNN weights, and ``model.prot_mpnn.*`` inside ``thermoMPNN_default.pt`` is
bit-identical to ``vanilla_model_weights/v_48_020.pt``, so one file is enough.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

LA_W = "light_attention.feature_convolution.weight"
LA_WC = "light_attention.feature_convolution.weight_center"


def _normalise(state: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    out = {}
    for k, v in state.items():
        if k.startswith("model."):
            k = k[len("model."):]
        v = np.asarray(v, dtype=np.float32)
        if k == LA_W:
            # LightAttention runs on a length-1 sequence (transfer_model.py:107-108): with kernel 9 / padding 4
            # only the centre tap ever multiplies data, and softmax over a length-1 axis is 1.
            k, v = LA_WC, np.ascontiguousarray(v[:, :, v.shape[2] // 2])
        if k.startswith("light_attention.attention_convolution"):
            continue
        out[k] = v
    return out


def load_state(path: str, vanilla_path: Optional[str] = None) -> Dict[str, np.ndarray]:
    """Load ``thermoMPNN_default.pt``-style Lightning checkpoints, bare state_dicts, or the slim ``.npz``
    fixture.  ``vanilla_path`` (a ``v_48_0xx.pt``) fills in ``prot_mpnn.*`` when the file lacks them."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith(".npz"):
        state = dict(np.load(path))
    else:
        import torch
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
        sd = ckpt.get("state_dict", ckpt.get("model_state_dict", ckpt))
        state = {k: v.detach().cpu().numpy() for k, v in sd.items() if hasattr(v, "detach")}
    state = _normalise(state)
    if vanilla_path is not None and not any(k.startswith("prot_mpnn.") for k in state):
        import torch
        van = torch.load(vanilla_path, map_location="cpu", weights_only=True)["model_state_dict"]
        state.update({"prot_mpnn." + k: v.numpy().astype(np.float32) for k, v in van.items()})
    return state


def random_state(seed: int = 0, cfg=None) -> Dict[str, np.ndarray]:
    """Random-init weights of the architecture ``cfg`` describes (``engine.ModelCfg``; default = the shipped model)."""
    from .engine import weight_table
    rng = np.random.default_rng(seed)
    state = {}
    for name, _, r, c in weight_table(cfg):
        if "norm" in name and name.endswith(".weight"):
            v = np.ones((r, c))
        elif name.endswith(".bias"):
            v = rng.normal(scale=0.02, size=(r, c))
        elif name == "ddg_out.weight":
            v = np.full((1, 1), 0.35)
        else:
            v = rng.uniform(-1, 1, size=(r, c)) * np.sqrt(6.0 / (r + c))
        state[name] = v.astype(np.float32)
    return state
