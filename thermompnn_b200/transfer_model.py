"""``TransferModel`` -- host-side mirror of the reference's model class.

Same constructor config keys, same ``forward(pdb, mutations) -> (list, None)`` contract
as ``transfer_model.py:42-121`` upstream, so it drops into ``analysis/custom_inference.py``
and ``analysis/thermompnn_benchmarking.py``.  All arithmetic happens in the CUDA library
behind the C ABI (``engine.Engine``); this class only packs inputs and hands out results.

What changes relative to the reference, deliberately:
  * one device pass produces the whole [L,21] head table; each requested mutation is a
    lookup ``T[pos, mut] - T[pos, wt]`` instead of a Python loop of tiny kernels;
  * ``ssm_tables`` takes MANY proteins at once (the reference is strictly B=1);
  * there is no CPU execution path: ``forward`` on a CPU module raises.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import engine as _engine
from .datasets import ALPHABET, Mutation
from .featurize import Packed, featurize_batch
from .weights import load_state, random_state

HIDDEN_DIM = 128
EMBED_DIM = 128
VOCAB_DIM = 21


class AttrDict(dict):
    """Minimal OmegaConf stand-in: attribute access, ``in``, attribute assignment."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    __setattr__ = dict.__setitem__

    @staticmethod
    def wrap(d):
        if isinstance(d, dict) and not isinstance(d, AttrDict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        return d


def default_config(thermompnn_dir: str = ".") -> AttrDict:
    """The inline config of ``analysis/custom_inference.py:31-49`` plus ``platform.thermompnn_dir``."""
    return AttrDict.wrap({
        "model": {"hidden_dims": [64, 32], "subtract_mut": True, "num_final_layers": 2, "freeze_weights": True,
                  "load_pretrained": True, "lightattn": True},
        "platform": {"thermompnn_dir": thermompnn_dir},
    })


class TransferModel(nn.Module):
    def __init__(self, cfg=None, state: Optional[Dict[str, np.ndarray]] = None):
        super().__init__()
        cfg = AttrDict.wrap(cfg) if isinstance(cfg, dict) else (cfg if cfg is not None else default_config())
        self.cfg = cfg
        self.hidden_dims = list(cfg.model.hidden_dims)
        self.subtract_mut = cfg.model.subtract_mut
        self.num_final_layers = cfg.model.num_final_layers
        self.lightattn = cfg.model.lightattn if "lightattn" in cfg.model else False
        if self.hidden_dims != [64, 32] or self.num_final_layers != 2 or not self.lightattn:
            raise NotImplementedError(
                "the CUDA path implements the shipped thermoMPNN_default architecture: hidden_dims=[64,32], "
                "num_final_layers=2, lightattn=True (analysis/custom_inference.py:31-49)")
        if state is None:
            state = self._initial_state(cfg)
        # The whole model is one flat fp32 blob in the C ABI's raw layout; keeping it as a frozen
        # Parameter gives .cuda()/.to()/.eval() and ``next(self.parameters()).device`` their usual meaning.
        self.weight_blob = nn.Parameter(torch.from_numpy(_engine.pack_weight_blob(state)), requires_grad=False)
        self._engine: Optional[_engine.Engine] = None
        self._engine_key = None

    @staticmethod
    def _initial_state(cfg):
        # get_protein_mpnn (transfer_model.py:17-37): vanilla weights from <thermompnn_dir>/vanilla_model_weights
        van = None
        if "platform" in cfg and "thermompnn_dir" in cfg.platform:
            p = os.path.join(cfg.platform.thermompnn_dir, "vanilla_model_weights", "v_48_020.pt")
            van = p if os.path.exists(p) else None
        state = random_state(0)
        if van is not None and cfg.model.load_pretrained:
            import torch as _t
            sd = _t.load(van, map_location="cpu", weights_only=True)["model_state_dict"]
            state.update({"prot_mpnn." + k: v.numpy().astype(np.float32) for k, v in sd.items()})
        return state

    # ---- checkpoint interface ---------------------------------------------
    def load_reference_state(self, state: Dict[str, np.ndarray]):
        """Replace all weights from a reference state_dict (``model.`` prefix optional)."""
        from .weights import _normalise
        blob = torch.from_numpy(_engine.pack_weight_blob(_normalise(state)))
        with torch.no_grad():
            self.weight_blob.copy_(blob.to(self.weight_blob.device))
        self._engine_key = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._engine_key = None   # weights moved: re-upload on next use
        return out

    def _get_engine(self) -> _engine.Engine:
        dev = self.weight_blob.device
        if dev.type != "cuda":
            raise RuntimeError("thermompnn_b200.TransferModel runs only on a CUDA device (B200): call model.cuda() "
                               "first; there is no CPU path")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (idx, self.weight_blob._version)
        if self._engine is None or self._engine.device != idx:
            self._engine = _engine.Engine(idx)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.load_weights(self.weight_blob.detach().cpu().numpy())
            self._engine_key = key
        return self._engine

    # ---- batched entry (new capability) ------------------------------------
    def head_tables(self, pdbs: Sequence[dict]) -> List[np.ndarray]:
        """[L_b, 21] head outputs T for each protein (host numpy).  ddG(i, wt->a) = T[i,a] - T[i,wt]."""
        packed = featurize_batch(pdbs)
        T = self._get_engine().run(packed.offsets, packed.X, packed.S, packed.residue_idx, packed.chain_enc, packed.mask)
        return [T[packed.offsets[b]:packed.offsets[b + 1]] for b in range(packed.n_proteins)]

    def ssm_tables(self, pdbs: Sequence[dict]) -> List[np.ndarray]:
        """Site-saturation ddG for every protein: [L_b, 20] in ``ALPHABET[:20]`` column order, wt->wt exactly 0,
        NaN rows for gap residues (for which the reference's callers emit ``None``, SSM.py:26-27)."""
        out = []
        for pdb, T in zip(pdbs, self.head_tables(pdbs)):
            seq = pdb["seq"]
            wt = np.asarray([ALPHABET.index(a) if a != "-" else 20 for a in seq])
            ddg = T[:, :20] - T[np.arange(len(seq)), wt][:, None]
            ddg[np.asarray([a == "-" for a in seq])] = np.nan
            out.append(ddg)
        return out

    # ---- reference-compatible entry -----------------------------------------
    def forward(self, pdb, mutations, tied_feat=True):
        """Same contract as upstream ``TransferModel.forward`` (transfer_model.py:75-121): uses ``pdb[0]``,
        returns ``([None | {"ddG": tensor[1]}, ...], None)`` index-aligned with ``mutations``."""
        device = self.weight_blob.device
        T = self.head_tables([pdb[0]])[0]
        L = T.shape[0]
        idx = [(i, m) for i, m in enumerate(mutations) if m is not None]
        out: List[Optional[dict]] = [None] * len(mutations)
        if not idx:
            return out, None
        pos = np.asarray([m.position for _, m in idx], dtype=np.int64)
        if (pos < 0).any() or (pos >= L).any():
            raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (int(pos[(pos < 0) | (pos >= L)][0]), L))
        mut = np.asarray([ALPHABET.index(m.mutation) for _, m in idx], dtype=np.int64)   # ValueError like upstream
        wt = np.asarray([ALPHABET.index(m.wildtype) for _, m in idx], dtype=np.int64)
        vals = T[pos, mut] - T[pos, wt] if self.subtract_mut else T[pos, mut]
        vals_t = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float32)).to(device)   # one H2D for all
        for k, (i, _) in enumerate(idx):
            out[i] = {"ddG": vals_t[k:k + 1]}
        return out, None


def get_trained_model(model_name: str, config=None, checkpt_dir: str = "models/", override_custom: bool = False):
    """Loader with the reference's signature (thermompnn_benchmarking.py:78-84): returns a
    ``TransferModel`` carrying the checkpoint's weights.  Accepts the stock Lightning ``.pt``,
    a bare state_dict ``.pt`` or the slim ``.npz``."""
    path = model_name if override_custom else os.path.join(checkpt_dir, model_name)
    cfg = config if config is not None else default_config()
    if isinstance(cfg, dict) and not isinstance(cfg, AttrDict):
        cfg = AttrDict.wrap(cfg)
    base = default_config()
    if "model" not in cfg:
        cfg["model"] = base.model
    van = None
    if "platform" in cfg and "thermompnn_dir" in cfg.platform:
        p = os.path.join(cfg.platform.thermompnn_dir, "vanilla_model_weights", "v_48_020.pt")
        van = p if os.path.exists(p) else None
    state = load_state(path, vanilla_path=van)
    return TransferModel(cfg, state=state)
