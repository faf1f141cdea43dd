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


def _cfg_get(cfg, key, default):
    return cfg[key] if key in cfg else default


def vanilla_state(cfg, version: str = "v_48_020.pt") -> Dict[str, np.ndarray]:
    """``get_protein_mpnn`` (transfer_model.py:17-37): the vanilla ProteinMPNN checkpoint under
    ``<thermompnn_dir>/vanilla_model_weights``; like upstream, a missing file is an error (no silent random weights)."""
    path = os.path.join(cfg.platform.thermompnn_dir, "vanilla_model_weights", version)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    if int(ckpt["num_edges"]) != 48:
        raise NotImplementedError(f"{version}: k_neighbors = {ckpt['num_edges']}; the CUDA path is built for 48 (v_48_*)")
    if not cfg.model.load_pretrained:
        return {}
    return {"prot_mpnn." + k: v.numpy().astype(np.float32) for k, v in ckpt["model_state_dict"].items()}


class _EngineModule(nn.Module):
    """Shared plumbing: the whole model is ONE flat fp32 blob in the C ABI's raw layout, kept as a frozen Parameter
    so that ``.cuda()/.to()/.eval()`` and ``next(self.parameters()).device`` mean what the reference's callers expect."""

    def _init_blob(self, state: Dict[str, np.ndarray], mcfg: _engine.ModelCfg):
        self._mcfg = mcfg
        self.weight_blob = nn.Parameter(torch.from_numpy(_engine.pack_weight_blob(state, mcfg)), requires_grad=False)
        self._engine: Optional[_engine.Engine] = None
        self._engine_key = None

    def load_reference_state(self, state: Dict[str, np.ndarray]):
        """Replace all weights from a reference state_dict (``model.`` prefix optional)."""
        from .weights import _normalise
        blob = torch.from_numpy(_engine.pack_weight_blob(_normalise(state), self._mcfg))
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
            raise RuntimeError("thermompnn_b200 models run only on a CUDA device (B200): call model.cuda() "
                               "first; there is no CPU path")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (idx, self.weight_blob._version)
        if self._engine is None or self._engine.device != idx:
            self._engine = _engine.Engine(idx)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.load_weights(self.weight_blob.detach().cpu().numpy(), self._mcfg)
            self._engine_key = key
        return self._engine

    @staticmethod
    def _positions(mutations, L):
        """Indices and positions of the non-None mutations; positions follow tensor indexing (-L..L-1), as upstream's
        ``mpnn_hid[0][mut.position]`` does (transfer_model.py:100)."""
        idx = [i for i, m in enumerate(mutations) if m is not None]
        pos = np.asarray([mutations[i].position for i in idx], dtype=np.int64)
        bad = (pos < -L) | (pos >= L)
        if bad.any():
            raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (int(pos[bad][0]), L))
        return idx, np.where(pos < 0, pos + L, pos)


class TransferModel(_EngineModule):
    def __init__(self, cfg=None, state: Optional[Dict[str, np.ndarray]] = None):
        super().__init__()
        cfg = AttrDict.wrap(cfg) if isinstance(cfg, dict) else (cfg if cfg is not None else default_config())
        self.cfg = cfg
        self.hidden_dims = list(cfg.model.hidden_dims)
        self.subtract_mut = cfg.model.subtract_mut
        self.num_final_layers = cfg.model.num_final_layers
        self.lightattn = cfg.model.lightattn if "lightattn" in cfg.model else False
        if not 0 <= int(self.num_final_layers) <= 3:
            raise ValueError("num_final_layers must be 0..3 (ProteinMPNN has three decoder layers)")
        mcfg = _engine.ModelCfg.make(self.num_final_layers, self.lightattn, self.hidden_dims, True)
        if state is None:
            state = self._initial_state(cfg, mcfg)
        self._init_blob(state, mcfg)

    @staticmethod
    def _initial_state(cfg, mcfg):
        """What ``TransferModel.__init__`` builds upstream (transfer_model.py:42-73): vanilla ProteinMPNN weights
        (``get_protein_mpnn``) and a freshly initialised head."""
        state = random_state(0, mcfg)
        state.update(vanilla_state(cfg))
        return state

    # ---- batched entries (new capability: the reference is strictly B = 1) ------------------
    def head_tables(self, pdbs: Sequence[dict]) -> List[np.ndarray]:
        """[L_b, 21] head outputs T for each protein (host numpy).  ddG(i, wt->a) = T[i,a] - T[i,wt]."""
        return self.head_tables_packed(featurize_batch(pdbs))

    def head_tables_packed(self, packed: Packed) -> List[np.ndarray]:
        """The same from an already packed batch (``featurize.featurize_batch``): packing costs ~1.4 us per residue on the
        host with upstream's nested-list coordinates (0.1 us with numpy arrays) -- more than the GPU pass on large batches --
        so callers that scan one set of structures repeatedly pack once."""
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

    def _lookup(self, T, mutations):
        idx, pos = self._positions(mutations, T.shape[0])
        mut = np.asarray([ALPHABET.index(mutations[i].mutation) for i in idx], dtype=np.int64)   # ValueError like upstream
        wt = np.asarray([ALPHABET.index(mutations[i].wildtype) for i in idx], dtype=np.int64)
        vals = T[pos, mut] - T[pos, wt] if self.subtract_mut else T[pos, mut]
        return idx, np.ascontiguousarray(vals, dtype=np.float32)

    def predict_batch(self, batch):
        """Listed-mutation batches (what the reference's dataset loops feed one protein at a time,
        thermompnn_benchmarking.py:97-99, datasets.py:133-164): ``batch`` = [(pdb, mutations), ...] with ``pdb`` a dict
        or the reference's one-element list.  ONE ragged launch for all proteins, one host->device copy for all
        values; returns one ``forward``-style list per protein ([None | {"ddG": tensor[1]}, ...])."""
        pdbs = [p[0] if isinstance(p, (list, tuple)) else p for p, _ in batch]
        tables = self.head_tables(pdbs)
        looked = [self._lookup(T, muts) for T, (_, muts) in zip(tables, batch)]
        flat = np.concatenate([v for _, v in looked]) if looked else np.zeros(0, np.float32)
        flat_t = torch.from_numpy(flat).to(self.weight_blob.device)
        views = flat_t.split(1) if flat_t.numel() else ()     # shape-[1] views of ONE device tensor (one C++ call, not one slice per mutation)
        outs, o = [], 0
        for (idx, v), (_, muts) in zip(looked, batch):
            out: List[Optional[dict]] = [None] * len(muts)
            for k, i in enumerate(idx):
                out[i] = {"ddG": views[o + k]}
            o += len(idx)
            outs.append(out)
        return outs

    # ---- reference-compatible entry -----------------------------------------
    def forward(self, pdb, mutations, tied_feat=True):
        """Same contract as upstream ``TransferModel.forward`` (transfer_model.py:75-121): uses ``pdb[0]``,
        returns ``([None | {"ddG": tensor[1]}, ...], None)`` index-aligned with ``mutations``."""
        return self.predict_batch([(pdb[0], mutations)])[0], None


class ProteinMPNNBaseline(_EngineModule):
    """ProteinMPNN's own log-probabilities as a ddG proxy: the second model of the reference's benchmarking script
    (analysis/thermompnn_benchmarking.py:38-65).  ``forward`` returns ``-log_probs[pos, mut]`` as "ddG" and
    ``+log_probs[pos, mut]`` as "dTm", plus the [1, L, 21] log_probs tensor itself."""

    def __init__(self, cfg=None, version: str = "v_48_020.pt", state: Optional[Dict[str, np.ndarray]] = None):
        super().__init__()
        cfg = AttrDict.wrap(cfg) if isinstance(cfg, dict) else (cfg if cfg is not None else default_config())
        self.cfg = cfg
        mcfg = _engine.ModelCfg.make(has_head=False)
        if state is None:
            state = random_state(0, mcfg)
            state.update(vanilla_state(cfg, version))
        self._init_blob(state, mcfg)

    def log_prob_tables(self, pdbs: Sequence[dict]) -> List[np.ndarray]:
        packed = featurize_batch(pdbs)
        _, lp = self._get_engine().run(packed.offsets, packed.X, packed.S, packed.residue_idx, packed.chain_enc, packed.mask,
                                       want_table=False, log_probs=True)
        return [lp[packed.offsets[b]:packed.offsets[b + 1]] for b in range(packed.n_proteins)]

    def unconditional_probs(self, pdbs: Sequence[dict]) -> List[np.ndarray]:
        """``ProteinMPNN.unconditional_probs`` (protein_mpnn_utils.py:1557-1587) for each protein: [L_b, 21] log-probabilities
        from the structure alone -- the decoder sees no sequence and the neighbours' encoder states."""
        packed = featurize_batch(pdbs)
        _, lp = self._get_engine().run(packed.offsets, packed.X, packed.S, packed.residue_idx, packed.chain_enc, packed.mask,
                                       want_table=False, log_probs=True, flags=_engine.UNCONDITIONAL)
        return [lp[packed.offsets[b]:packed.offsets[b + 1]] for b in range(packed.n_proteins)]

    def forward(self, pdb, mutations, tied_feat=True):
        lp = self.log_prob_tables([pdb[0]])[0]
        log_probs = torch.from_numpy(lp).to(self.weight_blob.device)        # one H2D; per-mutation results are views
        idx, pos = self._positions(mutations, lp.shape[0])
        aa = np.asarray([ALPHABET.index(mutations[i].mutation) for i in idx], dtype=np.int64)
        out: List[Optional[dict]] = [None] * len(mutations)
        if idx:
            flat = log_probs.reshape(-1)
            at = torch.from_numpy(pos * 21 + aa).to(flat.device)
            pred = flat[at]
            neg = -pred
            for k, i in enumerate(idx):
                out[i] = {"ddG": neg[k:k + 1], "dTm": pred[k:k + 1]}
        return out, log_probs.unsqueeze(0)


def get_trained_model(model_name: str, config=None, checkpt_dir: str = "models/", override_custom: bool = False):
    """Loader with the reference's signature and path rule (thermompnn_benchmarking.py:78-84): ``override_custom``
    takes ``model_name`` as the path, otherwise it is ``<config.platform.thermompnn_dir>/<checkpt_dir>/<model_name>``.
    Returns a ``TransferModel`` carrying the checkpoint's weights.  Accepts the stock Lightning ``.pt``,
    a bare state_dict ``.pt`` or the slim ``.npz``."""
    cfg = config if config is not None else default_config()
    if isinstance(cfg, dict) and not isinstance(cfg, AttrDict):
        cfg = AttrDict.wrap(cfg)
    if "model" not in cfg:
        cfg["model"] = default_config().model
    if override_custom:
        path = model_name
    else:
        root = cfg.platform.thermompnn_dir if "platform" in cfg and "thermompnn_dir" in cfg.platform else ""
        path = os.path.join(os.path.join(root, checkpt_dir), model_name)
    van = None
    if "platform" in cfg and "thermompnn_dir" in cfg.platform:
        p = os.path.join(cfg.platform.thermompnn_dir, "vanilla_model_weights", "v_48_020.pt")
        van = p if os.path.exists(p) else None
    state = load_state(path, vanilla_path=van)
    return TransferModel(cfg, state=state)
