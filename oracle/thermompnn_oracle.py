"""CPU oracle for the ThermoMPNN site-saturation ddG hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``thermompnn_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the timed CPU baseline -- never as the product path.

What it is: an independent, functional (no nn.Module) torch-CPU restatement of
the reference algorithm, op for op and in the reference's association order, so
that fp32 results track the reference to ~1e-6.  Each function cites the
reference lines (relative to the upstream repo root) it follows.

Parity pin: ``tests/golden/*.npz`` were produced by importing the UNMODIFIED
reference in the build container (``tools/make_golden.py``); ``tests/test_oracle.py``
checks this oracle against every one of them, and against the reference's only
shipped known-answer file (``examples/ThermoMPNN_inference_2OCJ.csv``, kept as
``tests/golden/ref_2OCJ_A_published.csv``; it pins to ~1e-3 only because it was
produced on the authors' GPU).

Weights are a flat ``dict[str, torch.Tensor]`` keyed by the reference checkpoint
names with the Lightning ``model.`` prefix stripped, e.g.
``prot_mpnn.encoder_layers.0.W1.weight``.  The LightAttention feature
convolution may be given either in full (``light_attention.feature_convolution.weight``
[384,384,9]) or as its centre tap (``...weight_center`` [384,384]).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

ALPHABET = "ACDEFGHIKLMNPQRSTVWYX"  # transfer_model.py:11
TOP_K = 48                           # checkpoint['num_edges'], transfer_model.py:26
NUM_RBF = 16
MAX_REL = 32                         # protein_mpnn_utils.py:897
SCALE = 30.0                         # EncLayer/DecLayer scale, protein_mpnn_utils.py:795,843

# Atom-pair order of the 25 RBF blocks, protein_mpnn_utils.py:1142-1167.
# (A, B) means atom A of residue i against atom B of neighbour j.
PAIRS = [("Ca", "Ca"), ("N", "N"), ("C", "C"), ("O", "O"), ("Cb", "Cb"),
         ("Ca", "N"), ("Ca", "C"), ("Ca", "O"), ("Ca", "Cb"), ("N", "C"),
         ("N", "O"), ("N", "Cb"), ("Cb", "C"), ("Cb", "O"), ("O", "C"),
         ("N", "Ca"), ("C", "Ca"), ("O", "Ca"), ("Cb", "Ca"), ("C", "N"),
         ("O", "N"), ("Cb", "N"), ("C", "Cb"), ("O", "Cb"), ("C", "O")]


# --------------------------------------------------------------------------
# host featurisation
# --------------------------------------------------------------------------
def featurize(pdb: dict):
    """Pack one parsed protein into flat arrays.

    Follows ``tied_featurize`` (protein_mpnn_utils.py:353-605) for the six
    tensors ``TransferModel.forward`` consumes (transfer_model.py:78-81):
    chain order = dict order of the ``seq_chain_*`` keys (:387); '-' -> 'X'
    (:454); atom order N, CA, C, O (:465-467); chain_encoding is 1-based (:471);
    residue_idx = 100*(c-1) + global position (:473); mask = all 12 coords
    finite (:576); NaN coords zeroed (:577).
    """
    letters = [k[-1:] for k in pdb if k[:10] == "seq_chain_"]
    xs, seqs, chain_enc, res_idx = [], [], [], []
    pos = 0
    for c, letter in enumerate(letters, start=1):
        seq = "".join(a if a != "-" else "X" for a in pdb[f"seq_chain_{letter}"])
        co = pdb[f"coords_chain_{letter}"]
        x = np.stack([np.asarray(co[f"{a}_chain_{letter}"], dtype=np.float64)
                      for a in ("N", "CA", "C", "O")], 1)
        n = len(seq)
        xs.append(x)
        seqs.append(seq)
        chain_enc.append(np.full(n, c, dtype=np.int64))
        res_idx.append(100 * (c - 1) + np.arange(pos, pos + n, dtype=np.int64))
        pos += n
    X = np.concatenate(xs, 0)
    mask = np.isfinite(X.sum((1, 2))).astype(np.float32)
    X = np.where(np.isnan(X), 0.0, X).astype(np.float32)
    S = np.asarray([ALPHABET.index(a) for a in "".join(seqs)], dtype=np.int64)
    return X, S, mask, np.concatenate(res_idx), np.concatenate(chain_enc)


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _lin(w, name, x):
    b = w.get(name + ".bias")
    return F.linear(x, w[name + ".weight"], b)


def _ln(w, name, x):
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], 1e-5)


def _gather_nodes(h, E_idx):
    """[L,C] at [L,K] -> [L,K,C]  (gather_nodes, protein_mpnn_utils.py:770-778)."""
    return h[E_idx]


# --------------------------------------------------------------------------
# graph + edge features
# --------------------------------------------------------------------------
def knn(Ca, mask, top_k=TOP_K):
    """k-NN over masked CA distances (ProteinFeatures._dist, :1101-1109).

    D[i,j] = m_i m_j sqrt(sum (Ca_j-Ca_i)^2 + 1e-6); masked pairs are pushed to
    the row maximum; the K = min(top_k, L) smallest are returned ascending.
    """
    m2 = mask[None, :] * mask[:, None]
    dX = Ca[None, :, :] - Ca[:, None, :]
    D = m2 * torch.sqrt(torch.sum(dX ** 2, 2) + 1e-6)
    D_max, _ = torch.max(D, -1, keepdim=True)
    D_adj = D + (1.0 - m2) * D_max
    return torch.topk(D_adj, int(min(top_k, Ca.shape[0])), dim=-1, largest=False)


def rbf(D):
    """16 Gaussian radial basis functions (ProteinFeatures._rbf, :1111-1119)."""
    mu = torch.linspace(2.0, 22.0, NUM_RBF, dtype=torch.float32).to(D.dtype)
    sigma = (22.0 - 2.0) / NUM_RBF
    return torch.exp(-(((D[..., None] - mu) / sigma) ** 2))


def backbone_atoms(X):
    """N, Ca, C, O and the virtual Cb (ProteinFeatures.forward, :1131-1138)."""
    b = X[:, 1] - X[:, 0]
    c = X[:, 2] - X[:, 1]
    a = torch.cross(b, c, dim=-1)
    Cb = -0.58273431 * a + 0.56802827 * b - 0.54067466 * c + X[:, 1]
    return {"N": X[:, 0], "Ca": X[:, 1], "C": X[:, 2], "O": X[:, 3], "Cb": Cb}


def edge_features(w, X, mask, residue_idx, chain_enc, top_k=TOP_K):
    """ProteinFeatures.forward (:1127-1180): returns E [L,K,128], E_idx [L,K]."""
    p = "prot_mpnn.features."
    atoms = backbone_atoms(X)
    D_nb, E_idx = knn(atoms["Ca"], mask, top_k)
    blocks = [rbf(D_nb)]                                   # pair 0 uses the masked top-k value (:1142)
    for A, B in PAIRS[1:]:
        D_ab = torch.sqrt(torch.sum((atoms[A][:, None, :] - atoms[B][None, :, :]) ** 2, -1) + 1e-6)
        blocks.append(rbf(torch.gather(D_ab, 1, E_idx)))   # _get_rbf (:1121-1125)
    rbf_all = torch.cat(blocks, -1)
    offset = torch.gather(residue_idx[:, None] - residue_idx[None, :], 1, E_idx)      # :1170-1171
    same = torch.gather((chain_enc[:, None] - chain_enc[None, :] == 0).long(), 1, E_idx)  # :1173-1175
    d = torch.clip(offset + MAX_REL, 0, 2 * MAX_REL) * same + (1 - same) * (2 * MAX_REL + 1)  # :904-905
    e_pos = F.linear(F.one_hot(d, 2 * MAX_REL + 2).to(X.dtype),
                     w[p + "embeddings.linear.weight"], w[p + "embeddings.linear.bias"])  # :906-907
    E = F.linear(torch.cat((e_pos, rbf_all), -1), w[p + "edge_embedding.weight"])         # :1177-1178
    return _ln(w, p + "norm_edges", E), E_idx                                              # :1179


# --------------------------------------------------------------------------
# message passing
# --------------------------------------------------------------------------
def _ffn(w, p, h):
    """PositionWiseFeedForward (:883-893)."""
    return _lin(w, p + "dense.W_out", F.gelu(_lin(w, p + "dense.W_in", h)))


def enc_layer(w, l, h_V, h_E, E_idx, mask, mask_attend):
    """EncLayer.forward (:816-839); dropout is identity in eval."""
    p = f"prot_mpnn.encoder_layers.{l}."
    K = E_idx.shape[1]

    def edge_in(hv):
        return torch.cat([hv[:, None, :].expand(-1, K, -1), h_E, _gather_nodes(hv, E_idx)], -1)

    msg = _lin(w, p + "W3", F.gelu(_lin(w, p + "W2", F.gelu(_lin(w, p + "W1", edge_in(h_V))))))
    msg = mask_attend[..., None] * msg
    h_V = _ln(w, p + "norm1", h_V + torch.sum(msg, -2) / SCALE)
    h_V = _ln(w, p + "norm2", h_V + _ffn(w, p, h_V))
    h_V = mask[:, None] * h_V
    msg = _lin(w, p + "W13", F.gelu(_lin(w, p + "W12", F.gelu(_lin(w, p + "W11", edge_in(h_V))))))
    h_E = _ln(w, p + "norm3", h_E + msg)
    return h_V, h_E


def dec_layer(w, l, h_V, h_ESV, mask):
    """DecLayer.forward (:859-880), called without mask_attend (:1272)."""
    p = f"prot_mpnn.decoder_layers.{l}."
    K = h_ESV.shape[1]
    h_EV = torch.cat([h_V[:, None, :].expand(-1, K, -1), h_ESV], -1)
    msg = _lin(w, p + "W3", F.gelu(_lin(w, p + "W2", F.gelu(_lin(w, p + "W1", h_EV)))))
    h_V = _ln(w, p + "norm1", h_V + torch.sum(msg, -2) / SCALE)
    h_V = _ln(w, p + "norm2", h_V + _ffn(w, p, h_V))
    return mask[:, None] * h_V


def mpnn_forward(w, X, S, mask, residue_idx, chain_enc, top_k=TOP_K, trace: Optional[dict] = None):
    """ProteinMPNN.forward (:1222-1277) with the decoding-order mask already
    overwritten by ones (:1259) -- i.e. mask_bw = mask_i, mask_fw = 0, so the
    ``h_EXV_encoder_fw`` term (:1264-1267) vanishes and :1242-1257 are dead code.

    Returns ([dec3, dec2, dec1], h_S) like ``list(reversed(all_hidden)), h_S``.
    """
    E, E_idx = edge_features(w, X, mask, residue_idx, chain_enc, top_k)
    h_E = _lin(w, "prot_mpnn.W_e", E)                       # :1229
    h_V = torch.zeros((E.shape[0], E.shape[-1]), dtype=E.dtype)   # :1228
    mask_attend = mask[:, None] * mask[E_idx]                # :1232-1233
    if trace is not None:
        trace["E_idx"] = E_idx
        trace["E"] = E
        trace["h_E0"] = h_E
    for l in range(3):                                       # :1234-1235
        h_V, h_E = enc_layer(w, l, h_V, h_E, E_idx, mask, mask_attend)
        if trace is not None:
            trace[f"enc{l}_h_V"] = h_V
            trace[f"enc{l}_h_E"] = h_E
    h_S = w["prot_mpnn.W_s.weight"][S]                       # :1238
    h_ES = torch.cat([h_E, _gather_nodes(h_S, E_idx)], -1)   # :1239
    hidden = []
    for l in range(3):                                       # :1268-1273
        h_ESV = torch.cat([h_ES, _gather_nodes(h_V, E_idx)], -1)
        h_ESV = mask[:, None, None] * h_ESV
        h_V = dec_layer(w, l, h_V, h_ESV, mask)
        hidden.append(h_V)
        if trace is not None:
            trace[f"dec{l}_h_V"] = h_V
    return hidden[::-1], h_S


def unconditional_log_probs(w, X, mask, residue_idx, chain_enc, top_k=TOP_K):
    """ProteinMPNN.unconditional_probs (protein_mpnn_utils.py:1557-1587): the encoder as in ``mpnn_forward``; every
    decoder layer then sees the FIXED input mask_i * [h_E | 0 | h_V_enc[j]] (no sequence, the neighbours' ENCODER states,
    the order mask all zero so mask_bw = 0 and mask_fw = mask_i); log_softmax(W_out h_V) of the last layer, [L,21]."""
    E, E_idx = edge_features(w, X, mask, residue_idx, chain_enc, top_k)
    h_E = _lin(w, "prot_mpnn.W_e", E)
    h_V = torch.zeros((E.shape[0], E.shape[-1]), dtype=E.dtype)
    mask_attend = mask[:, None] * mask[E_idx]
    for l in range(3):
        h_V, h_E = enc_layer(w, l, h_V, h_E, E_idx, mask, mask_attend)
    h_EX = torch.cat([h_E, torch.zeros_like(h_E)], -1)                 # cat_neighbors_nodes(zeros, h_E)   (:1573)
    h_EXV = torch.cat([h_EX, _gather_nodes(h_V, E_idx)], -1)           # ... + h_V_enc[j]                  (:1574)
    h_EXV_fw = mask[:, None, None] * h_EXV                             # mask_fw = mask_1D * (1 - 0)       (:1576-1582)
    for l in range(3):
        h_V = dec_layer(w, l, h_V, h_EXV_fw, mask)                     # (:1583-1584)
    return log_probs(w, h_V)


# --------------------------------------------------------------------------
# transfer head
# --------------------------------------------------------------------------
def _la_center(w):
    if "light_attention.feature_convolution.weight_center" in w:
        return w["light_attention.feature_convolution.weight_center"]
    return w["light_attention.feature_convolution.weight"][:, :, 4]


def head_table(w, hidden, h_S, num_final_layers: int = 2, lightattn: bool = True):
    """All 21 head outputs at every position, [L,21].

    ``TransferModel.forward`` (transfer_model.py:96-116) evaluated once per
    position instead of once per mutation, for any ``cfg.model`` (``num_final_layers``
    decoder states, optional LightAttention, the ``both_out`` MLP found in ``w``;
    transfer_model.py:45-73).  LightAttention on a length-1
    sequence (:107-108, :148-155) reduces to its centre tap: the k=9, pad=4
    convolution only multiplies zeros elsewhere, and the softmax over a length-1
    axis is 1.
    """
    x = torch.cat(list(hidden[:num_final_layers]) + [h_S], -1)   # :82-83, :96-103
    z = x
    if lightattn:
        z = F.linear(x, _la_center(w), w["light_attention.feature_convolution.bias"])
    i = 1
    while f"both_out.{i}.weight" in w:                            # both_out = (ReLU, Linear)*, :67-71
        z = _lin(w, f"both_out.{i}", F.relu(z))
        i += 2
    return z * w["ddg_out.weight"].reshape(()) + w["ddg_out.bias"].reshape(())   # :73, :110-111


def log_probs(w, h_V_last):
    """ProteinMPNN's output layer (protein_mpnn_utils.py:1275-1276): log_softmax(W_out h_V), [L,21]; what
    ``ProteinMPNNBaseline`` reads (analysis/thermompnn_benchmarking.py:52-58)."""
    return F.log_softmax(_lin(w, "prot_mpnn.W_out", h_V_last), dim=-1)


def head_per_mutation(w, hidden, h_S, position: int, wt: int, mut: int):
    """One iteration of the reference's per-mutation loop, as written upstream
    (transfer_model.py:92-116, LightAttention.forward :140-155): two k=9
    Conv1d's over a length-1 sequence, softmax, the MLP, the 1x1 affine and the
    wild-type subtraction.  Needs the full [384,384,9] conv weights.  This is
    what ``bench.py --impl reference`` times.
    """
    x = torch.cat([hidden[0][position], hidden[1][position], h_S[position]], -1)
    x = x[None, :, None]
    o = F.conv1d(x, w["light_attention.feature_convolution.weight"],
                 w["light_attention.feature_convolution.bias"], padding=4)
    a = F.conv1d(x, w["light_attention.attention_convolution.weight"],
                 w["light_attention.attention_convolution.bias"], padding=4)
    z = torch.squeeze(o * torch.softmax(a, dim=-1))
    for i in (1, 3, 5):
        z = _lin(w, f"both_out.{i}", F.relu(z))
    t = F.linear(z[:, None], w["ddg_out.weight"], w["ddg_out.bias"])
    return t[mut][0] - t[wt][0]


# --------------------------------------------------------------------------
# entry points
# --------------------------------------------------------------------------
def cast_weights(w: Dict[str, "torch.Tensor|np.ndarray"], dtype=torch.float32):
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in w.items()}


def ssm_table(w, pdb: dict, dtype=torch.float32, trace: Optional[dict] = None, num_final_layers: int = 2,
              lightattn: bool = True):
    """Site-saturation table for one protein.

    Returns (T [L,21], ddG [L,20], S [L]) with ``ddG[i,a] = T[i,a] - T[i,S_i]``
    (transfer_model.py:113-114), columns in ALPHABET order.  Rows of gap
    residues are still computed (callers emit ``None`` for them, SSM.py:26-27).
    """
    X, S, mask, ridx, cenc = featurize(pdb)
    with torch.no_grad():
        Xt = torch.from_numpy(X).to(dtype)
        mt = torch.from_numpy(mask).to(dtype)
        St = torch.from_numpy(S)
        hidden, h_S = mpnn_forward(w, Xt, St, mt, torch.from_numpy(ridx), torch.from_numpy(cenc), trace=trace)
        T = head_table(w, hidden, h_S, num_final_layers, lightattn)
        ddg = T[:, :20] - torch.gather(T, 1, St[:, None])
        if trace is not None:
            trace["log_probs"] = log_probs(w, hidden[0])
    if trace is not None:
        trace["h_S"] = h_S
        trace["T"] = T
    return T, ddg, S


def forward_as_reference(w, pdb: dict, mutations: Sequence, dtype=torch.float32) -> List[Optional[float]]:
    """``TransferModel.forward`` as written upstream: one MPNN pass, then the
    per-mutation head loop (transfer_model.py:75-121).  ``mutations`` holds
    objects with ``position``, ``wildtype``, ``mutation`` (or ``None``).
    """
    X, S, mask, ridx, cenc = featurize(pdb)
    out: List[Optional[float]] = []
    with torch.no_grad():
        hidden, h_S = mpnn_forward(w, torch.from_numpy(X).to(dtype), torch.from_numpy(S),
                                   torch.from_numpy(mask).to(dtype), torch.from_numpy(ridx),
                                   torch.from_numpy(cenc))
        for m in mutations:
            if m is None:
                out.append(None)
                continue
            out.append(float(head_per_mutation(w, hidden, h_S, m.position,
                                               ALPHABET.index(m.wildtype), ALPHABET.index(m.mutation))))
    return out
