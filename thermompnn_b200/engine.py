"""ctypes binding of the C ABI in ``include/thermompnn_b200.h``.

This is the only door into the CUDA path.  There is deliberately no fallback:
if the shared library is missing or no B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TMPNN_LIB=dev selects the development build (-DTMPNN_DEV: the TMPNN_PDL / TMPNN_TC_TRACE measurement switches)
LIB_PATH = os.path.join(_HERE, "libthermompnn_b200_dev.so" if os.environ.get("TMPNN_LIB") == "dev" else "libthermompnn_b200.so")
MAX_HIDDEN = 8

IN_DEVICE, OUT_DEVICE, NO_SYNC, UNCONDITIONAL = 1, 2, 4, 8

# every symbol include/thermompnn_b200.h declares
SYMBOLS = [
    "tmpnn_abi_version", "tmpnn_create", "tmpnn_destroy", "tmpnn_last_error", "tmpnn_weight_count",
    "tmpnn_weight_info", "tmpnn_weight_blob_floats", "tmpnn_load_weights", "tmpnn_run_batch", "tmpnn_fetch",
    "tmpnn_set_debug_stop", "tmpnn_set_profiling", "tmpnn_profile_count", "tmpnn_profile_get",
    "tmpnn_profile_reset", "tmpnn_launch_count", "tmpnn_graph_launch_count", "tmpnn_default_cfg", "tmpnn_cfg_weight_count",
    "tmpnn_cfg_weight_info", "tmpnn_cfg_weight_blob_floats", "tmpnn_load_weights_cfg", "tmpnn_run_batch_ex", "tmpnn_set_graphs",
]


class ModelCfg(C.Structure):
    """``tmpnn_model_cfg``: which head a checkpoint carries (cfg.model.* of transfer_model.py:45-73)."""
    _fields_ = [("num_final_layers", C.c_int32), ("lightattn", C.c_int32), ("n_hidden", C.c_int32),
                ("hidden_dims", C.c_int32 * MAX_HIDDEN), ("has_head", C.c_int32)]

    @classmethod
    def make(cls, num_final_layers=2, lightattn=True, hidden_dims=(64, 32), has_head=True) -> "ModelCfg":
        hidden_dims = list(hidden_dims)
        if len(hidden_dims) > MAX_HIDDEN:
            raise ValueError(f"at most {MAX_HIDDEN} hidden layers")
        c = cls()
        c.num_final_layers, c.lightattn, c.n_hidden, c.has_head = int(num_final_layers), int(bool(lightattn)), len(hidden_dims), int(bool(has_head))
        for i, d in enumerate(hidden_dims):
            c.hidden_dims[i] = int(d)
        return c

    def key(self):
        return (self.num_final_layers, self.lightattn, tuple(self.hidden_dims[:self.n_hidden]), self.has_head)

_lib = None


class EngineError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen the in-tree library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). thermompnn_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    lib.tmpnn_abi_version.restype = C.c_int
    lib.tmpnn_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.tmpnn_destroy.argtypes = [vp]
    lib.tmpnn_destroy.restype = None
    lib.tmpnn_last_error.argtypes = [vp]
    lib.tmpnn_last_error.restype = C.c_char_p
    lib.tmpnn_weight_count.restype = C.c_int
    lib.tmpnn_weight_info.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.tmpnn_weight_blob_floats.restype = i64
    lib.tmpnn_load_weights.argtypes = [vp, vp, i64]
    lib.tmpnn_run_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, u32, vp]
    lib.tmpnn_fetch.argtypes = [vp, C.c_char_p, vp, i64]
    lib.tmpnn_set_debug_stop.argtypes = [vp, C.c_int]
    lib.tmpnn_set_profiling.argtypes = [vp, C.c_int]
    lib.tmpnn_profile_count.argtypes = [vp]
    lib.tmpnn_profile_get.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(i64)]
    lib.tmpnn_profile_reset.argtypes = [vp]
    lib.tmpnn_launch_count.argtypes = [vp]
    lib.tmpnn_launch_count.restype = i64
    lib.tmpnn_graph_launch_count.argtypes = [vp]
    lib.tmpnn_graph_launch_count.restype = i64
    cfgp = C.POINTER(ModelCfg)
    lib.tmpnn_default_cfg.argtypes = [cfgp]
    lib.tmpnn_default_cfg.restype = None
    lib.tmpnn_cfg_weight_count.argtypes = [cfgp]
    lib.tmpnn_cfg_weight_info.argtypes = [cfgp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.tmpnn_cfg_weight_blob_floats.argtypes = [cfgp]
    lib.tmpnn_cfg_weight_blob_floats.restype = i64
    lib.tmpnn_load_weights_cfg.argtypes = [vp, cfgp, vp, i64]
    lib.tmpnn_run_batch_ex.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp]
    lib.tmpnn_set_graphs.argtypes = [vp, C.c_int]
    _lib = lib
    return lib


def weight_table(cfg: Optional[ModelCfg] = None) -> List[Tuple[str, int, int, int]]:
    """[(name, offset, rows, cols)] of the raw checkpoint contract for a model configuration (no GPU needed)."""
    lib = load_library()
    cfg = cfg if cfg is not None else ModelCfg.make()
    n = lib.tmpnn_cfg_weight_count(C.byref(cfg))
    if n < 0:
        raise ValueError("invalid model configuration")
    out = []
    for i in range(n):
        name, off, r, c = C.c_char_p(), C.c_int64(), C.c_int32(), C.c_int32()
        lib.tmpnn_cfg_weight_info(C.byref(cfg), i, C.byref(name), C.byref(off), C.byref(r), C.byref(c))
        out.append((name.value.decode(), off.value, r.value, c.value))
    return out


def pack_weight_blob(state: Dict[str, np.ndarray], cfg: Optional[ModelCfg] = None) -> np.ndarray:
    """Pack a reference state_dict (``model.`` prefix stripped) into the blob ``tmpnn_load_weights[_cfg]`` takes."""
    lib = load_library()
    cfg = cfg if cfg is not None else ModelCfg.make()
    blob = np.zeros(lib.tmpnn_cfg_weight_blob_floats(C.byref(cfg)), dtype=np.float32)
    for name, off, r, c in weight_table(cfg):
        if name not in state:
            raise KeyError(f"checkpoint is missing tensor '{name}'")
        v = np.asarray(state[name], dtype=np.float32)
        if v.size != r * c:
            raise ValueError(f"tensor '{name}' has {v.size} elements, expected {r}x{c}")
        blob[off:off + r * c] = v.reshape(-1)
    return blob


def _ptr(a) -> int:
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class Engine:
    """One GPU's ThermoMPNN engine (wraps a ``tmpnn_ctx``)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.tmpnn_create(device, C.byref(self._ctx))
        if rc != 0:
            raise EngineError(self._lib.tmpnn_last_error(None).decode())
        self.device = device
        self._n_last = 0

    def close(self):
        ctx = getattr(self, "_ctx", None)
        if ctx is not None and ctx.value:
            self._ctx = None
            self._lib.tmpnn_destroy(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(f"[{rc}] " + self._lib.tmpnn_last_error(self._ctx).decode())

    def load_weights(self, blob: np.ndarray, cfg: Optional[ModelCfg] = None):
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        cfg = cfg if cfg is not None else ModelCfg.make()
        self._check(self._lib.tmpnn_load_weights_cfg(self._ctx, C.byref(cfg), blob.ctypes.data, blob.size))
        self.cfg = cfg

    @staticmethod
    def _host_out(out, N, what):
        """A caller-supplied host result array goes to C as a raw pointer: check it before that."""
        if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous and out.size == N * 21):
            raise ValueError(f"{what} must be a C-contiguous float32 array of {N} x 21 elements")
        return out

    def run(self, offsets: np.ndarray, X, S, residue_idx, chain_enc, mask, out=None, *, flags: int = 0,
            stream: int = 0, log_probs=None, want_table: bool = True):
        """Run one ragged batch.  With flags == 0 every array is a host numpy array (``offsets`` always is);
        with IN_DEVICE / OUT_DEVICE pass raw device pointers (ints) for the inputs / outputs.

        ``out``: the [N,21] head table (allocated when None, skipped with ``want_table=False``).
        ``log_probs``: True (allocate) / array / device pointer to also get ProteinMPNN's log-probabilities [N,21].
        Returns the table, or (table, log_probs) when log_probs were asked for."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if offsets.ndim != 1 or offsets.size < 2:
            raise ValueError("offsets must be a 1-D array of B+1 entries")
        B, N = offsets.size - 1, int(offsets[-1])
        if not (flags & IN_DEVICE):
            X = np.ascontiguousarray(X, dtype=np.float32)
            S = np.ascontiguousarray(S, dtype=np.int32)
            residue_idx = np.ascontiguousarray(residue_idx, dtype=np.int32)
            chain_enc = np.ascontiguousarray(chain_enc, dtype=np.int32)
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if not (X.size == N * 12 and S.size == N and residue_idx.size == N and chain_enc.size == N and mask.size == N):
                raise ValueError("input array sizes do not match offsets[-1]")
        elif any(isinstance(a, np.ndarray) for a in (X, S, residue_idx, chain_enc, mask)):
            raise ValueError("IN_DEVICE takes raw device pointers (ints), not host arrays")
        dev_out = bool(flags & OUT_DEVICE)
        copy_out = copy_lp = False
        if not want_table:
            out = None
        elif out is None:
            if dev_out:
                raise ValueError("OUT_DEVICE needs an explicit device pointer")
            # results land in a per-engine buffer with a stable address (so that repeated call shapes can replay the
            # captured graph, whose D2H copy targets it) and are handed out as a fresh array
            out, copy_out = self._result_buffer("table", N), True
        elif not dev_out:
            self._host_out(out, N, "out")
        elif isinstance(out, np.ndarray):
            raise ValueError("OUT_DEVICE takes a raw device pointer for `out`")
        lp = None
        if log_probs is not None and log_probs is not False:
            if log_probs is True:
                if dev_out:
                    raise ValueError("OUT_DEVICE needs an explicit device pointer for log_probs")
                lp, copy_lp = self._result_buffer("log_probs", N), True
            elif not dev_out:
                lp = self._host_out(log_probs, N, "log_probs")
            else:
                lp = log_probs
        if out is None and lp is None:
            raise ValueError("nothing to compute: want_table=False and no log_probs")
        self._check(self._lib.tmpnn_run_batch_ex(self._ctx, B, offsets.ctypes.data, _ptr(X), _ptr(S), _ptr(residue_idx),
                                                 _ptr(chain_enc), _ptr(mask), None if out is None else _ptr(out),
                                                 None if lp is None else _ptr(lp), flags, stream or None))
        self._n_last = N
        if copy_out:
            out = out.copy()
        if copy_lp:
            lp = lp.copy()
        return out if lp is None else (out, lp)

    def _result_buffer(self, what: str, N: int) -> np.ndarray:
        cache = self.__dict__.setdefault("_result_cache", {})
        buf = cache.get(what)
        if buf is None or buf.shape[0] < N:
            buf = cache[what] = np.empty((max(N, 64), 21), dtype=np.float32)
        return buf[:N]

    def set_graphs(self, on: bool):
        self._check(self._lib.tmpnn_set_graphs(self._ctx, int(on)))

    def graph_launch_count(self) -> int:
        return int(self._lib.tmpnn_graph_launch_count(self._ctx))

    # ---- introspection ----------------------------------------------------
    _SHAPES = {"E_idx": ((48,), np.int32), "D_nb": ((48,), np.float32), "h_E": ((48, 128), np.float32),
               "h_V": ((128,), np.float32), "dec0": ((128,), np.float32), "dec1": ((128,), np.float32),
               "dec2": ((128,), np.float32), "table": ((21,), np.float32), "log_probs": ((21,), np.float32)}

    def fetch(self, name: str) -> np.ndarray:
        shape, dt = self._SHAPES[name]
        a = np.empty((self._n_last,) + shape, dtype=dt)
        self._check(self._lib.tmpnn_fetch(self._ctx, name.encode(), a.ctypes.data, a.nbytes))
        return a

    def fetch_trace(self) -> np.ndarray:
        """TMPNN_TC_TRACE=1 only: [2 kernels][16 tiles][32 slots] clock64 stamps of block 0."""
        a = np.zeros((2, 16, 32), dtype=np.int64)
        self._check(self._lib.tmpnn_fetch(self._ctx, b"tc_trace", a.ctypes.data, a.nbytes))
        return a

    def fetch_node_trace(self) -> np.ndarray:
        """TMPNN_TC_TRACE=1 only: clock64 stamps of the node_update cluster kernel (encoder layer 1, cluster 0)."""
        a = np.zeros(64, dtype=np.int64)
        self._check(self._lib.tmpnn_fetch(self._ctx, b"node_trace", a.ctypes.data, a.nbytes))
        return a

    def set_debug_stop(self, stage: int):
        self._check(self._lib.tmpnn_set_debug_stop(self._ctx, stage))

    def set_profiling(self, on: bool):
        self._check(self._lib.tmpnn_set_profiling(self._ctx, int(on)))

    def profile_reset(self):
        self._check(self._lib.tmpnn_profile_reset(self._ctx))

    def profile(self) -> Dict[str, Tuple[float, int]]:
        out = {}
        for i in range(self._lib.tmpnn_profile_count(self._ctx)):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
            self._lib.tmpnn_profile_get(self._ctx, i, C.byref(name), C.byref(ms), C.byref(n))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    def launch_count(self) -> int:
        return int(self._lib.tmpnn_launch_count(self._ctx))
