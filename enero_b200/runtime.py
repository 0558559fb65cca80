"""Device context and the stateless batched operators (numpy in / numpy out).

One Context per (process, device).  All compute goes through libenero_b200.so;
there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .topology import Topology

WEIGHT_SHAPES = ((40, 20), (20,), (20, 60), (20, 60), (2, 60), (20, 20), (20,), (20, 20), (20,), (20, 1), (1,))
WEIGHT_SIZES = tuple(int(np.prod(s)) for s in WEIGHT_SHAPES)


def flatten_weights(tensors) -> np.ndarray:
    """11 tensors in Keras variable order -> flat fp32[4201]."""
    if isinstance(tensors, np.ndarray) and tensors.ndim == 1:
        flat = np.ascontiguousarray(tensors, dtype=np.float32)
    else:
        if len(tensors) != 11:
            raise ValueError("expected the 11 model tensors in Keras variable order")
        parts = []
        for t, shp in zip(tensors, WEIGHT_SHAPES):
            a = np.asarray(t, dtype=np.float32)
            if a.shape != shp:
                raise ValueError("weight shape %s, expected %s" % (a.shape, shp))
            parts.append(a.reshape(-1))
        flat = np.ascontiguousarray(np.concatenate(parts))
    if flat.size != _lib.NPARAMS:
        raise ValueError("expected %d parameters" % _lib.NPARAMS)
    return flat


def unflatten_weights(flat: np.ndarray) -> list:
    out, o = [], 0
    for shp, n in zip(WEIGHT_SHAPES, WEIGHT_SIZES):
        out.append(np.array(flat[o:o + n], dtype=np.float32).reshape(shp))
        o += n
    return out


class Context:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        if self._lib.enero_device_count() <= 0:
            raise _lib.EneroError("enero_b200 needs a CUDA device: none visible (there is no CPU fallback)")
        h = C.c_void_p()
        _lib.check(self._lib.enero_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._children = weakref.WeakSet()       # EpisodeBatch objects created on this context

    def _adopt(self, child):
        self._children.add(child)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            for child in list(getattr(self, "_children", ())):      # batches hold pointers into the context
                child.close()
            self._lib.enero_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._lib.enero_ctx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.enero_ctx_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.enero_ctx_stream(self._h) or 0)

    @property
    def math(self) -> str:
        """'fast' (MUFU ex2/rcp forms: the default) or 'accurate' (libdevice expf/tanhf, IEEE division)"""
        return "accurate" if self._lib.enero_ctx_get_math(self._h) == _lib.MATH_ACCURATE else "fast"

    @math.setter
    def math(self, mode: str):
        if mode not in ("fast", "accurate"):
            raise ValueError("math mode must be 'fast' or 'accurate'")
        _lib.check(self._lib.enero_ctx_set_math(self._h, _lib.MATH_ACCURATE if mode == "accurate" else _lib.MATH_FAST))

    @property
    def engine(self) -> str:
        """'ffma' (k_mp_iter, FP32 pipe) or 'tc' (k_mp_tc, tcgen05 tensor cores with 3xTF32 products)"""
        return "tc" if self._lib.enero_ctx_get_engine(self._h) == 1 else "ffma"

    @engine.setter
    def engine(self, name: str):
        if name not in ("ffma", "tc"):
            raise ValueError("engine must be 'ffma' or 'tc'")
        _lib.check(self._lib.enero_ctx_set_engine(self._h, 1 if name == "tc" else 0))

    def fp32_peak_tflops(self) -> float:
        """FP32 FMA peak of this device, measured live (FFMA2 register loop)"""
        v = C.c_double(0)
        _lib.check(self._lib.enero_ctx_fp32_peak(self._h, C.byref(v)))
        return v.value

    def set_stream(self, stream=None):
        """launch on the caller's CUDA stream (raw cudaStream_t as int, or a torch.cuda.Stream); None = the context's own"""
        raw = getattr(stream, "cuda_stream", stream)
        _lib.check(self._lib.enero_ctx_set_stream(self._h, C.c_void_p(int(raw)) if raw else None))

    def profile_begin(self):
        _lib.check(self._lib.enero_ctx_profile_begin(self._h))

    def profile_end(self):
        ms = C.c_double(0)
        n = C.c_int64(0)
        _lib.check(self._lib.enero_ctx_profile_end(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- weights -------------------------------------------------------------------------
    def upload_weights(self, which: int, tensors):
        """raw upload into the context's actor / critic buffer; a model object that owned the buffer (models.bind)
        loses it, so its next use re-uploads its own weights instead of computing with these"""
        prev = self.__dict__.setdefault("_weight_owner", {}).get(which)
        if prev is not None and prev._device_dirty:
            prev.sync_from_device()                  # keep what a GPU optimiser step left there
        self._weight_owner.pop(which, None)
        flat = flatten_weights(tensors)
        _lib.check(self._lib.enero_weights_upload(self._h, which, _lib.ptr(flat)))

    def download_weights(self, which: int) -> list:
        flat = np.zeros(_lib.NPARAMS, dtype=np.float32)
        _lib.check(self._lib.enero_weights_download(self._h, which, _lib.ptr(flat)))
        return unflatten_weights(flat)

    # ---- signature-compatible GNN ------------------------------------------------------------
    def gnn_forward(self, which, link_state, graph_id, first, second, n_graphs, T=5) -> np.ndarray:
        ls = np.ascontiguousarray(link_state, dtype=np.float32)
        if ls.ndim != 2 or ls.shape[1] != 20:
            raise ValueError("link_state must be [n, 20]")
        f = np.ascontiguousarray(first, dtype=np.int32).reshape(-1)
        s = np.ascontiguousarray(second, dtype=np.int32).reshape(-1)
        if f.size != s.size:
            raise ValueError("first and second must have the same length")
        gid = None if graph_id is None else np.ascontiguousarray(graph_id, dtype=np.int32).reshape(-1)
        if gid is not None and gid.size != ls.shape[0]:
            raise ValueError("graph_id must have one entry per link_state row")
        out = np.zeros(int(n_graphs), dtype=np.float32)
        _lib.check(self._lib.enero_gnn_forward(self._h, which, _lib.ptr(ls), _lib.ptr(gid), _lib.ptr(f), _lib.ptr(s),
                                               ls.shape[0], f.size, int(n_graphs), int(T), _lib.ptr(out), _lib.HOST))
        return out

    # ---- fused decision batch ---------------------------------------------------------------
    def decide_batch(self, topo: Topology, loads, sd, bw):
        """-> (logits[B,Kmax], probs[B,Kmax], nK[B], action[B]) for B independent decisions."""
        loads = np.ascontiguousarray(loads, dtype=np.float64)
        sd = np.ascontiguousarray(sd, dtype=np.int32)
        bw = np.ascontiguousarray(bw, dtype=np.float64).reshape(-1)
        B = bw.size
        if loads.shape != (B, topo.E) or sd.shape != (B, 2):
            raise ValueError("loads must be [B,E] and sd [B,2]")
        _lib.check(self._lib.enero_topo_upload(self._h, topo.handle))
        logits = np.zeros((B, topo.Kmax), dtype=np.float32)
        probs = np.zeros((B, topo.Kmax), dtype=np.float32)
        nK = np.zeros(B, dtype=np.int32)
        action = np.zeros(B, dtype=np.int32)
        _lib.check(self._lib.enero_decide_batch(self._h, topo.handle, B, _lib.ptr(loads), _lib.ptr(sd), _lib.ptr(bw),
                                                _lib.ptr(logits), _lib.ptr(probs), _lib.ptr(nK), _lib.ptr(action), _lib.HOST))
        return logits, probs, nK, action

    def decide_value_batch(self, topo: Topology, loads, sd, bw):
        """-> (probs[B,Kmax], nK[B], values[B]): actor and critic on the same states, one stream round trip."""
        loads = np.ascontiguousarray(loads, dtype=np.float64)
        sd = np.ascontiguousarray(sd, dtype=np.int32)
        bw = np.ascontiguousarray(bw, dtype=np.float64).reshape(-1)
        B = bw.size
        if loads.shape != (B, topo.E) or sd.shape != (B, 2):
            raise ValueError("loads must be [B,E] and sd [B,2]")
        _lib.check(self._lib.enero_topo_upload(self._h, topo.handle))
        probs = np.zeros((B, topo.Kmax), dtype=np.float32)
        nK = np.zeros(B, dtype=np.int32)
        values = np.zeros(B, dtype=np.float32)
        _lib.check(self._lib.enero_decide_value_batch(self._h, topo.handle, B, _lib.ptr(loads), _lib.ptr(sd), _lib.ptr(bw),
                                                      _lib.ptr(probs), _lib.ptr(nK), _lib.ptr(values)))
        return probs, nK, values

    def value_batch(self, topo: Topology, loads) -> np.ndarray:
        loads = np.ascontiguousarray(loads, dtype=np.float64)
        B = loads.shape[0]
        _lib.check(self._lib.enero_topo_upload(self._h, topo.handle))
        out = np.zeros(B, dtype=np.float32)
        _lib.check(self._lib.enero_value_batch(self._h, topo.handle, B, _lib.ptr(loads), _lib.ptr(out), _lib.HOST))
        return out


_default = {}


def default_context(device: int = 0) -> Context:
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
