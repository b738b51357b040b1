"""ctypes binding of libqinco_b200.so (C ABI: include/qinco_b200.h).  No torch types cross this boundary: tensors
are passed as raw device pointers (`tensor.data_ptr()`), streams as the CUstream handle.

There is no CPU fallback: if the library is missing and cannot be built, or no sm_100 device is present, creating a
model raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libqinco_b200.so")

QB_OK = 0
STATUS = {0: "QB_OK", -1: "QB_ERR_INVALID", -2: "QB_ERR_CUDA", -3: "QB_ERR_WORKSPACE", -4: "QB_ERR_KERNEL",
          -5: "QB_ERR_NOMEM"}

# every symbol include/qinco_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["qb_version", "qb_last_error", "qb_model_create", "qb_model_destroy", "qb_encode_workspace_bytes",
           "qb_decode_workspace_bytes", "qb_encode", "qb_decode", "qb_encode_host", "qb_decode_host", "qb_check",
           "qb_launch_count", "qb_timing_enable", "qb_timing_read", "qb_model_info", "qb_debug_step", "qb_plan_export", "qb_plan_pack", "qb_plan_pack_pre", "qb_plan_tables",
           "qb_pairwise_create", "qb_pairwise_destroy", "qb_pairwise_decode", "qb_pairwise_check", "qb_pairwise_launch_count",
           "qb_pairwise_last_error", "qb_encode_ivf", "qb_decode_ivf", "qb_encode_ivf_host", "qb_decode_ivf_host", "qb_codes_pack", "qb_codes_unpack"]

_fpp = C.POINTER(C.POINTER(C.c_float))


class QbModelDesc(C.Structure):
    _fields_ = [
        ("D", C.c_int32), ("De", C.c_int32), ("Dh", C.c_int32), ("L", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
        ("A", C.c_int32), ("B", C.c_int32), ("qinco1_mode", C.c_int32), ("device", C.c_int32),
        ("codebook", _fpp), ("substep_codebook", _fpp), ("concat_w", _fpp), ("concat_b", _fpp), ("up_w", _fpp),
        ("down_w", _fpp), ("in_proj", _fpp), ("out_proj", _fpp),
        ("data_mean", C.POINTER(C.c_float)), ("data_std", C.c_float),
        ("opt_hc", C.c_int32), ("opt_n_tiles", C.c_int32), ("opt_slot_bytes", C.c_int32), ("opt_max_stage", C.c_int32),
        ("opt_max_slab_k", C.c_int32), ("opt_stagger", C.c_int32),
        ("ivf_K", C.c_int32), ("ivf_centroids", C.POINTER(C.c_float)),
    ]


class QbPairwiseDesc(C.Structure):
    """Mirror of qb_pairwise_desc (include/qinco_b200.h)."""
    _fields_ = [("D", C.c_int32), ("M", C.c_int32), ("K", C.c_int32), ("Mt", C.c_int32), ("ivf_K", C.c_int32),
                ("device", C.c_int32), ("codebook", C.POINTER(C.c_float)), ("combine", C.POINTER(C.c_int64)),
                ("ivf_code_map", C.POINTER(C.c_int64))]


class QbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the library (building it first if the sources are newer); raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("QINCO_B200_LIB")      # A/B tests of compile-time kernel variants: load exactly this build
    if override:
        if not os.path.exists(override):
            raise RuntimeError(f"QINCO_B200_LIB={override} does not exist")
        build_if_missing = False
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine: use the prebuilt library if there is one
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"libqinco_b200.so is missing and could not be built: {e}") from e
    if not override and not os.path.exists(LIB_PATH):
        raise RuntimeError("libqinco_b200.so is missing; run `python -m qinco_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(override or LIB_PATH)
    vp, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
    lib.qb_version.restype = C.c_int
    lib.qb_last_error.restype = C.c_char_p
    lib.qb_model_create.argtypes = [C.POINTER(QbModelDesc), C.POINTER(vp)]
    lib.qb_model_destroy.argtypes = [vp]
    lib.qb_encode_workspace_bytes.argtypes = [vp, i64]
    lib.qb_encode_workspace_bytes.restype = sz
    lib.qb_decode_workspace_bytes.argtypes = [vp, i64]
    lib.qb_decode_workspace_bytes.restype = sz
    lib.qb_encode.argtypes = [vp, vp, i64, C.c_int, vp, vp, vp, sz, vp]
    lib.qb_decode.argtypes = [vp, vp, i64, C.c_int, vp, vp, sz, vp]
    lib.qb_encode_ivf.argtypes = [vp, vp, i64, C.c_int, vp, vp, vp, vp, sz, vp]
    lib.qb_decode_ivf.argtypes = [vp, vp, vp, i64, C.c_int, vp, vp, sz, vp]
    lib.qb_encode_ivf_host.argtypes = [vp, vp, i64, C.c_int, vp, vp, vp]
    lib.qb_decode_ivf_host.argtypes = [vp, vp, vp, i64, C.c_int, vp]
    lib.qb_encode_host.argtypes = [vp, vp, i64, C.c_int, vp, vp]
    lib.qb_decode_host.argtypes = [vp, vp, i64, C.c_int, vp]
    lib.qb_codes_pack.argtypes = [vp, vp, C.c_int, i64, i64, i64, vp, vp, vp]
    lib.qb_codes_unpack.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.qb_check.argtypes = [vp]
    lib.qb_launch_count.argtypes = [vp]
    lib.qb_launch_count.restype = i64
    lib.qb_timing_enable.argtypes = [vp, C.c_int]
    lib.qb_timing_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
    lib.qb_model_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.c_int]
    lib.qb_debug_step.argtypes = [vp, C.c_int, vp, vp, i64, vp, vp, sz, vp]
    lib.qb_pairwise_create.argtypes = [C.POINTER(QbPairwiseDesc), C.POINTER(vp)]
    lib.qb_pairwise_destroy.argtypes = [vp]
    lib.qb_pairwise_decode.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.qb_pairwise_check.argtypes = [vp]
    lib.qb_pairwise_launch_count.argtypes = [vp]
    lib.qb_pairwise_launch_count.restype = i64
    lib.qb_pairwise_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != QB_OK:
        raise QbError(rc, load().qb_last_error().decode(errors="replace"))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class _PtrArray:
    """A C array of float* built from numpy arrays (kept alive here)."""

    def __init__(self, arrays):
        self.keep = [None if a is None else _f32(a) for a in arrays]
        self.arr = (C.POINTER(C.c_float) * len(self.keep))()
        for i, a in enumerate(self.keep):
            self.arr[i] = a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None

    @property
    def ptr(self):
        return C.cast(self.arr, _fpp)


INFO_FIELDS = ["D", "De", "Dh", "L", "K", "has_proj", "skip", "n_ops_block", "n_ops_out", "hc", "n_hchunk", "n_tiles",
               "oc", "n_ochunk", "slot_bytes", "n_stage", "smem_total", "block_w_bytes", "w_blob_bytes", "n_sm",
               "default_chunk", "pair", "decode_loop", "loop_n_stage", "loop_smem_total"]


class Handle:
    """Owns one qb_model.  `weights` is a dict keyed like the reference state dict (numpy fp32 arrays)."""

    def __init__(self, cfg: dict, weights: dict, device: int = 0, plan_opts: dict | None = None):
        lib = load()
        D, De, Dh, L, M, K, A, B = (int(cfg[k]) for k in ("D", "de", "dh", "L", "M", "K", "A", "B"))
        g = weights.get
        keep = []

        def arr(fmt, rng):
            pa = _PtrArray([g(fmt.format(m=m)) if m >= 1 else None for m in rng])
            keep.append(pa)
            return pa.ptr

        ivf_K = int(cfg.get("ivf_K") or 0)
        S = M + 1 if ivf_K else M          # quantisation steps (cfg._M_ivf); per-step arrays are indexed by step
        cb = _PtrArray([None if (ivf_K and m == 0) else weights[f"steps.{m}.codebook.weight"] for m in range(S)])
        keep.append(cb)
        d = QbModelDesc()
        d.D, d.De, d.Dh, d.L, d.M, d.K, d.A, d.B = D, De, Dh, L, M, K, A, B
        d.qinco1_mode = int(bool(cfg["qinco1_mode"]))
        d.device = int(device)
        d.codebook = cb.ptr
        steps = range(S)
        if ivf_K:
            cent = _f32(weights["steps.0.ivf_centroids.weight"])
            assert cent.shape == (ivf_K, D), f"steps.0.ivf_centroids.weight must be [{ivf_K}, {D}]"
            keep.append(cent)
            d.ivf_K = ivf_K
            d.ivf_centroids = cent.ctypes.data_as(C.POINTER(C.c_float))
        d.substep_codebook = arr("steps.{m}.substep.codebook.weight", steps) if A > 0 else None
        d.concat_w = arr("steps.{m}.concat.mlp.weight", steps)
        d.concat_b = arr("steps.{m}.concat.mlp.bias", steps)
        ups = _PtrArray([g(f"steps.{m}.residual_blocks.{l}.up_proj.weight") if m >= 1 else None
                         for m in range(S) for l in range(L)])
        downs = _PtrArray([g(f"steps.{m}.residual_blocks.{l}.down_proj.weight") if m >= 1 else None
                           for m in range(S) for l in range(L)])
        keep += [ups, downs]
        d.up_w, d.down_w = ups.ptr, downs.ptr
        if De != D:
            d.in_proj = arr("steps.{m}.in_proj.weight", steps)
            d.out_proj = arr("steps.{m}.out_proj.weight", steps)
        mean = _f32(weights["data_mean"]) if "data_mean" in weights else np.zeros(D, np.float32)
        keep.append(mean)
        d.data_mean = mean.ctypes.data_as(C.POINTER(C.c_float))
        d.data_std = float(np.asarray(weights.get("data_std", 1.0)).reshape(-1)[0])
        for k, v in (plan_opts or {}).items():
            setattr(d, "opt_" + k, int(v))
        h = C.c_void_p()
        check(lib.qb_model_create(C.byref(d), C.byref(h)))
        del keep
        self._h = h
        self._lib = lib
        self.M, self.D, self.De, self.K = M, D, De, K

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qb_model_destroy(self._h)
            self._h = None

    __del__ = close

    # --- thin wrappers, raw pointers only ---
    def encode_workspace_bytes(self, n: int) -> int:
        return int(self._lib.qb_encode_workspace_bytes(self._h, n))

    def decode_workspace_bytes(self, n: int) -> int:
        return int(self._lib.qb_decode_workspace_bytes(self._h, n))

    def encode(self, x_ptr, n, normalize, codes_ptr, xhat_ptr, ws_ptr, ws_bytes, stream):
        check(self._lib.qb_encode(self._h, x_ptr, n, int(normalize), codes_ptr, xhat_ptr, ws_ptr, ws_bytes, stream))

    def decode(self, codes_ptr, n, denormalize, out_ptr, ws_ptr, ws_bytes, stream):
        check(self._lib.qb_decode(self._h, codes_ptr, n, int(denormalize), out_ptr, ws_ptr, ws_bytes, stream))

    def encode_ivf(self, x_ptr, n, normalize, ivf_ptr, codes_ptr, xhat_ptr, ws_ptr, ws_bytes, stream):
        check(self._lib.qb_encode_ivf(self._h, x_ptr, n, int(normalize), ivf_ptr, codes_ptr, xhat_ptr, ws_ptr, ws_bytes, stream))

    def decode_ivf(self, ivf_ptr, codes_ptr, n, denormalize, out_ptr, ws_ptr, ws_bytes, stream):
        check(self._lib.qb_decode_ivf(self._h, ivf_ptr, codes_ptr, n, int(denormalize), out_ptr, ws_ptr, ws_bytes, stream))

    def encode_ivf_host(self, x: np.ndarray, normalize: bool, want_xhat: bool = False):
        """IVF models, host buffers: -> (ivf codes int32 [n], codes uint8 [n, M], xhat or None)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        ivf = np.empty((n,), np.int32)
        codes = np.empty((n, self.M), np.uint8)
        xhat = np.empty((n, self.D), np.float32) if want_xhat else None
        check(self._lib.qb_encode_ivf_host(self._h, x.ctypes.data, n, int(normalize), ivf.ctypes.data, codes.ctypes.data,
                                           xhat.ctypes.data if want_xhat else None))
        return ivf, codes, xhat

    def decode_ivf_host(self, ivf: np.ndarray, codes: np.ndarray, denormalize: bool) -> np.ndarray:
        ivf = np.ascontiguousarray(ivf, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        out = np.empty((codes.shape[0], self.D), np.float32)
        check(self._lib.qb_decode_ivf_host(self._h, ivf.ctypes.data, codes.ctypes.data, codes.shape[0], int(denormalize),
                                           out.ctypes.data))
        return out

    def encode_host(self, x: np.ndarray, normalize: bool, want_xhat: bool = False):
        x = _f32(x)
        n = len(x)
        codes = np.empty((n, self.M), np.uint8)
        xhat = np.empty((n, self.D), np.float32) if want_xhat else None
        check(self._lib.qb_encode_host(self._h, x.ctypes.data, n, int(normalize), codes.ctypes.data,
                                       xhat.ctypes.data if want_xhat else None))
        return codes, xhat

    def decode_host(self, codes: np.ndarray, denormalize: bool) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        n = len(codes)
        out = np.empty((n, self.D), np.float32)
        check(self._lib.qb_decode_host(self._h, codes.ctypes.data, n, int(denormalize), out.ctypes.data))
        return out

    def codes_pack(self, src_ptr, elem_bytes, stride_row, stride_col, n, codes_ptr, ivf_ptr, stream):
        check(self._lib.qb_codes_pack(self._h, src_ptr, elem_bytes, stride_row, stride_col, n, codes_ptr, ivf_ptr, stream))

    def codes_unpack(self, codes_ptr, ivf_ptr, n, dst_ptr, stream):
        check(self._lib.qb_codes_unpack(self._h, codes_ptr, ivf_ptr, n, dst_ptr, stream))

    def debug_step(self, step, xhat_ptr, codes_ptr, n, out_ptr, ws_ptr, ws_bytes, stream):
        check(self._lib.qb_debug_step(self._h, step, xhat_ptr, codes_ptr, n, out_ptr, ws_ptr, ws_bytes, stream))

    def check(self):
        check(self._lib.qb_check(self._h))

    KINDS = ["prep", "mlp_score", "select", "mlp_apply", "other", "ivf"]

    def timing_enable(self, on: bool = True):
        check(self._lib.qb_timing_enable(self._h, int(on)))

    def timing_read(self) -> dict:
        """{kind: (ms, launches, rows)} since the last read (synchronises on the recorded events)."""
        n = len(self.KINDS)
        ms, cnt, rows = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
        rc = self._lib.qb_timing_read(self._h, ms, cnt, rows, n)
        if rc < 0:
            check(rc)
        return {k: (ms[i], cnt[i], rows[i]) for i, k in enumerate(self.KINDS)}

    @property
    def launch_count(self) -> int:
        return int(self._lib.qb_launch_count(self._h))

    def info(self, step: int = 1) -> dict:
        buf = (C.c_int32 * 32)()
        n = self._lib.qb_model_info(self._h, step, buf, 32)
        if n < 0:
            check(n)
        return dict(zip(INFO_FIELDS, list(buf)[:n]))
