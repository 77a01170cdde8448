"""ctypes binding of oracle/sfod_oracle.c (TEST INFRASTRUCTURE ONLY; never imported by the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsfod_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sfod_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsfod_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def nms(boxes, scores, thr: float) -> np.ndarray:
    b, bp = _f(boxes); s, sp = _f(scores)
    n = b.shape[0]
    keep = np.zeros(max(n, 1), dtype=np.int64); nk = C.c_int64(0)
    rc = lib().orc_nms(bp, sp, C.c_int64(n), C.c_double(thr), keep.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nk))
    assert rc == 0
    return keep[: nk.value].copy()


def argsort_desc_stable(scores) -> np.ndarray:
    s, sp = _f(scores)
    o = np.zeros(max(s.shape[0], 1), dtype=np.int64)
    assert lib().orc_argsort_desc_stable(sp, C.c_int64(s.shape[0]), o.ctypes.data_as(C.POINTER(C.c_int64))) == 0
    return o[: s.shape[0]]


def roi_align_fwd(x, rois, out_hw, scale, sampling_ratio, aligned) -> np.ndarray:
    x, xp = _f(x); r, rp = _f(rois)
    N, Cc, H, W = x.shape; R = r.shape[0]; PH, PW = out_hw
    out = np.zeros((R, Cc, PH, PW), dtype=np.float32)
    assert lib().orc_roi_align_fwd(xp, rp, N, Cc, H, W, R, PH, PW, C.c_float(scale), int(sampling_ratio), int(bool(aligned)),
                                   out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return out


def roi_align_bwd(grad_out, rois, in_shape, scale, sampling_ratio, aligned) -> np.ndarray:
    g, gp = _f(grad_out); r, rp = _f(rois)
    N, Cc, H, W = in_shape; R, _, PH, PW = g.shape
    gi = np.zeros(in_shape, dtype=np.float32)
    assert lib().orc_roi_align_bwd(gp, rp, N, Cc, H, W, R, PH, PW, C.c_float(scale), int(sampling_ratio), int(bool(aligned)),
                                   gi.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return gi


def roi_pool_fwd(x, rois, out_hw, scale):
    x, xp = _f(x); r, rp = _f(rois)
    N, Cc, H, W = x.shape; R = r.shape[0]; PH, PW = out_hw
    out = np.zeros((R, Cc, PH, PW), dtype=np.float32); am = np.zeros((R, Cc, PH, PW), dtype=np.int32)
    assert lib().orc_roi_pool_fwd(xp, rp, N, Cc, H, W, R, PH, PW, C.c_float(scale), out.ctypes.data_as(C.POINTER(C.c_float)),
                                  am.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    return out, am


def roi_pool_bwd(grad_out, rois, argmax, in_shape) -> np.ndarray:
    g, gp = _f(grad_out); r, rp = _f(rois)
    am = np.ascontiguousarray(argmax, dtype=np.int32)
    N, Cc, H, W = in_shape; R, _, PH, PW = g.shape
    gi = np.zeros(in_shape, dtype=np.float32)
    assert lib().orc_roi_pool_bwd(gp, rp, am.ctypes.data_as(C.POINTER(C.c_int32)), N, Cc, H, W, R, PH, PW,
                                  gi.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return gi


def apply_deltas(deltas, boxes, weights, scale_clamp) -> np.ndarray:
    d, dp = _f(deltas); b, bp = _f(boxes); w, wp = _f(np.asarray(weights, dtype=np.float32))
    R = b.shape[0]; k = d.shape[1] // 4
    out = np.zeros_like(d)
    assert lib().orc_apply_deltas(dp, bp, C.c_int64(R), k, wp, C.c_float(scale_clamp), out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return out


def softmax(x) -> np.ndarray:
    x, xp = _f(x); out = np.zeros_like(x)
    assert lib().orc_softmax(xp, C.c_int64(x.shape[0]), x.shape[1], out.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return out


def ema_f32(student, teacher, keep_rate: float) -> np.ndarray:
    s, sp = _f(student); t = np.array(teacher, dtype=np.float32, copy=True, order="C")
    assert lib().orc_ema_f32(sp, t.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(s.size), C.c_double(keep_rate)) == 0
    return t


def ema_i64(student, teacher, keep_rate: float) -> np.ndarray:
    s = np.ascontiguousarray(student, dtype=np.int64); t = np.array(teacher, dtype=np.int64, copy=True, order="C")
    assert lib().orc_ema_i64(s.ctypes.data_as(C.POINTER(C.c_int64)), t.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(s.size),
                             C.c_double(keep_rate)) == 0
    return t


def bn_stats(x):
    x, xp = _f(x); N, Cc = x.shape[:2]; HW = int(np.prod(x.shape[2:]))
    m = np.zeros(Cc, np.float32); v = np.zeros(Cc, np.float32)
    assert lib().orc_bn_stats(xp, N, Cc, C.c_int64(HW), m.ctypes.data_as(C.POINTER(C.c_float)), v.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return m, v


def bn_update_running(mean, var_b, n, momentum, running_mean, running_var):
    m, mp = _f(mean); v, vp = _f(var_b)
    rm = np.array(running_mean, dtype=np.float32, copy=True); rv = np.array(running_var, dtype=np.float32, copy=True)
    assert lib().orc_bn_update_running(mp, vp, m.shape[0], C.c_double(n), C.c_double(momentum), rm.ctypes.data_as(C.POINTER(C.c_float)),
                                       rv.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return rm, rv


def bn_apply(x, mean, var_b, weight, bias, eps):
    x, xp = _f(x); m, mp = _f(mean); v, vp = _f(var_b); w, wp = _f(weight); b, bp = _f(bias)
    N, Cc = x.shape[:2]; HW = int(np.prod(x.shape[2:])); y = np.zeros_like(x)
    assert lib().orc_bn_apply(xp, N, Cc, C.c_int64(HW), mp, vp, wp, bp, C.c_double(eps), y.ctypes.data_as(C.POINTER(C.c_float))) == 0
    return y
