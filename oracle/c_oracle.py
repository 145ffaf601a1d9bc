"""ctypes binding of oracle/coords_oracle.c (TEST INFRASTRUCTURE; numpy in, numpy out)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdrosfm_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_u8 = ctypes.POINTER(ctypes.c_ubyte)


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "coords_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t=_f):
    return None if a is None else a.ctypes.data_as(t)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def inv2depth(inv):
    inv = _c(inv)
    out = np.empty_like(inv)
    lib().drosfm_oracle_inv2depth(_p(inv), _p(out), ctypes.c_size_t(inv.size))
    return out


def reconstruct(depth, K, Twc=None):
    """depth [B,1,H,W], K [B,3,3] float32, Twc [B,4,4] or None (frame 'c') -> [B,3,H,W]."""
    depth, K, Twc = _c(depth), _c(K), _c(Twc)
    B, _, H, W = depth.shape
    out = np.empty((B, 3, H, W), np.float32)
    lib().drosfm_oracle_reconstruct(_p(depth), _p(K), _p(Twc), _p(out), B, H, W)
    return out


def project(points, K, Tcw=None, normalize=True):
    points, K, Tcw = _c(points), _c(K), _c(Tcw)
    B, _, H, W = points.shape
    out = np.empty((B, H, W, 2), np.float32)
    lib().drosfm_oracle_project(_p(points), _p(K), _p(Tcw), _p(out), B, H, W, int(normalize))
    return out


def warp_coords(depth, K, Kref, T, sx=1.0, sy=None, normalize=True, want_mask=False):
    """Fused reconstruct(identity target camera) -> project(source camera at pose T)."""
    depth, K, Kref, T = _c(depth), _c(K), _c(Kref), _c(T)
    sy = sx if sy is None else sy
    B, _, H, W = depth.shape
    uv = np.empty((B, H, W, 2), np.float32)
    mask = np.empty((B, H, W, 2), np.uint8) if want_mask else None
    lib().drosfm_oracle_warp_coords(_p(depth), _p(K), _p(Kref), _p(T), ctypes.c_float(sx), ctypes.c_float(sy),
                                    _p(uv), _p(mask, _u8), B, H, W, int(normalize))
    return (uv, mask.astype(bool)) if want_mask else uv


def euler_from_trig(trig, angle_z, fma=False):
    """(sin x, cos x, sin y, cos y, sin z, cos z) [N,6] + z angles [N] -> R [N,3,3] in euler2mat's operation order.
    fma=False: the accumulation of torch's CPU bmm; fma=True: the FMA chain of a CUDA bmm (see coords_oracle.c)."""
    trig, angle_z = _c(trig), _c(angle_z)
    N = trig.shape[0]
    out = np.empty((N, 3, 3), np.float32)
    fn = lib().drosfm_oracle_euler_from_trig
    for i in range(N):
        fn(_p(trig[i]), ctypes.c_float(float(angle_z[i])), int(bool(fma)), _p(out[i]))
    return out
