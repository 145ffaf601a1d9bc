"""ctypes binding of libdrosfm_b200.so (the C ABI declared in include/drosfm_b200.h).

There is no CPU fallback: every operator of this package goes through this module, and loading
fails loudly when the CUDA library has not been built (``python -m dro_sfm_b200.build``).
PyTorch is used for device memory and streams only.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("DROSFM_SO") or os.path.join(_HERE, "libdrosfm_b200.so")     # DROSFM_SO: alternative build (experiments)

ABI_VERSION = 4
MAX_VIEWS = 8
MAX_PREDS = 16
MAX_COST_JOBS = 9

POSE_IDENTITY, POSE_MAT4, POSE_EULER6 = 0, 1, 2
PAD_ZEROS, PAD_BORDER = 0, 1
F32, F64 = 0, 1
DEPTH, INV_DEPTH, DISP = 0, 1, 2
REDUCE_MIN, REDUCE_MEAN = 0, 1
NCHW, NHWC = 0, 1
ACCUMULATE_FMAP = 1
PHOTO_WARPED_READY, PHOTO_NO_ADJOINT, PHOTO_FUSE_BWD = 1, 2, 4
SLOT_BYTES = 128

_vp = ctypes.c_void_p
_int = ctypes.c_int
_f32 = ctypes.c_float
_pp = ctypes.POINTER(ctypes.c_void_p)


class Cams(ctypes.Structure):
    """drosfm_cams_t"""
    _fields_ = [("K", _vp), ("Kref", _vp), ("k_dtype", ctypes.c_int32), ("sx", _f32), ("sy", _f32),
                ("Twc", _vp), ("pose", _vp), ("pose_kind", ctypes.c_int32)]


class PhotoOpts(ctypes.Structure):
    """drosfm_photo_opts_t"""
    _fields_ = [("ssim_w", _f32), ("C1", _f32), ("C2", _f32), ("padding", ctypes.c_int32),
                ("reduce_op", ctypes.c_int32), ("automask", ctypes.c_int32), ("gamma", _f32), ("clip_loss", _f32),
                ("clip_scratch", _vp)]


class CostJob(ctypes.Structure):
    """drosfm_cost_job_t"""
    _fields_ = [("fmap", _vp), ("fmap_ref", _pp), ("depth", _vp), ("depth_kind", ctypes.c_int32), ("n_views", ctypes.c_int32),
                ("poses", _pp), ("cost", _vp), ("disp_min", _f32), ("disp_range", _f32)]


class CostJobGrads(ctypes.Structure):
    """drosfm_cost_job_grads_t"""
    _fields_ = [("g_cost", _vp), ("g_fmap", _vp), ("g_fmap_ref", _pp), ("g_depth", _vp), ("g_poses", _pp),
                ("flags", ctypes.c_int32)]


_cp = ctypes.POINTER(Cams)
_op = ctypes.POINTER(PhotoOpts)
_jp = ctypes.POINTER(CostJob)
_gp = ctypes.POINTER(CostJobGrads)

# name -> argtypes; every symbol include/drosfm_b200.h declares (tests check the export list)
SIGNATURES = {
    "drosfm_version": ([], _int),
    "drosfm_last_error": ([], ctypes.c_char_p),
    "drosfm_launch_count": ([], ctypes.c_ulonglong),
    "drosfm_ws_bytes": ([_int], ctypes.c_size_t),
    "drosfm_pose_vec2mat_fwd": ([_vp, _vp, _int, _vp], _int),
    "drosfm_pose_vec2mat_bwd": ([_vp, _vp, _vp, _int, _vp], _int),
    "drosfm_reconstruct_fwd": ([_vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp], _int),
    "drosfm_reconstruct_bwd": ([_vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp], _int),
    "drosfm_project_fwd": ([_vp, _vp, _int, _vp, _vp, _int, _int, _int, _int, _vp], _int),
    "drosfm_project_bwd": ([_vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp], _int),
    "drosfm_warp_coords_fwd": ([_vp, _int, _cp, _vp, _vp, _int, _int, _int, _int, _vp], _int),
    "drosfm_warp_coords_bwd": ([_vp, _vp, _int, _cp, _vp, _vp, _vp, _int, _int, _int, _int, _vp], _int),
    "drosfm_selftest_rcp": ([_vp, _vp], _int),
    "drosfm_grid_gather_fwd": ([_vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_grid_gather_bwd": ([_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_view_synthesis_fwd": ([_vp, _vp, _int, _cp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_view_synthesis_bwd": ([_vp, _vp, _vp, _int, _cp, _vp, _vp, _vp, _vp,
                                   _int, _int, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_feat_cost_fwd": ([_vp, _pp, _vp, _int, _cp, _pp, _int, _vp, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_feat_cost_bwd": ([_vp, _vp, _pp, _vp, _int, _cp, _pp, _int, _vp, _pp, _vp, _pp, _vp,
                              _int, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_feat_cost_batch_fwd": ([_jp, _int, _cp, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_feat_cost_batch_bwd": ([_jp, _gp, _int, _cp, _vp, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_automask_fwd": ([_vp, _pp, _int, _op, _vp, _int, _int, _int, _vp], _int),
    "drosfm_photometric_fwd": ([_vp, _pp, _int, _pp, _int, _int, _cp, _pp, _vp, _op, _vp, _vp, _vp, _vp, _vp,
                                _int, _int, _int, _int, _vp], _int),
    "drosfm_photometric_bwd": ([_vp, _vp, _pp, _int, _pp, _int, _int, _cp, _pp, _vp, _op, _pp, _pp, _vp, _vp, _vp,
                                _int, _int, _int, _int, _vp], _int),
    "drosfm_warp_sources_fwd": ([_pp, _int, _pp, _int, _int, _cp, _pp, _int, _vp, _vp, _int, _int, _int, _vp], _int),
    "drosfm_warp_sources_bwd": ([_vp, _pp, _int, _pp, _int, _int, _cp, _pp, _int, _vp, _vp, _pp, _pp, _vp, _int,
                                 _int, _int, _int, _vp], _int),
    "drosfm_smoothness_fwd": ([_vp, _pp, _int, _f32, _vp, _vp, _vp, _vp, _int, _int, _int, _vp], _int),
    "drosfm_smoothness_bwd": ([_vp, _vp, _pp, _int, _f32, _vp, _pp, _int, _vp, _int, _int, _int, _vp], _int),
    "drosfm_sup_depth_loss_fwd": ([_vp, _pp, _int, _f32, _f32, _f32, _vp, _vp, _int, _int, _int, _vp], _int),
    "drosfm_sup_depth_loss_bwd": ([_vp, _vp, _pp, _int, _f32, _f32, _f32, _pp, _int, _int, _int, _vp], _int),
    "drosfm_relayout": ([_vp, _vp, _int, _int, _int, _int, _int, _vp], _int),
    "drosfm_post_process_inv_depth": ([_vp, _vp, _vp, _int, _int, _int, _int, _vp], _int),
    "drosfm_eval_ws_bytes": ([_int], ctypes.c_size_t),
    "drosfm_depth_metrics": ([_vp, _vp, _int, _int, _int, _int, _int, _f32, _f32, _int, _int, _vp, _vp, _vp], _int),
    "drosfm_images_u8_to_f32": ([_vp, _vp, ctypes.c_size_t, _vp], _int),
    "drosfm_upsample_depth_fwd": ([_vp, _vp, _vp, _int, _int, _int, _int, _f32, _f32, _vp], _int),
    "drosfm_upsample_depth_bwd": ([_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _f32, _vp], _int),
    "drosfm_reproj_loss_fwd": ([_vp, _int, _cp, _pp, _pp, _int, _int, _f32, _f32, _f32, _vp, _vp,
                                _int, _int, _int, _vp], _int),
    "drosfm_reproj_loss_bwd": ([_vp, _vp, _int, _cp, _pp, _pp, _int, _int, _f32, _f32, _f32, _pp, _vp,
                                _int, _int, _int, _vp], _int),
}

_lib = None
_lock = threading.Lock()


class _TimedLib:
    """Proxy that brackets every kernel entry point with CUDA events on the launching stream
    (bench.py uses it for the per-kernel roofline; never active on the product path)."""

    def __init__(self, handle):
        self._h = handle
        self.records = []      # (name, args, start_event, end_event)

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.endswith(("_fwd", "_bwd")):
            return fn

        def timed(*args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            self.records.append((name, args, s, e))
            return rc
        return timed


_timed = None


def profile_begin():
    """Start per-call CUDA-event timing of the C-ABI entry points."""
    global _timed
    _timed = _TimedLib(lib())
    return _timed


def profile_end():
    """Stop timing; returns [(name, args, milliseconds)] after synchronising."""
    global _timed
    t, _timed = _timed, None
    torch.cuda.synchronize()
    return [(name, args, s.elapsed_time(e)) for name, args, s, e in t.records]


def lib():
    """The loaded library; raises RuntimeError if it is missing (no fallback path exists)."""
    global _lib
    if _timed is not None:
        return _timed
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise RuntimeError(
                        "dro_sfm_b200: %s not found. Build it with `python -m dro_sfm_b200.build` "
                        "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % SO_PATH)
                handle = ctypes.CDLL(SO_PATH)
                missing = [n for n in SIGNATURES if not hasattr(handle, n)]
                if missing:
                    raise RuntimeError("dro_sfm_b200: %s does not export %s (stale build?)" % (SO_PATH, missing))
                for name, (argtypes, restype) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.argtypes = argtypes
                    fn.restype = restype
                if handle.drosfm_version() != ABI_VERSION:
                    raise RuntimeError("dro_sfm_b200: ABI version mismatch (library %d, binding %d)"
                                       % (handle.drosfm_version(), ABI_VERSION))
                _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().drosfm_last_error().decode("utf-8", "replace")
        kind = ValueError if rc < 0 else RuntimeError
        raise kind("drosfm %s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    """Host array of device pointers (const float* const*); entries may be None."""
    arr = (ctypes.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return ctypes.cast(arr, _pp)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dro_sfm_b200 operators run on CUDA tensors only (got a %s tensor); "
                               "there is no CPU fallback" % t.device)


def f32c(t):
    """float32 contiguous view/copy of t."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def k_arg(K):
    """Intrinsics for the kernels: float64 is consumed directly (rounded in-kernel == K.float())."""
    if K.dtype not in (torch.float32, torch.float64):
        K = K.float()
    K = K.contiguous()
    return K, (F64 if K.dtype == torch.float64 else F32)


def make_cams(K, Kref, sx=1.0, sy=None, Twc=None, pose=None, pose_kind=POSE_IDENTITY):
    """Builds a drosfm_cams_t; returns (struct, keepalive tuple)."""
    K, kd = k_arg(K)
    if Kref is K:
        Kref_c, kd2 = K, kd
    else:
        Kref_c, kd2 = k_arg(Kref)
        if kd2 != kd:
            K, Kref_c, kd = K.float().contiguous(), Kref_c.float().contiguous(), F32
    c = Cams()
    c.K, c.Kref, c.k_dtype = K.data_ptr(), Kref_c.data_ptr(), kd
    c.sx = float(sx)
    c.sy = float(sx if sy is None else sy)
    c.Twc = None if Twc is None else Twc.data_ptr()
    c.pose = None if pose is None else pose.data_ptr()
    c.pose_kind = int(pose_kind)
    return c, (K, Kref_c, Twc, pose)


_workspaces = {}


_side_streams = {}


def side_stream(device):
    """The library's second stream on `device` (independent parts of a loss run there, joined before returning)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        st = _side_streams[idx] = torch.cuda.Stream(device)
    return st


_eval_ws = {}


def eval_workspace(device, B):
    """Zero-initialised, self-cleaning workspace of drosfm_depth_metrics, one per (device, stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    need = int(lib().drosfm_eval_ws_bytes(int(B)))
    ws = _eval_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _eval_ws[key] = torch.zeros(need, dtype=torch.uint8, device=device)
    return ws


def workspace(device, slots):
    """Zero-initialised, self-cleaning reduction workspace, one per (device, stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    need = int(lib().drosfm_ws_bytes(max(64, int(slots))))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need * 2, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws
