"""Drop-in for the recurrent optimiser's feature-metric cost calls
(reference: dro_sfm/networks/depth_pose/DepthPoseNet.py:76-105).

Same argument lists as ``DepthPoseNet.get_cost_each`` / ``depth_cost_calc`` so that the closures
built in ``DepthPoseNet.forward`` (:159-167) and invoked from ``update.py:161,189`` keep working.
Each call is ONE kernel launch: pose vector -> matrix, intrinsics cast/scaling, inv2depth, the
coordinate chain, the bilinear gather of all source views and the squared-difference reduction are
fused; nothing but the [B,C,h,w] cost map is written.
"""
import os
import weakref

import torch

from .. import ops
from .. import _lib as L

# The encoder hands over NCHW feature maps, and the same maps are consumed by all 2*V*T cost calls of a
# forward pass (DepthPoseNet.py:113-115,159-167).  The channels-last kernels move 512 contiguous bytes per
# tap instead of 4, so each distinct map is converted ONCE per forward (an autograd-visible
# ``.contiguous(memory_format=channels_last)``; its gradient comes back through the same node) and the copy
# is remembered for as long as the original tensor object is alive and unmodified.  The returned cost map is
# a logical [B,C,h,w] tensor in channels_last storage, which is also what cuDNN prefers for the 1x1
# convolution that consumes it (update.py:81).  DROSFM_COST_LAYOUT=nchw keeps the caller's layout.
_LAYOUT = os.environ.get("DROSFM_COST_LAYOUT", "nhwc").lower()
_cl_cache = {}


def _channels_last(t):
    if _LAYOUT != "nhwc" or t.dim() != 4 or t.shape[1] % 4 != 0 or t.shape[1] < 32:
        return t
    if t.is_contiguous(memory_format=torch.channels_last):
        return t
    key = id(t)
    hit = _cl_cache.get(key)
    if hit is not None:
        ref, version, grad_mode, converted = hit
        # valid for the same tensor object, unmodified, in the same grad mode, and only until the backward
        # pass of the step that created it has run (its autograd node and gradient sink are then spent)
        if (ref() is t and version == t._version and grad_mode == torch.is_grad_enabled()
                and not converted._drosfm_sink_state.consumed):
            return converted
    # autograd-visible conversion whose backward hands over the in-kernel sum of all cost gradients
    converted = ops.to_channels_last_sink(t)
    if len(_cl_cache) > 64:
        for k in [k for k, (r, _, _, _) in _cl_cache.items() if r() is None]:
            del _cl_cache[k]
    _cl_cache[key] = (weakref.ref(t, lambda _r, k=key: _cl_cache.pop(k, None)), t._version, torch.is_grad_enabled(), converted)
    return converted


def split_feature_maps(stacked, sizes):
    """torch.split(stacked, sizes, dim=0) of the encoder's output for [target, source_1..V] (DepthPoseNet.py:113-118), with
    the channels_last conversion the cost kernels want done ONCE for the stacked tensor: one conversion launch, one
    zero-filled gradient buffer and one back-conversion per step instead of one per map (ops.split_channels_last_sink)."""
    t = stacked
    usable = (_LAYOUT == "nhwc" and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0
              and t.shape[1] >= 32 and t.is_contiguous())
    if not usable:
        return list(torch.split(stacked, list(sizes), dim=0))
    return ops.split_channels_last_sink(stacked, sizes)


def _cost(pose_list, fmap, fmaps_ref, depth, K, ref_K, scale_factor, inverse_depth):
    # poses arrive as [B,6] euler vectors (Pose.from_vec(pose, "euler") in the reference)
    fmap = _channels_last(fmap)
    fmaps_ref = [_channels_last(f) for f in fmaps_ref]
    return ops.feat_cost(depth, fmap, fmaps_ref, list(pose_list), K, ref_K, scale_factor,
                         inverse_depth=inverse_depth)


def get_cost_each(pose, fmap, fmap_ref, depth, K, ref_K, scale_factor):
    """(fmap - warp(fmap_ref))**2 -> [B,C,h,w]   (DepthPoseNet.py:76-96)."""
    return _cost([pose], fmap, [fmap_ref], depth, K, ref_K, scale_factor, False)


def depth_cost_calc(inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale_factor):
    """mean over views of get_cost_each on inv2depth(inv_depth)   (DepthPoseNet.py:98-105)."""
    pose_list, fmaps_ref = list(pose_list), list(fmaps_ref)
    if len(pose_list) <= L.MAX_VIEWS:
        return _cost(pose_list, fmap, fmaps_ref, inv_depth, K, ref_K, scale_factor, True)
    # more views than one launch takes: accumulate chunks (sum of per-chunk means, reweighted)
    total = None
    for s in range(0, len(pose_list), L.MAX_VIEWS):
        ps, fs = pose_list[s:s + L.MAX_VIEWS], fmaps_ref[s:s + L.MAX_VIEWS]
        part = _cost(ps, fmap, fs, inv_depth, K, ref_K, scale_factor, True) * (len(ps) / len(pose_list))
        total = part if total is None else total + part
    return total


def cost_batch(jobs, K, ref_K, scale_factor):
    """Several independent cost evaluations in ONE kernel launch (forward) and one more (backward).

    jobs: (depth_or_inv_depth, fmap, fmaps_ref, pose_list, kind) tuples -- a depth_cost_calc call is
    (inv_depth, fmap, fmaps_ref, pose_list, True), a get_cost_each call (depth, fmap, [fmap_ref], [pose], False);
    kind = ("disp", min_depth, max_depth) hands over the network's raw disparity and lets the kernel apply
    scale_inv_depth / disp_to_depth (DepthPoseNet.py:38-41, layers.py:11-20) itself.
    Used by the lock-step schedule (networks/lockstep.py); results equal the individual calls."""
    prepared = [(d, _channels_last(f), [_channels_last(r) for r in frs], list(ps), inv) for d, f, frs, ps, inv in jobs]
    return ops.feat_cost_batch(prepared, K, ref_K, scale_factor)


def upsample_depth(depth, mask, ratio=8, disp_range=None):
    """Convex up-sampling of the low-resolution inverse depth (DepthPoseNet.py:63-74), one fused kernel; with
    disp_range = (min_depth, max_depth) the disp_to_depth scaling that follows it in DepthPoseNet.forward is its epilogue."""
    return ops.upsample_depth(depth, mask, ratio, disp_range)


class FeatureMetricCost:
    """Mixin with the reference's method names; ``patch.install`` grafts it onto DepthPoseNet."""

    def upsample_depth(self, depth, mask, ratio=8):
        return upsample_depth(depth, mask, ratio)        # the reference's signature (DepthPoseNet.py:63)

    def get_cost_each(self, pose, fmap, fmap_ref, depth, K, ref_K, scale_factor):
        return get_cost_each(pose, fmap, fmap_ref, depth, K, ref_K, scale_factor)

    def depth_cost_calc(self, inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale_factor):
        return depth_cost_calc(inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale_factor)
