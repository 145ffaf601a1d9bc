"""Drop-in for the recurrent optimiser's feature-metric cost calls
(reference: dro_sfm/networks/depth_pose/DepthPoseNet.py:76-105).

Same argument lists as ``DepthPoseNet.get_cost_each`` / ``depth_cost_calc`` so that the closures
built in ``DepthPoseNet.forward`` (:159-167) and invoked from ``update.py:161,189`` keep working.
Each call is ONE kernel launch: pose vector -> matrix, intrinsics cast/scaling, inv2depth, the
coordinate chain, the bilinear gather of all source views and the squared-difference reduction are
fused; nothing but the [B,C,h,w] cost map is written.
"""
from .. import ops
from .. import _lib as L


def _cost(pose_list, fmap, fmaps_ref, depth, K, ref_K, scale_factor, inverse_depth):
    # poses arrive as [B,6] euler vectors (Pose.from_vec(pose, "euler") in the reference)
    return ops.feat_cost(depth, fmap, list(fmaps_ref), list(pose_list), K, ref_K, scale_factor,
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


class FeatureMetricCost:
    """Mixin with the reference's method names; ``patch.install`` grafts it onto DepthPoseNet."""

    def get_cost_each(self, pose, fmap, fmap_ref, depth, K, ref_K, scale_factor):
        return get_cost_each(pose, fmap, fmap_ref, depth, K, ref_K, scale_factor)

    def depth_cost_calc(self, inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale_factor):
        return depth_cost_calc(inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale_factor)
