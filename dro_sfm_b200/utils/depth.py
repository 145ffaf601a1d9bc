"""Drop-ins for the evaluation helpers of dro_sfm.utils.depth (reference: dro_sfm/utils/depth.py:230-340), the tail of
``ModelWrapper.evaluate_depth`` (models/model_wrapper.py:355-399).

``post_process_inv_depth`` is one launch; ``compute_depth_metrics`` replaces the per-sample Python loop (boolean
indexing, sort-based ``torch.median``, ~60 ATen launches per sample and mode) by an exact radix selection of the median
ratio plus one fused pass over the batch -- no host synchronisation, the result stays on the device.
"""
import torch

from .. import ops


def post_process_inv_depth(inv_depth, inv_depth_flipped, method='mean'):
    """Fuse a prediction with the prediction of the horizontally flipped image (depth.py:230-258)."""
    return ops.post_process_inv_depth(inv_depth, inv_depth_flipped, method)


def compute_depth_metrics(config, gt, pred, use_gt_scale=True):
    """[abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, SILog, iabs_diff] averaged over the batch (depth.py:261-340).

    config: object with ``crop`` ('garg' | 'eigen_nyu' | anything else = none), ``min_depth``, ``max_depth``."""
    crop = getattr(config, "crop", "")
    return ops.depth_metrics(gt, pred, float(config.min_depth), float(config.max_depth), crop, bool(use_gt_scale)).type_as(gt)
