"""Drop-in for dro_sfm.losses.supervised_loss.SupervisedDepthPoseLoss
(reference: dro_sfm/losses/supervised_loss.py:201-371).

The reprojection pose loss (get_ref_coords / calc_pose_loss, :279-325) runs as one fused kernel per
direction, the masked-L1 depth term (:244-277) as one more.
"""
import torch

from .. import ops
from ..geometry.pose import Pose
from .loss_base import LossBase, ProgressiveScaling


def _materialize_rows(rows):
    """[B,4,4] matrices of rows of poses, same nesting; one conversion launch for all lazy euler vectors."""
    flat = Pose.materialize([p for row in rows for p in row])
    out, k = [], 0
    for row in rows:
        out.append(flat[k:k + len(row)])
        k += len(row)
    return out


class SupervisedDepthPoseLoss(LossBase):
    def __init__(self, supervised_method='sparse-l1', supervised_num_scales=4, progressive_scaling=0.0,
                 min_depth=0.1, max_depth=100, **kwargs):
        super().__init__()
        if supervised_method not in ('sparse-l1', 'l1'):
            # calculate_loss of the reference never calls self.loss_func; the name is only validated
            raise ValueError('Unknown supervised loss {}'.format(supervised_method))
        self.supervised_method = supervised_method
        self.n = supervised_num_scales
        self.progressive_scaling = ProgressiveScaling(progressive_scaling, self.n)
        self.min_depth = min_depth
        self.max_depth = max_depth

    @property
    def logs(self):
        return {'supervised_num_scales': self.n}

    def calculate_loss(self, inv_depths, gt_inv_depths):
        """gamma-weighted masked L1 on inverse depth (supervised_loss.py:244-277), one fused kernel.
        ``gt_inv_depths`` is the per-prediction list the reference builds with match_scales; all entries are
        the same full-resolution map."""
        return ops.sup_depth_loss(list(inv_depths[:self.n]), gt_inv_depths[0], self.min_depth, self.max_depth, 0.85)

    def get_ref_coords(self, pose, K, ref_K, depth, scale_factor, device):
        """Projected coordinates and their in-range mask (supervised_loss.py:279-291)."""
        mat = pose.mat if isinstance(pose, Pose) or hasattr(pose, "mat") else pose
        return ops.warp_coords(depth, mat, K, ref_K, scale_factor, True, want_mask=True)

    def calc_pose_loss(self, pred_poses, gt_pose_context, gt_depth, K, ref_K):
        """Reprojection loss of the predicted poses on GT depth (supervised_loss.py:293-325)."""
        preds = _materialize_rows([pv[:self.n] for pv in pred_poses])
        gts = [g.mat if hasattr(g, "mat") else g for g in gt_pose_context]
        return ops.reproj_pose_loss(preds, gts, gt_depth, K, ref_K, self.min_depth, self.max_depth, 0.85)

    def forward(self, image, context, inv_depths, gt_inv_depth, gt_pose_context, K, ref_K, poses,
                return_logs=False, progress=0.0):
        """Same contract as the reference (supervised_loss.py:328-371)."""
        self.n = len(inv_depths)
        for d in inv_depths:
            if tuple(d.shape) != tuple(gt_inv_depth.shape):
                raise NotImplementedError("dro_sfm_b200: predictions must be at the ground-truth resolution")
        loss_depth = self.calculate_loss(inv_depths, [gt_inv_depth] * self.n)
        # calc_pose_loss (and the reference, supervised_loss.py:306-312) use the first self.n predictions of every view
        # (one conversion launch for all V x n euler vectors that are still lazy, see Pose.materialize)
        preds = _materialize_rows([pv[:self.n] for pv in poses])
        gts = [g.mat if hasattr(g, "mat") else g for g in gt_pose_context]
        # inv2depth of the GT map is fused into the kernel
        loss_pose = ops.reproj_pose_loss(preds, gts, gt_inv_depth, K, ref_K, self.min_depth, self.max_depth, 0.85,
                                         inverse_depth=True)
        self.add_metric('depth_loss', loss_depth)
        self.add_metric('pose_loss', loss_pose)
        self.add_metric('all_loss', loss_depth + loss_pose)
        loss = loss_depth + loss_pose
        return {'loss': loss.unsqueeze(0), 'metrics': self.metrics}
