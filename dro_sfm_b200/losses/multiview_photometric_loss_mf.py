"""Drop-in for dro_sfm.losses.multiview_photometric_loss_mf
(reference: dro_sfm/losses/multiview_photometric_loss_mf.py:15-361).

``MultiViewPhotometricDecayLoss(**config.model.loss).forward(image, context, inv_depths, K, ref_K,
poses, return_logs, progress)`` keeps its signature, return structure and metric keys.  Underneath,
the ~4 800 ATen launches of the reference (V=2, n=4) become: one auto-mask pass, one fused
warp+SSIM+L1+min+decay forward, two smoothness passes, and in backward one photometric and one
smoothness kernel.
"""
import torch

from .. import ops
from .loss_base import LossBase, ProgressiveScaling


def SSIM(x, y, C1=1e-4, C2=9e-4, kernel_size=3, stride=1):
    """Stand-alone SSIM map (multiview_photometric_loss_mf.py:15-54).  Host-side helper kept for API
    parity (visualisation / metrics); the training loss evaluates SSIM inside the fused kernel."""
    import torch.nn.functional as F
    if kernel_size != 3 or stride != 1:
        raise NotImplementedError("dro_sfm_b200: SSIM supports kernel_size=3, stride=1")
    x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x.pow(2), 3, 1) - mu_x.pow(2)
    sigma_y = F.avg_pool2d(y.pow(2), 3, 1) - mu_y.pow(2)
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    return ((2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)) / ((mu_x.pow(2) + mu_y.pow(2) + C1) * (sigma_x + sigma_y + C2))


class MultiViewPhotometricDecayLoss(LossBase):
    """Self-supervised multi-view photometric loss with gamma-decay over the recurrent predictions.

    Constructor arguments are those of the reference (lines 92-95); unknown keys of the config node
    are swallowed by **kwargs exactly as there."""

    def __init__(self, num_scales=4, ssim_loss_weight=0.85, occ_reg_weight=0.1, smooth_loss_weight=0.1,
                 C1=1e-4, C2=9e-4, photometric_reduce_op='mean', disp_norm=True, clip_loss=0.5,
                 progressive_scaling=0.0, padding_mode='zeros', automask_loss=False, **kwargs):
        super().__init__()
        self.n = 1
        self.ssim_loss_weight = ssim_loss_weight
        self.occ_reg_weight = occ_reg_weight
        self.smooth_loss_weight = smooth_loss_weight
        self.C1 = C1
        self.C2 = C2
        self.photometric_reduce_op = photometric_reduce_op
        self.disp_norm = disp_norm
        self.clip_loss = clip_loss
        self.padding_mode = padding_mode
        self.automask_loss = automask_loss
        self.progressive_scaling = ProgressiveScaling(progressive_scaling, self.n)
        # diagnostic of the last forward: [n,B,H,W] uint8, the source view that won the per-pixel min (255: an un-warped
        # map, i.e. the pixel is auto-masked); None for the 'mean' reduce op
        self.last_selection = None
        if self.automask_loss:
            assert self.photometric_reduce_op == 'min', \
                'For automasking only the min photometric_reduce_op is supported.'
        if self.photometric_reduce_op not in ('min', 'mean'):
            raise NotImplementedError('Unknown photometric_reduce_op: {}'.format(self.photometric_reduce_op))

    @property
    def logs(self):
        return {'num_scales': self.n}

    @staticmethod
    def _pose_mats(poses, n):
        """poses: per view either a list of n Pose / tensors (training) or one Pose (evaluation)."""
        out = []
        for pv in poses:
            if not isinstance(pv, (list, tuple)):
                pv = [pv] * n
            out.append([p.kernel_arg() if hasattr(p, "kernel_arg") else (p.mat if hasattr(p, "mat") else p) for p in pv])
        # one encoding per call: if euler vectors and matrices are mixed, materialise the matrices
        if len({t.dim() for row in out for t in row}) > 1:
            out = [[(p.mat if hasattr(p, "mat") else p) for p in (pv if isinstance(pv, (list, tuple)) else [pv] * n)]
                   for pv in poses]
        return out

    def forward(self, image, context, inv_depths, K, ref_K, poses, return_logs=False, progress=0.0):
        """Same contract as the reference (lines 303-361): returns {'loss': [1], 'metrics': {...}}."""
        self.n = len(inv_depths)
        total, terms, self.last_selection = ops.photometric_loss(
            image, list(context), list(inv_depths), K, ref_K, self._pose_mats(poses, self.n),
            ssim_w=self.ssim_loss_weight, C1=self.C1, C2=self.C2, reduce_op=self.photometric_reduce_op,
            padding_mode=self.padding_mode, automask=self.automask_loss, smooth_w=self.smooth_loss_weight,
            gamma=0.85, inverse_depth=True, want_selection=True, clip=self.clip_loss)
        if self.smooth_loss_weight > 0.0:
            self.add_metric('smoothness_loss', terms[1])
            # reference quirk: 'photometric_loss' aliases the tensor that `loss += smoothness` then
            # updates in place (lines 268, 356), so the logged value is the total
            self.add_metric('photometric_loss', total[0])
        else:
            self.add_metric('photometric_loss', terms[0])
        return {'loss': total, 'metrics': self.metrics}
