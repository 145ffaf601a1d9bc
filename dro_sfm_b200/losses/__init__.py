from .loss_base import LossBase, ProgressiveScaling
from .multiview_photometric_loss_mf import MultiViewPhotometricDecayLoss, SSIM
from .supervised_loss import SupervisedDepthPoseLoss

__all__ = ["LossBase", "ProgressiveScaling", "MultiViewPhotometricDecayLoss", "SSIM", "SupervisedDepthPoseLoss"]
