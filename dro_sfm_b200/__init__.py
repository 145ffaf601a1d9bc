"""dro_sfm_b200 -- B200-native (sm_100a) dense depth-pose warping path of dro-sfm.

Hand-written CUDA kernels behind a C ABI (include/drosfm_b200.h) and the reference's own Python
operator surface:

    from dro_sfm_b200.geometry import Camera, Pose, view_synthesis
    from dro_sfm_b200.losses import MultiViewPhotometricDecayLoss, SupervisedDepthPoseLoss
    from dro_sfm_b200.networks import get_cost_each, depth_cost_calc

There is no CPU / PyTorch fallback: the operators raise if the CUDA library is missing or the
tensors are not on a CUDA device.
"""
__version__ = "0.1.0"
