"""Install the B200 operators under an importable reference tree.

    import dro_sfm_b200.patch as patch
    patch.install()          # after this, `import dro_sfm` code paths hit the CUDA kernels

Replaces (by name, keeping signatures) the hot-path symbols of the reference:
  dro_sfm.geometry.camera.Camera, dro_sfm.geometry.pose.Pose,
  dro_sfm.geometry.camera_utils.view_synthesis,
  dro_sfm.losses.multiview_photometric_loss_mf.MultiViewPhotometricDecayLoss,
  dro_sfm.losses.supervised_loss.SupervisedDepthPoseLoss,
  dro_sfm.utils.depth.post_process_inv_depth / compute_depth_metrics (evaluation path),
  DepthPoseNet.get_cost_each / depth_cost_calc / upsample_depth, and (lockstep=True, the default) DepthPoseNet.forward
  by the lock-step schedule of networks/lockstep.py: the same sub-modules and arithmetic, with the 1 + V cost
  evaluations of every inner step of the recurrent optimiser submitted as one kernel launch.
Modules that did `from x import Name` before install() keep their old binding, so the known importers
(SelfSupModelMF.py:3, SupModelMF.py:3, SemiSupModelMF.py:3-4, DepthPoseNet.py:11, SfmModelMF.py) are
re-bound too.  Nothing else of the reference is touched: trainer, configs and checkpoints are as-is
(the hot-path modules own no parameters or buffers).
"""
import importlib
import sys


def _rebind(module_name, **symbols):
    try:
        mod = importlib.import_module(module_name)
    except Exception:
        return False
    for name, value in symbols.items():
        if hasattr(mod, name):
            setattr(mod, name, value)
    return True


def install(lockstep=True):
    """Returns the list of reference modules that were patched."""
    from .geometry import Camera, Pose, view_synthesis
    from .losses import MultiViewPhotometricDecayLoss, SupervisedDepthPoseLoss
    from .networks.cost import FeatureMetricCost
    from .utils.depth import post_process_inv_depth, compute_depth_metrics

    patched = []
    table = [
        ("dro_sfm.geometry.pose", dict(Pose=Pose)),
        ("dro_sfm.geometry.camera", dict(Camera=Camera, Pose=Pose)),
        ("dro_sfm.geometry.camera_utils", dict(view_synthesis=view_synthesis)),
        ("dro_sfm.losses.multiview_photometric_loss_mf",
         dict(MultiViewPhotometricDecayLoss=MultiViewPhotometricDecayLoss, Camera=Camera, view_synthesis=view_synthesis)),
        ("dro_sfm.losses.supervised_loss", dict(SupervisedDepthPoseLoss=SupervisedDepthPoseLoss, Camera=Camera, Pose=Pose)),
        ("dro_sfm.models.SelfSupModelMF", dict(MultiViewPhotometricDecayLoss=MultiViewPhotometricDecayLoss)),
        ("dro_sfm.models.SupModelMF", dict(SupervisedDepthPoseLoss=SupervisedDepthPoseLoss)),
        ("dro_sfm.models.SemiSupModelMF", dict(SupervisedDepthPoseLoss=SupervisedDepthPoseLoss,
                                              MultiViewPhotometricDecayLoss=MultiViewPhotometricDecayLoss)),
        ("dro_sfm.models.SfmModelMF", dict(Pose=Pose)),
        ("dro_sfm.networks.depth_pose.DepthPoseNet", dict(Camera=Camera, Pose=Pose)),
        # evaluation path (model_wrapper.py:355-399): CUDA tensors only -- the reference's own functions stay in place for
        # CPU evaluation because these wrappers refuse non-CUDA tensors
        ("dro_sfm.utils.depth", dict(post_process_inv_depth=post_process_inv_depth, compute_depth_metrics=compute_depth_metrics)),
        ("dro_sfm.models.model_wrapper", dict(post_process_inv_depth=post_process_inv_depth,
                                              compute_depth_metrics=compute_depth_metrics)),
    ]
    for module_name, symbols in table:
        if _rebind(module_name, **symbols):
            patched.append(module_name)
    net_mod = sys.modules.get("dro_sfm.networks.depth_pose.DepthPoseNet")
    if net_mod is not None and hasattr(net_mod, "DepthPoseNet"):
        net_mod.DepthPoseNet.get_cost_each = FeatureMetricCost.get_cost_each
        net_mod.DepthPoseNet.depth_cost_calc = FeatureMetricCost.depth_cost_calc
        net_mod.DepthPoseNet.upsample_depth = FeatureMetricCost.upsample_depth
        if lockstep:
            from .networks import lockstep as _lockstep
            net_mod.DepthPoseNet.forward = _lockstep.forward
    return patched
