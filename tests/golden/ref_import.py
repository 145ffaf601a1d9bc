"""Import the UNMODIFIED reference from /root/reference on CPU (build container only).

The reference cannot travel to the GPU box, so this module is used solely by
``make_golden.py`` (fixture generation) and by ``tests/test_oracle_vs_reference.py``
(skipped when /root/reference is absent).  Harness-side shims, SURVEY.md section 8c:
  * stub modules for yacs / matplotlib.cm / numpy.lib.type_check / termcolor,
    which the reference imports but the hot path never uses;
  * ``Tensor.get_device`` returns the device object on CPU, because
    ``warp_ref_image`` (multiview_photometric_loss_mf.py:156) feeds it to ``.to()``.
"""
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("DROSFM_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "dro_sfm"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    import numpy as np
    import torch

    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    logging.disable(logging.WARNING)  # every hot-path ctor logs at WARNING

    class _CfgNode(dict):
        pass

    yacs = _stub("yacs")
    yacs.config = _stub("yacs.config", CfgNode=_CfgNode)
    try:
        import matplotlib.cm  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.cm = _stub("matplotlib.cm", get_cmap=lambda *a, **k: None)
    try:
        from numpy.lib.type_check import imag  # noqa: F401
    except Exception:
        _stub("numpy.lib.type_check", imag=np.imag)
    try:
        import termcolor  # noqa: F401
    except Exception:
        _stub("termcolor", colored=lambda s, *a, **k: s)

    if not getattr(torch.Tensor.get_device, "_drosfm_patched", False):
        orig = torch.Tensor.get_device

        def get_device(t):
            return orig(t) if t.is_cuda else t.device
        get_device._drosfm_patched = True
        torch.Tensor.get_device = get_device

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    from dro_sfm.geometry.camera import Camera
    from dro_sfm.geometry.pose import Pose
    from dro_sfm.geometry.camera_utils import view_synthesis, scale_intrinsics
    from dro_sfm.utils.depth import inv2depth, calc_smoothness
    from dro_sfm.losses.multiview_photometric_loss_mf import MultiViewPhotometricDecayLoss, SSIM
    from dro_sfm.losses.supervised_loss import SupervisedDepthPoseLoss
    from dro_sfm.networks.depth_pose.DepthPoseNet import DepthPoseNet

    net = object.__new__(DepthPoseNet)  # get_cost_each / depth_cost_calc use no module state
    return types.SimpleNamespace(
        Camera=Camera, Pose=Pose, view_synthesis=view_synthesis, scale_intrinsics=scale_intrinsics,
        inv2depth=inv2depth, calc_smoothness=calc_smoothness, SSIM=SSIM,
        MultiViewPhotometricDecayLoss=MultiViewPhotometricDecayLoss,
        SupervisedDepthPoseLoss=SupervisedDepthPoseLoss,
        get_cost_each=lambda *a, **k: DepthPoseNet.get_cost_each(net, *a, **k),
        depth_cost_calc=lambda *a, **k: DepthPoseNet.depth_cost_calc(net, *a, **k),
        upsample_depth=lambda *a, **k: DepthPoseNet.upsample_depth(net, *a, **k),
    )
