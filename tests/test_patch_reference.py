"""patch.install() against the real reference tree (build container only; skipped on the GPU box)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_install_rebinds_the_hot_path_symbols():
    ref_import.load()
    import dro_sfm_b200.patch as patch
    from dro_sfm_b200.geometry import Camera, Pose, view_synthesis
    from dro_sfm_b200.losses import MultiViewPhotometricDecayLoss, SupervisedDepthPoseLoss
    from dro_sfm_b200.networks.cost import FeatureMetricCost
    patched = patch.install()
    assert "dro_sfm.geometry.camera" in patched and "dro_sfm.losses.multiview_photometric_loss_mf" in patched
    import dro_sfm.geometry.camera as rc
    import dro_sfm.geometry.camera_utils as rcu
    import dro_sfm.losses.multiview_photometric_loss_mf as rl
    import dro_sfm.losses.supervised_loss as rs
    import dro_sfm.networks.depth_pose.DepthPoseNet as rn
    assert rc.Camera is Camera and rc.Pose is Pose
    assert rcu.view_synthesis is view_synthesis
    assert rl.MultiViewPhotometricDecayLoss is MultiViewPhotometricDecayLoss
    assert rs.SupervisedDepthPoseLoss is SupervisedDepthPoseLoss
    assert rn.DepthPoseNet.get_cost_each is FeatureMetricCost.get_cost_each
    assert rn.DepthPoseNet.depth_cost_calc is FeatureMetricCost.depth_cost_calc
    assert rn.DepthPoseNet.upsample_depth is FeatureMetricCost.upsample_depth
    # evaluation path (model_wrapper.py:355-399 imports these two names from dro_sfm.utils.depth)
    import dro_sfm.utils.depth as rd
    from dro_sfm_b200.utils.depth import post_process_inv_depth, compute_depth_metrics
    assert rd.post_process_inv_depth is post_process_inv_depth and rd.compute_depth_metrics is compute_depth_metrics
    # the lock-step schedule is grafted onto the network's forward
    from dro_sfm_b200.networks import lockstep
    assert rn.DepthPoseNet.forward is lockstep.forward


def test_dropin_signatures_match_the_reference():
    """Same parameter names in the same order as the reference's public entry points."""
    import inspect
    ref = ref_import.load()
    from dro_sfm_b200.geometry import Camera, view_synthesis
    from dro_sfm_b200.losses import MultiViewPhotometricDecayLoss, SupervisedDepthPoseLoss
    from dro_sfm_b200.networks.cost import FeatureMetricCost

    def params(f):
        return list(inspect.signature(f).parameters)

    import importlib
    orig_cam = importlib.import_module("dro_sfm.geometry.camera")
    # the reference modules may already be patched by the previous test: compare against the source text instead
    src = open(os.path.join(ref_import.REF_ROOT, "dro_sfm/geometry/camera.py")).read()
    assert "def reconstruct(self, depth, frame='w')" in src and params(Camera.reconstruct) == ["self", "depth", "frame"]
    assert "def project(self, X, frame='w', normalize=True)" in src and params(Camera.project) == ["self", "X", "frame", "normalize"]
    assert params(view_synthesis) == ["ref_image", "depth", "ref_cam", "cam", "mode", "padding_mode"]
    assert params(MultiViewPhotometricDecayLoss.forward) == ["self", "image", "context", "inv_depths", "K", "ref_K", "poses",
                                                             "return_logs", "progress"]
    assert params(SupervisedDepthPoseLoss.forward) == ["self", "image", "context", "inv_depths", "gt_inv_depth",
                                                       "gt_pose_context", "K", "ref_K", "poses", "return_logs", "progress"]
    assert params(FeatureMetricCost.get_cost_each) == ["self", "pose", "fmap", "fmap_ref", "depth", "K", "ref_K", "scale_factor"]
    assert params(FeatureMetricCost.depth_cost_calc) == ["self", "inv_depth", "fmap", "fmaps_ref", "pose_list", "K", "ref_K",
                                                         "scale_factor"]
    assert params(FeatureMetricCost.upsample_depth) == ["self", "depth", "mask", "ratio"]
    from dro_sfm_b200.utils import depth as my_depth
    src = open(os.path.join(ref_import.REF_ROOT, "dro_sfm/utils/depth.py")).read()
    assert "def post_process_inv_depth(inv_depth, inv_depth_flipped, method='mean')" in src
    assert params(my_depth.post_process_inv_depth) == ["inv_depth", "inv_depth_flipped", "method"]
    assert "def compute_depth_metrics(config, gt, pred, use_gt_scale=True)" in src
    assert params(my_depth.compute_depth_metrics) == ["config", "gt", "pred", "use_gt_scale"]
    assert orig_cam is not None and ref is not None
