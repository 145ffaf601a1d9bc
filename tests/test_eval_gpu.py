"""Evaluation path on the GPU: post_process_inv_depth and compute_depth_metrics through the C ABI
against the reference's golden outputs and against the op-for-op torch restatement on the same device."""
import types

import numpy as np
import pytest
import torch

from oracle import eval_oracle as eo
from dro_sfm_b200 import ops
from dro_sfm_b200.utils import depth as depth_utils
from dro_sfm_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CROPS = {0: "", 1: "garg", 2: "eigen_nyu"}
RTOL = 1e-5          # north_star: fp32 outputs within 1e-5 relative


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("tag", ["pp_a", "pp_b"])
@pytest.mark.parametrize("method", ["mean", "max", "min"])
def test_post_process_golden(golden, tag, method):
    G = golden("eval")
    out = depth_utils.post_process_inv_depth(dev(G[tag + "_inv"]), dev(G[tag + "_inv_flipped"]), method=method)
    # the host golden carries torch's vectorised CPU linspace (one ulp off the scalar rule in places): 1e-6, not bits
    np.testing.assert_allclose(out.cpu().numpy(), G["%s_%s" % (tag, method)], rtol=1e-6, atol=0)


@pytest.mark.parametrize("shape", [(2, 1, 6, 41), (1, 1, 5, 64), (4, 1, 320, 960), (2, 1, 192, 641), (1, 1, 3, 1), (1, 2, 4, 2), (1, 1, 7, 1242)])
@pytest.mark.parametrize("method", ["mean", "max", "min"])
def test_post_process_bit_exact_vs_torch_cuda(shape, method):
    g = syn.gen(11)
    B, C, H, W = shape
    a = syn.inv_depth(g, B * C, H, W, 0.1, 100.0).view(B, C, H, W).to(DEV)
    b = syn.inv_depth(g, B * C, H, W, 0.1, 100.0).view(B, C, H, W).to(DEV)
    want = eo.post_process_inv_depth_torch(a, b, method)
    got = ops.post_process_inv_depth(a, b, method)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_post_process_errors():
    a = torch.ones(1, 1, 4, 4, device=DEV)
    with pytest.raises(ValueError):
        ops.post_process_inv_depth(a, a, "median")
    with pytest.raises(ValueError):
        ops.post_process_inv_depth(a, a[..., :3].contiguous(), "mean")
    with pytest.raises((ValueError, RuntimeError)):
        ops.post_process_inv_depth(a.cpu(), a.cpu(), "mean")
    assert ops.post_process_inv_depth(a[:0], a[:0]).shape == (0, 1, 4, 4)


@pytest.mark.parametrize("tag", ["m_garg", "m_none", "m_nyu"])
@pytest.mark.parametrize("scale", [1, 0])
def test_depth_metrics_golden(golden, tag, scale):
    G = golden("eval")
    lo, hi, crop = G[tag + "_cfg"]
    cfg = types.SimpleNamespace(crop=CROPS[int(crop)], min_depth=float(lo), max_depth=float(hi))
    out = depth_utils.compute_depth_metrics(cfg, dev(G[tag + "_gt"]), dev(G[tag + "_pred"]), use_gt_scale=bool(scale))
    assert out.shape == (9,) and out.dtype == torch.float32 and out.is_cuda
    np.testing.assert_allclose(out.cpu().numpy(), G["%s_scale%d" % (tag, scale)], rtol=RTOL, atol=1e-7)
    # the workspace cleans itself: a second call on the same stream gives the same bits
    again = depth_utils.compute_depth_metrics(cfg, dev(G[tag + "_gt"]), dev(G[tag + "_pred"]), use_gt_scale=bool(scale))
    assert torch.equal(out, again)


@pytest.mark.parametrize("crop,B,H,W,h,w", [("garg", 4, 375, 1242, 320, 960), ("", 12, 240, 320, 240, 320), ("eigen_nyu", 3, 480, 640, 240, 320)])
@pytest.mark.parametrize("scale", [True, False])
def test_depth_metrics_full_size_vs_torch_cuda(crop, B, H, W, h, w, scale):
    """BASELINE-size evaluation batches (KITTI ground truth 375x1242 against a 320x960 prediction; ScanNet 240x320)."""
    g = syn.gen(29)
    lo, hi = (0.1, 80.0) if crop == "garg" else (0.1, 10.0)
    gt = (lo * 0.5 + torch.rand(B, 1, H, W, generator=g) * (hi * 1.1 - lo * 0.5)) * (torch.rand(B, 1, H, W, generator=g) < 0.3)
    pred = (lo + torch.rand(B, 1, h, w, generator=g) * (hi - lo) * 0.6)
    pred = torch.nn.functional.avg_pool2d(pred, 5, 1, 2)          # smooth, like a network output
    gt, pred = gt.to(DEV), pred.to(DEV)
    want = eo.compute_depth_metrics_torch(crop, lo, hi, gt, pred, scale)
    got = ops.depth_metrics(gt, pred, lo, hi, crop, scale)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=RTOL, atol=1e-7)


def test_median_is_exact_lower_median():
    """With gt = k * pred (k a power of two, so gt / pred is exact) on an even number of valid pixels drawn from two
    distinct ratios, torch.median picks the LOWER middle value; abs_rel then is exactly that of the lower ratio."""
    H, W = 8, 16
    pred = torch.full((1, 1, H, W), 2.0)
    gt = pred * 2.0
    gt[0, 0, :4] = pred[0, 0, :4] * 4.0         # half the pixels ratio 4, half ratio 2 -> lower median 2
    got = ops.depth_metrics(gt.to(DEV), pred.to(DEV), 0.1, 100.0, "", True).cpu().numpy()
    want = eo.compute_depth_metrics("", 0.1, 100.0, gt.numpy(), pred.numpy(), True)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-8)
    # ratio-2 half: pred * 2 = gt exactly; ratio-4 half: |8 - 4| / 8 = 0.5 -> abs_rel = 0.25
    assert abs(got[0] - 0.25) < 1e-7


def test_depth_metrics_scale_invariance_and_empty():
    g = syn.gen(31)
    gt = (0.5 + torch.rand(3, 1, 60, 90, generator=g) * 50).to(DEV)
    pred = (0.5 + torch.rand(3, 1, 60, 90, generator=g) * 50).to(DEV)
    a = ops.depth_metrics(gt, pred, 0.1, 80.0, "", True)
    b = ops.depth_metrics(gt, pred * 4.0, 0.1, 80.0, "", True)          # exact power-of-two rescale: median scaling undoes it
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-6)
    # no valid pixel anywhere: all zeros, as the reference's untouched accumulators
    z = ops.depth_metrics(torch.zeros_like(gt), pred, 0.1, 80.0, "garg", True)
    assert torch.count_nonzero(z) == 0
    with pytest.raises(ValueError):
        ops.depth_metrics(gt[:, 0], pred, 0.1, 80.0)
