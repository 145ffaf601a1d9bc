"""The CPU oracle against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import c_oracle
from conftest import t, assert_close

TAGS = ("kitti", "flip", "scan8")


def same(a, b):
    """Exact float equality (+0 == -0), NaN-safe."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("tag", TAGS)
def test_c_oracle_coords_bit_exact(golden, tag):
    g = golden("coords")
    K, s = g[f"{tag}_K"].astype(np.float32), float(g[f"{tag}_scale"])
    depth, T = g[f"{tag}_depth"], g[f"{tag}_T"]
    Ks = np.empty_like(K)
    Ki = np.empty_like(K)
    lib = c_oracle.lib()
    for b in range(len(K)):
        lib.drosfm_oracle_scale_K(c_oracle._p(K[b]), c_oracle.ctypes.c_float(s), c_oracle.ctypes.c_float(s),
                                  c_oracle._p(Ks[b]))
        lib.drosfm_oracle_K_inverse(c_oracle._p(Ks[b]), c_oracle._p(Ki[b]))
    assert same(Ks, g[f"{tag}_Ks"]) and same(Ki, g[f"{tag}_Kinv"])
    assert same(c_oracle.reconstruct(depth, Ks, None), g[f"{tag}_Pc"])
    eye = np.tile(np.eye(4, dtype=np.float32), (len(K), 1, 1))
    Pw = c_oracle.reconstruct(depth, Ks, eye)
    assert same(Pw, g[f"{tag}_Pw"])
    assert same(c_oracle.project(Pw, Ks, T, True), g[f"{tag}_uv"])
    assert same(c_oracle.project(Pw, Ks, T, False), g[f"{tag}_uv_raw"])
    assert same(c_oracle.project(Pw, Ks, None, True), g[f"{tag}_uv_c"])
    assert same(c_oracle.warp_coords(depth, K, K, T, s, s, True), g[f"{tag}_uv"])
    assert same(c_oracle.warp_coords(depth, K, K, T, s, s, False), g[f"{tag}_uv_raw"])
    Twc2 = oracle.invert_T(t(g[f"{tag}_T2"])).numpy()
    assert same(c_oracle.reconstruct(depth, Ks, Twc2), g[f"{tag}_Pw2"])


@pytest.mark.parametrize("tag", TAGS)
def test_torch_oracle_coords_bit_exact(golden, tag):
    g = golden("coords")
    K, s = t(g[f"{tag}_K"]).float(), float(g[f"{tag}_scale"])
    depth, T = t(g[f"{tag}_depth"]), t(g[f"{tag}_T"])
    Ks = oracle.scale_K(K, s)
    assert same(Ks, g[f"{tag}_Ks"]) and same(oracle.K_inverse(Ks), g[f"{tag}_Kinv"])
    Pw = oracle.reconstruct(depth, Ks, None, "w")
    assert same(Pw, g[f"{tag}_Pw"])
    assert same(oracle.reconstruct(depth, Ks, None, "c"), g[f"{tag}_Pc"])
    assert same(oracle.reconstruct(depth, Ks, t(g[f"{tag}_T2"]), "w"), g[f"{tag}_Pw2"])
    assert same(oracle.project(Pw, Ks, T, "w", True), g[f"{tag}_uv"])
    assert same(oracle.project(Pw, Ks, T, "w", False), g[f"{tag}_uv_raw"])
    assert same(oracle.project(Pw, Ks, T, "c", True), g[f"{tag}_uv_c"])
    assert same(oracle.warp_coords(depth, t(g[f"{tag}_K"]), t(g[f"{tag}_K"]), T, s), g[f"{tag}_uv"])


def test_inv2depth_mask_bit_exact(golden):
    g = golden("view_synthesis")
    inv = g["inv_depth"]
    assert (inv <= 0).any()
    assert same(c_oracle.inv2depth(inv), oracle.inv2depth(t(inv)).numpy())


def test_supervised_coords_and_mask_bit_exact(golden):
    g = golden("supervised")
    depth = c_oracle.inv2depth(g["gt_inv_depth"])
    K = g["K"].astype(np.float32)
    uv, mask = c_oracle.warp_coords(depth, K, K, g["gt_T0"], 1.0, 1.0, True, want_mask=True)
    assert same(uv, g["coords_gt0"])
    assert np.array_equal(mask, g["mask_gt0"])
    c, m = oracle.reproj_coords(t(g["gt_T0"]), t(g["K"]), t(g["K"]), oracle.inv2depth(t(g["gt_inv_depth"])))
    assert same(c, g["coords_gt0"]) and np.array_equal(m.numpy(), g["mask_gt0"])


def _grads(out, gout, *xs):
    return torch.autograd.grad(out, xs, gout, allow_unused=True)


@pytest.mark.parametrize("dt,tag", [(torch.float32, "f32"), (torch.float64, "f64")])
@pytest.mark.parametrize("pad", ["zeros", "border"])
def test_view_synthesis(golden, pad, dt, tag):
    g = golden("view_synthesis")
    src, inv, T = (t(g[k], dt).requires_grad_(True) for k in ("src", "inv_depth", "T"))
    K = t(g["K"]).float().to(dt)
    y = oracle.view_synthesis(src, oracle.inv2depth(inv), K, K, T, 1.0, pad)
    gs, gi, gT = _grads(y, t(g["gout"], dt), src, inv, T)
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float32 else dict(rtol=1e-12, atol=1e-13)
    assert_close(y.detach(), g[f"{pad}_{tag}_out"], what="out", **tol)
    assert_close(gs, g[f"{pad}_{tag}_g_src"], what="g_src", **tol)
    assert_close(gi, g[f"{pad}_{tag}_g_inv"], what="g_inv", **(tol if dt == torch.float64 else dict(rtol=1e-5, atol=1e-6)))
    assert_close(gT, g[f"{pad}_{tag}_g_T"], what="g_T", **(tol if dt == torch.float64 else dict(rtol=1e-5, atol=1e-5)))


@pytest.mark.parametrize("dt,tag", [(torch.float32, "f32"), (torch.float64, "f64")])
def test_feat_cost(golden, dt, tag):
    g = golden("feat_cost")
    K = t(g["K"]).float().to(dt)
    tol = dict(rtol=1e-6, atol=1e-6) if dt == torch.float32 else dict(rtol=1e-12, atol=1e-12)
    gtol = dict(rtol=1e-5, atol=1e-4) if dt == torch.float32 else tol
    fmap, fref, pose = (t(g[k], dt).requires_grad_(True) for k in ("fmap", "fref0", "pose0"))
    c = oracle.feat_cost_each(pose, fmap, fref, oracle.inv2depth(t(g["inv_depth"], dt)), K, K, 0.125)
    gp, gf, gr = _grads(c, t(g["gout"], dt), pose, fmap, fref)
    assert_close(c.detach(), g[f"each_{tag}_cost"], what="cost", **tol)
    assert_close(gf, g[f"each_{tag}_g_fmap"], what="g_fmap", **tol)
    assert_close(gr, g[f"each_{tag}_g_fref"], what="g_fref", **tol)
    assert_close(gp, g[f"each_{tag}_g_pose"], what="g_pose", **gtol)
    fmap, f0, f1, inv = (t(g[k], dt).requires_grad_(True) for k in ("fmap", "fref0", "fref1", "inv_depth"))
    c = oracle.depth_cost(inv, fmap, [f0, f1], [t(g["pose0"], dt), t(g["pose1"], dt)], K, K, 0.125)
    gi, gf, g0, g1 = _grads(c, t(g["gout"], dt), inv, fmap, f0, f1)
    assert_close(c.detach(), g[f"depth_{tag}_cost"], what="cost", **tol)
    assert_close(gi, g[f"depth_{tag}_g_inv"], what="g_inv", **gtol)
    assert_close(gf, g[f"depth_{tag}_g_fmap"], what="g_fmap", **tol)
    assert_close(g0, g[f"depth_{tag}_g_fref0"], what="g_fref0", **tol)
    assert_close(g1, g[f"depth_{tag}_g_fref1"], what="g_fref1", **tol)


PHOTO_VARIANTS = {
    "default": dict(automask=True, reduce_op="min", clip=0.0, smooth_w=0.001, padding_mode="zeros", ssim_w=0.85),
    "nomask_mean_border": dict(automask=False, reduce_op="mean", clip=0.0, smooth_w=0.1, padding_mode="border", ssim_w=0.85),
    "min_nomask_clip": dict(automask=False, reduce_op="min", clip=0.5, smooth_w=0.0, padding_mode="zeros", ssim_w=0.85),
    "l1only": dict(automask=True, reduce_op="min", clip=0.0, smooth_w=0.0, padding_mode="zeros", ssim_w=0.0),
}


@pytest.mark.parametrize("dt,tag", [(torch.float32, "f32"), (torch.float64, "f64")])
@pytest.mark.parametrize("name", list(PHOTO_VARIANTS))
def test_photometric_loss(golden, name, dt, tag):
    g = golden("photometric")
    V, n = 2, 3
    K = t(g["K"]).float().to(dt)
    image = t(g["image"], dt)
    context = [t(g[f"context{v}"], dt) for v in range(V)]
    invs = [t(g[f"inv{i}"], dt).requires_grad_(True) for i in range(n)]
    Ts = [[t(g[f"T{v}_{i}"], dt).requires_grad_(True) for i in range(n)] for v in range(V)]
    loss, metrics = oracle.multiview_photometric_decay_loss(image, context, invs, K, K, Ts, **PHOTO_VARIANTS[name])
    flat = invs + [x for tv in Ts for x in tv]
    grads = torch.autograd.grad(loss.sum(), flat, allow_unused=True)
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float32 else dict(rtol=1e-12, atol=1e-13)
    assert_close(loss.detach(), g[f"{name}_{tag}_loss"], what="loss", **tol)
    for k, val in metrics.items():
        assert_close(val, g[f"{name}_{tag}_{k}"], what=k, **tol)
    gtol = dict(rtol=1e-5, atol=1e-7) if dt == torch.float32 else tol
    for i in range(n):
        assert_close(grads[i], g[f"{name}_{tag}_g_inv{i}"], what=f"g_inv{i}", **gtol)
    k = n
    for v in range(V):
        for i in range(n):
            assert_close(grads[k], g[f"{name}_{tag}_g_T{v}_{i}"], what=f"g_T{v}_{i}", **gtol)
            k += 1


def test_ssim_and_photometric_map(golden):
    g = golden("photometric")
    assert_close(oracle.ssim(t(g["context0"]), t(g["image"])), g["ssim"], rtol=1e-6, atol=1e-7)
    assert_close(oracle.photometric_map(t(g["context0"]), t(g["image"])), g["photometric_map_unwarped"],
                 rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("dt,tag", [(torch.float32, "f32"), (torch.float64, "f64")])
def test_supervised_losses(golden, dt, tag):
    g = golden("supervised")
    V, n = 2, 3
    K = t(g["K"]).float().to(dt)
    gt_inv = t(g["gt_inv_depth"], dt)
    invs = [t(g[f"inv{i}"], dt).requires_grad_(True) for i in range(n)]
    pred = [[t(g[f"pred_T{v}_{i}"], dt).requires_grad_(True) for i in range(n)] for v in range(V)]
    gtT = [t(g[f"gt_T{v}"], dt) for v in range(V)]
    lp = oracle.reproj_pose_loss(pred, gtT, oracle.inv2depth(gt_inv), K, K, 0.2, 80.0)
    ld = oracle.supervised_depth_loss(invs, gt_inv, 0.2, 80.0)
    flat = invs + [x for tv in pred for x in tv]
    grads = torch.autograd.grad(lp + ld, flat)
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float32 else dict(rtol=1e-12, atol=1e-13)
    assert_close(lp.detach(), g[f"{tag}_pose_loss"], what="pose_loss", **tol)
    assert_close(ld.detach(), g[f"{tag}_depth_loss"], what="depth_loss", **tol)
    gtol = dict(rtol=1e-5, atol=1e-7) if dt == torch.float32 else tol
    for i in range(n):
        assert_close(grads[i], g[f"{tag}_g_inv{i}"], what=f"g_inv{i}", **gtol)
    k = n
    for v in range(V):
        for i in range(n):
            assert_close(grads[k], g[f"{tag}_g_T{v}_{i}"], what=f"g_T{v}_{i}", **gtol)
            k += 1


@pytest.mark.parametrize("dt,tag", [(torch.float32, "f32"), (torch.float64, "f64")])
def test_upsample_depth(golden, dt, tag):
    g = golden("upsample")
    d, m = t(g["depth"], dt).requires_grad_(True), t(g["mask"], dt).requires_grad_(True)
    y = oracle.upsample_depth(d, m, 8)
    gd, gm = torch.autograd.grad(y, (d, m), t(g["gout"], dt))
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float32 else dict(rtol=1e-12, atol=1e-13)
    assert_close(y.detach(), g[f"{tag}_out"], what="out", **tol)
    assert_close(gd, g[f"{tag}_g_depth"], what="g_depth", **(tol if dt == torch.float64 else dict(rtol=1e-5, atol=1e-6)))
    assert_close(gm, g[f"{tag}_g_mask"], what="g_mask", **tol)


@pytest.mark.parametrize("scale", [0.01, 0.05, 2.0])
def test_euler_restatement_matches_torch_cpu(scale):
    """oracle/coords_oracle.c:drosfm_oracle_euler_from_trig (non-FMA accumulation) reproduces euler2mat as torch
    evaluates it on the CPU bit for bit -- signed zeros included -- when fed torch's own sin / cos."""
    import numpy as np
    torch.manual_seed(11)
    ang = (torch.randn(4096, 3) * scale).float()
    ang[:64] = -ang[:64].abs()
    ang[64:72] = 0.0
    ang[72:80] = -0.0
    R = oracle.euler_to_R(ang).numpy()
    x, y, z = ang[:, 0], ang[:, 1], ang[:, 2]
    trig = torch.stack([torch.sin(x), torch.cos(x), torch.sin(y), torch.cos(y), torch.sin(z), torch.cos(z)], 1).numpy()
    Rc = c_oracle.euler_from_trig(trig, z.numpy(), fma=False)
    assert np.array_equal(R.view(np.uint32), Rc.view(np.uint32))
    # the FMA accumulation (what a CUDA bmm evaluates) differs from it by at most one ulp of 1.0
    Rf = c_oracle.euler_from_trig(trig, z.numpy(), fma=True)
    assert np.abs(Rf.astype(np.float64) - R).max() <= 1.2e-7
