"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the exact inputs and the reference's outputs (float32, CPU,
torch 2.11.0) -- and, for differentiable paths, the reference's autograd gradients in
float32 plus a float64 re-run of the same reference code (intrinsics kept in float64
by feeding float64 tensors; the reference's ``K.float()`` casts are honoured, i.e. the
float64 run uses the float32-rounded intrinsics, exactly what the fp32 path sees).
Sizes are small so the .npz files stay a few hundred KB in total.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_import  # noqa: E402
from dro_sfm_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)


def np_(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (np_(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


class _F64K:
    """Context manager for the float64 re-runs of the reference: make ``K.float()`` a no-op (the
    intrinsics fed in are already rounded through float32, so the value is the same) and let
    ``Pose.identity`` (pose.py:31-33, default dtype float32) follow the default dtype, which is
    switched to float64 for the duration.  No effect on a float32 run."""
    def __init__(self, dt=torch.float64):
        self.dt = dt

    def __enter__(self):
        self.orig = torch.Tensor.float
        self.orig_default = torch.get_default_dtype()
        if self.dt == torch.float64:
            from dro_sfm.geometry.pose import Pose
            torch.Tensor.float = lambda t, *a, **k: t
            torch.set_default_dtype(torch.float64)
            self.orig_identity = Pose.identity
            Pose.identity = classmethod(lambda cls, N=1, device=None, dtype=torch.float64:
                                        cls(torch.eye(4, device=device, dtype=dtype).repeat([N, 1, 1])))
        return self

    def __exit__(self, *exc):
        if self.dt == torch.float64:
            from dro_sfm.geometry.pose import Pose
            torch.Tensor.float = self.orig
            torch.set_default_dtype(self.orig_default)
            Pose.identity = self.orig_identity


def leaf(x, dt):
    """A fresh leaf copy of x in dtype dt that records gradients."""
    return x.detach().clone().to(dt).requires_grad_(True)


def rand_T(ref, g, B, dataset):
    return ref.Pose.from_vec(syn.pose_vec(g, B, dataset), "euler").mat.contiguous()


def case_coords(ref):
    """Camera.reconstruct / Camera.project (camera.py:111-194), incl. depth==0, Z<1e-5, flipped fx."""
    g = syn.gen(101)
    B, H, W = 2, 24, 40
    out = {}
    for tag, dataset, flip, scale in (("kitti", "kitti", False, 1.0), ("flip", "kitti", True, 1.0),
                                      ("scan8", "scannet", False, 0.125)):
        K = syn.intrinsics(dataset, B, H * int(1 / scale), W * int(1 / scale), flip=flip)
        depth = ref.inv2depth(syn.inv_depth(g, B, H, W, 0.2, 80.0, frac_nonpos=0.1))
        T = rand_T(ref, g, B, dataset)
        if tag == "flip":
            T[1, 2, 3] = -30.0     # push sample 1 behind the source camera: exercises Z.clamp(min=1e-5)
        cam = ref.Camera(K=K.float()).scaled(scale)
        rcam = ref.Camera(K=K.float(), Tcw=ref.Pose(T.clone())).scaled(scale)
        Pw = cam.reconstruct(depth, frame="w")
        Pc = cam.reconstruct(depth, frame="c")
        uv = rcam.project(Pw, frame="w", normalize=True)
        uv_raw = rcam.project(Pw, frame="w", normalize=False)
        uv_c = rcam.project(Pw, frame="c", normalize=True)
        # a target camera that is NOT at the identity
        T2 = rand_T(ref, g, B, dataset)
        Pw2 = ref.Camera(K=K.float(), Tcw=ref.Pose(T2.clone())).scaled(scale).reconstruct(depth, frame="w")
        out.update({f"{tag}_K": K, f"{tag}_scale": scale, f"{tag}_depth": depth, f"{tag}_T": T,
                    f"{tag}_Pw": Pw, f"{tag}_Pc": Pc, f"{tag}_uv": uv, f"{tag}_uv_raw": uv_raw,
                    f"{tag}_uv_c": uv_c, f"{tag}_T2": T2, f"{tag}_Pw2": Pw2,
                    f"{tag}_Ks": cam.K, f"{tag}_Kinv": cam.Kinv})
    save("coords", **out)


def case_view_synthesis(ref):
    """view_synthesis (camera_utils.py:23-56), zeros and border, with input gradients."""
    g = syn.gen(202)
    B, H, W = 2, 24, 40
    K = syn.intrinsics("kitti", B, H, W)
    src = syn.images(g, B, H, W)
    inv = syn.inv_depth(g, B, H, W, 0.5, 80.0, frac_nonpos=0.05)
    T = rand_T(ref, g, B, "kitti")
    T[:, 0, 3] += 0.8       # sideways shift so that a band of pixels leaves the image
    gout = torch.randn(B, 3, H, W, generator=g)
    out = {"K": K, "src": src, "inv_depth": inv, "T": T, "gout": gout}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        for pad in ("zeros", "border"):
            s = leaf(src, dt)
            d = leaf(inv, dt)
            Tm = leaf(T, dt)
            Kd = K.float().to(dt)
            with _F64K(dt):
                cam = ref.Camera(K=Kd)
                rcam = ref.Camera(K=Kd, Tcw=ref.Pose(Tm))
                y = ref.view_synthesis(s, ref.inv2depth(d), rcam, cam, padding_mode=pad)
            y.backward(gout.to(dt))
            out.update({f"{pad}_{tag}_out": y, f"{pad}_{tag}_g_src": s.grad, f"{pad}_{tag}_g_inv": d.grad,
                        f"{pad}_{tag}_g_T": Tm.grad})
    save("view_synthesis", **out)


def case_feat_cost(ref):
    """DepthPoseNet.get_cost_each / depth_cost_calc (DepthPoseNet.py:76-105) with gradients."""
    g = syn.gen(303)
    B, C, h, w, V = 2, 16, 12, 20, 2
    K = syn.intrinsics("kitti", B, h * 8, w * 8)
    fmap = syn.features(g, B, C, h, w)
    frefs = [syn.features(g, B, C, h, w) for _ in range(V)]
    inv = syn.inv_depth(g, B, h, w, 0.5, 80.0, frac_nonpos=0.05)
    poses = [syn.pose_vec(g, B, "kitti", 1.0 if v == 0 else -1.0) for v in range(V)]
    gout = torch.randn(B, C, h, w, generator=g)
    out = {"K": K, "fmap": fmap, "inv_depth": inv, "gout": gout}
    for v in range(V):
        out[f"fref{v}"] = frefs[v]
        out[f"pose{v}"] = poses[v]
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        Kd = K.float().to(dt)
        # pose cost: grads to pose, fmap, fmap_ref; depth arrives pre-inverted & detached
        f = leaf(fmap, dt)
        fr = leaf(frefs[0], dt)
        p = leaf(poses[0], dt)
        depth = ref.inv2depth(inv.to(dt))
        with _F64K(dt):
            c = ref.get_cost_each(p, f, fr, depth, Kd, Kd, 0.125)
        c.backward(gout.to(dt))
        out.update({f"each_{tag}_cost": c, f"each_{tag}_g_pose": p.grad, f"each_{tag}_g_fmap": f.grad,
                    f"each_{tag}_g_fref": fr.grad})
        # depth cost: grads to inv_depth, fmap, all fmaps_ref
        f = leaf(fmap, dt)
        frs = [leaf(x, dt) for x in frefs]
        d = leaf(inv, dt)
        with _F64K(dt):
            c = ref.depth_cost_calc(d, f, frs, [q.to(dt) for q in poses], Kd, Kd, 0.125)
        c.backward(gout.to(dt))
        out.update({f"depth_{tag}_cost": c, f"depth_{tag}_g_inv": d.grad, f"depth_{tag}_g_fmap": f.grad})
        for v in range(V):
            out[f"depth_{tag}_g_fref{v}"] = frs[v].grad
    save("feat_cost", **out)


def _photo_inputs(g, B, H, W, V, n, dataset, ref):
    K = syn.intrinsics(dataset, B, H, W)
    image = syn.images(g, B, H, W)
    # context views = small shifts of the target, so that warped and unwarped errors compete (automask)
    context = [torch.roll(image, shifts=(v + 1) * (1 if v % 2 == 0 else -1), dims=3) * 0.9
               + 0.1 * syn.images(g, B, H, W) for v in range(V)]
    invs = [syn.inv_depth(g, B, H, W, 0.5, 80.0, frac_nonpos=0.02) for _ in range(n)]
    Ts = [[rand_T(ref, g, B, dataset) for _ in range(n)] for _ in range(V)]
    for v in range(V):
        for i in range(n):
            Ts[v][i][:, :3, 3] *= 0.2
    return K, image, context, invs, Ts


def case_photometric(ref):
    """SSIM, calc_photometric_loss and the full MultiViewPhotometricDecayLoss.forward
    (multiview_photometric_loss_mf.py:15-54,194-361) for several option sets, with gradients."""
    g = syn.gen(404)
    B, H, W, V, n = 2, 24, 40, 2, 3
    K, image, context, invs, Ts = _photo_inputs(g, B, H, W, V, n, "kitti", ref)
    out = {"K": K, "image": image}
    for v in range(V):
        out[f"context{v}"] = context[v]
        for i in range(n):
            out[f"T{v}_{i}"] = Ts[v][i]
    for i in range(n):
        out[f"inv{i}"] = invs[i]
    out["ssim"] = ref.SSIM(context[0], image)
    variants = {
        "default": dict(automask_loss=True, photometric_reduce_op="min", clip_loss=0.0, smooth_loss_weight=0.001,
                        padding_mode="zeros", ssim_loss_weight=0.85),
        "nomask_mean_border": dict(automask_loss=False, photometric_reduce_op="mean", clip_loss=0.0,
                                   smooth_loss_weight=0.1, padding_mode="border", ssim_loss_weight=0.85),
        "min_nomask_clip": dict(automask_loss=False, photometric_reduce_op="min", clip_loss=0.5,
                                smooth_loss_weight=0.0, padding_mode="zeros", ssim_loss_weight=0.85),
        "l1only": dict(automask_loss=True, photometric_reduce_op="min", clip_loss=0.0, smooth_loss_weight=0.0,
                       padding_mode="zeros", ssim_loss_weight=0.0),
    }
    for name, kw in variants.items():
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            loss_mod = ref.MultiViewPhotometricDecayLoss(**kw)
            d = [leaf(x, dt) for x in invs]
            Tm = [[leaf(t, dt) for t in tv] for tv in Ts]
            Kd = K.float().to(dt)
            with _F64K(dt):
                res = loss_mod(image.to(dt), [c.to(dt) for c in context], d, Kd, Kd,
                               [[ref.Pose(t) for t in tv] for tv in Tm])
            res["loss"].sum().backward()
            out[f"{name}_{tag}_loss"] = res["loss"]
            for k, val in res["metrics"].items():
                out[f"{name}_{tag}_{k}"] = val
            for i in range(n):
                out[f"{name}_{tag}_g_inv{i}"] = d[i].grad
                for v in range(V):
                    out[f"{name}_{tag}_g_T{v}_{i}"] = Tm[v][i].grad
            if name == "default" and tag == "f32":
                pm = loss_mod.calc_photometric_loss([context[0]] * n, [image] * n)
                out["photometric_map_unwarped"] = pm[0]
    save("photometric", **out)


def case_supervised(ref):
    """SupervisedDepthPoseLoss.calc_pose_loss / calculate_loss / get_ref_coords (supervised_loss.py:244-325)."""
    g = syn.gen(505)
    B, H, W, V, n = 2, 24, 40, 2, 3
    K = syn.intrinsics("kitti", B, H, W)
    gt_inv = syn.inv_depth(g, B, H, W, 0.2, 80.0) * (torch.rand(B, 1, H, W, generator=g) > 0.3)
    invs = [syn.inv_depth(g, B, H, W, 0.2, 80.0) for _ in range(n)]
    gt_T = [rand_T(ref, g, B, "kitti") for _ in range(V)]
    pred_T = [[gt_T[v] @ rand_T(ref, g, B, "scannet") for _ in range(n)] for v in range(V)]
    out = {"K": K, "gt_inv_depth": gt_inv}
    for i in range(n):
        out[f"inv{i}"] = invs[i]
    for v in range(V):
        out[f"gt_T{v}"] = gt_T[v]
        for i in range(n):
            out[f"pred_T{v}_{i}"] = pred_T[v][i]
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        mod = ref.SupervisedDepthPoseLoss(min_depth=0.2, max_depth=80.0)
        mod.n = n
        Kd = K.float().to(dt)
        P = [[leaf(t, dt) for t in tv] for tv in pred_T]
        d = [leaf(x, dt) for x in invs]
        gt_depth = ref.inv2depth(gt_inv.to(dt))
        with _F64K(dt):
            lp = mod.calc_pose_loss([[ref.Pose(t) for t in tv] for tv in P], [t.to(dt) for t in gt_T],
                                    gt_depth, Kd, Kd)
            ld = mod.calculate_loss(d, [gt_inv.to(dt)] * n)
            coords, mask = mod.get_ref_coords(gt_T[0].to(dt), Kd, Kd, gt_depth, 1, "cpu")
        (lp + ld).backward()
        out[f"{tag}_pose_loss"] = lp
        out[f"{tag}_depth_loss"] = ld
        for v in range(V):
            for i in range(n):
                out[f"{tag}_g_T{v}_{i}"] = P[v][i].grad
        for i in range(n):
            out[f"{tag}_g_inv{i}"] = d[i].grad
        if tag == "f32":
            out["coords_gt0"] = coords
            out["mask_gt0"] = mask
    save("supervised", **out)


def case_upsample(ref):
    """DepthPoseNet.upsample_depth (DepthPoseNet.py:63-74) with gradients."""
    g = syn.gen(606)
    N, H, W = 2, 5, 7
    depth = syn.inv_depth(g, N, H, W, 0.5, 80.0)
    mask = torch.randn(N, 576, H, W, generator=g) * 2.0
    gout = torch.randn(N, 1, 8 * H, 8 * W, generator=g)
    out = {"depth": depth, "mask": mask, "gout": gout}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        d, m = leaf(depth, dt), leaf(mask, dt)
        y = ref.upsample_depth(d, m, ratio=8)
        y.backward(gout.to(dt))
        out.update({f"{tag}_out": y, f"{tag}_g_depth": d.grad, f"{tag}_g_mask": m.grad})
    save("upsample", **out)


def case_eval(ref):
    """Evaluation path (model_wrapper.py:355-399): post_process_inv_depth (utils/depth.py:230-258) and
    compute_depth_metrics (utils/depth.py:261-340), the reference's own functions on the CPU."""
    import types
    g = syn.gen(707)
    out = {}
    # post-processing: odd and even widths (the linspace mask is built from both ends)
    for tag, (B, H, W) in (("pp_a", (2, 6, 41)), ("pp_b", (1, 5, 64))):
        a = syn.inv_depth(g, B, H, W, 0.5, 80.0)
        b = syn.inv_depth(g, B, H, W, 0.5, 80.0)
        out[tag + "_inv"], out[tag + "_inv_flipped"] = a, b
        for method in ("mean", "max", "min"):
            out["%s_%s" % (tag, method)] = ref.post_process_inv_depth(a, b, method=method)
    # metrics: sparse ground truth with out-of-range values, a prediction at another resolution, one sample without a
    # single valid pixel, every crop mode, with and without median scaling
    cases = {"m_garg": ("garg", 0.1, 80.0, (3, 48, 160), (24, 80)), "m_none": ("", 0.5, 10.0, (2, 30, 50), (30, 50)),
             "m_nyu": ("eigen_nyu", 0.1, 10.0, (1, 480, 640), (60, 80))}
    for tag, (crop, lo, hi, (B, H, W), (h, w)) in cases.items():
        gt = lo * 0.5 + torch.rand(B, 1, H, W, generator=g) * (hi * 1.2 - lo * 0.5)
        gt = gt * (torch.rand(B, 1, H, W, generator=g) < (0.4 if H < 400 else 0.08))  # sparse (LiDAR-like)
        pred = lo + torch.rand(B, 1, h, w, generator=g) * (hi - lo) * 0.7
        pred[0, 0, 0, :3] = 0.0                                                        # exercised by clamp(min=1e-6)
        if B > 2:
            gt[1] = 0.0                                                                # sample without valid pixels
        cfg = types.SimpleNamespace(crop=crop, min_depth=lo, max_depth=hi)
        out[tag + "_gt"], out[tag + "_pred"] = gt, pred
        out[tag + "_cfg"] = np.array([lo, hi, {"": 0, "garg": 1, "eigen_nyu": 2}[crop]], np.float64)
        for scale in (True, False):
            out["%s_scale%d" % (tag, int(scale))] = ref.compute_depth_metrics(cfg, gt, pred, use_gt_scale=scale)
    save("eval", **out)


def main():
    ref = ref_import.load()
    if len(sys.argv) > 1 and sys.argv[1] == "upsample":
        return case_upsample(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "eval":
        return case_eval(ref)
    case_coords(ref)
    case_view_synthesis(ref)
    case_feat_cost(ref)
    case_photometric(ref)
    case_supervised(ref)
    case_upsample(ref)
    case_eval(ref)


if __name__ == "__main__":
    main()
