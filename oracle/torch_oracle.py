"""Functional PyTorch-CPU restatement of the reference hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line it restates (paths relative to the
upstream repo root).  The arithmetic goes through the same ATen CPU operators
the reference uses (``bmm``, ``grid_sample``, ``avg_pool2d``, reflection pad),
so on the same host it reproduces the reference bit for bit; that claim is
checked against fixtures produced by the real reference
(``tests/golden/make_golden.py``).  All functions are dtype-generic: feed
float64 tensors to obtain the high-precision comparator used for the
gradient tolerances.

Conventions: intrinsics ``K`` are ``[B,3,3]``; rigid transforms ``T`` are
``[B,4,4]`` (only the top 3x4 block is used); pose vectors are ``[B,6]`` =
``(tx,ty,tz,rx,ry,rz)`` with euler rotation ``R = Rx @ Ry @ Rz``.
"""
import torch
import torch.nn.functional as F

__all__ = [
    "scale_K", "K_inverse", "invert_T", "euler_to_R", "pose_vec_to_T", "pixel_grid",
    "inv2depth", "reconstruct", "project", "warp_coords", "grid_gather",
    "view_synthesis", "feat_cost_each", "depth_cost", "ssim", "photometric_map",
    "photometric_loss", "smoothness_loss", "multiview_photometric_decay_loss",
    "reproj_coords", "reproj_pose_loss", "supervised_depth_loss", "upsample_depth",
]


# ----------------------------------------------------------------------------------------------
# intrinsics / poses
# ----------------------------------------------------------------------------------------------
def scale_K(K, x_scale, y_scale=None):
    """dro_sfm/geometry/camera.py:83-107 + camera_utils.py:13-19 (returns K itself when scale==1)."""
    if y_scale is None:
        y_scale = x_scale
    if x_scale == 1. and y_scale == 1.:
        return K
    Ks = K.clone()
    Ks[..., 0, 0] *= x_scale
    Ks[..., 1, 1] *= y_scale
    Ks[..., 0, 2] = (Ks[..., 0, 2] + 0.5) * x_scale - 0.5
    Ks[..., 1, 2] = (Ks[..., 1, 2] + 0.5) * y_scale - 0.5
    return Ks


def K_inverse(K):
    """dro_sfm/geometry/camera.py:70-79 -- closed form; every other entry is copied from K."""
    Ki = K.clone()
    fx, fy, cx, cy = K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]
    Ki[:, 0, 0] = 1. / fx
    Ki[:, 1, 1] = 1. / fy
    Ki[:, 0, 2] = -1. * cx / fx
    Ki[:, 1, 2] = -1. * cy / fy
    return Ki


def invert_T(T):
    """dro_sfm/geometry/pose_utils.py:89-94."""
    Ti = torch.eye(4, device=T.device, dtype=T.dtype).repeat([len(T), 1, 1])
    Ti[:, :3, :3] = torch.transpose(T[:, :3, :3], -2, -1)
    Ti[:, :3, -1] = torch.bmm(-1. * Ti[:, :3, :3], T[:, :3, -1].unsqueeze(-1)).squeeze(-1)
    return Ti


def euler_to_R(angle):
    """dro_sfm/geometry/pose_utils.py:40-69 -- R = Rx @ Ry @ Rz."""
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    zero = z.detach() * 0
    one = zero.detach() + 1
    cz, sz = torch.cos(z), torch.sin(z)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], dim=1).view(B, 3, 3)
    cy, sy = torch.cos(y), torch.sin(y)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], dim=1).view(B, 3, 3)
    cx, sx = torch.cos(x), torch.sin(x)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], dim=1).view(B, 3, 3)
    return Rx.bmm(Ry).bmm(Rz)


def pose_vec_to_T(vec):
    """dro_sfm/geometry/pose.py:38-45 + pose_utils.py:73-85 (mode 'euler')."""
    R = euler_to_R(vec[:, 3:])
    T = torch.eye(4, device=vec.device, dtype=vec.dtype).repeat([len(vec), 1, 1])
    T[:, :3, :3] = R
    T[:, :3, -1] = vec[:, :3]
    return T


def pixel_grid(B, H, W, dtype, device="cpu"):
    """dro_sfm/utils/image.py:267-332 -- [x, y, 1] grid, un-normalised."""
    xs = torch.linspace(0, W - 1, W, device=device, dtype=dtype)
    ys = torch.linspace(0, H - 1, H, device=device, dtype=dtype)
    ys, xs = torch.meshgrid([ys, xs], indexing="ij")
    xs, ys = xs.repeat([B, 1, 1]), ys.repeat([B, 1, 1])
    return torch.stack([xs, ys, torch.ones_like(xs)], dim=1)


def _transform(T, X):
    """dro_sfm/geometry/pose.py:79-85 -- bmm, then a separate broadcast add."""
    B, _, H, W = X.shape
    out = T[:, :3, :3].bmm(X.reshape(B, 3, -1)) + T[:, :3, -1].unsqueeze(-1)
    return out.view(B, 3, H, W)


def _eye_T(B, dtype, device):
    return torch.eye(4, device=device, dtype=dtype).repeat([B, 1, 1])


# ----------------------------------------------------------------------------------------------
# depth helpers
# ----------------------------------------------------------------------------------------------
def inv2depth(inv_depth):
    """dro_sfm/utils/depth.py:102-121."""
    depth = 1. / inv_depth.clamp(min=1e-6)
    depth[inv_depth <= 0.] = 0.
    return depth


# ----------------------------------------------------------------------------------------------
# Camera.reconstruct / Camera.project
# ----------------------------------------------------------------------------------------------
def reconstruct(depth, K, Tcw=None, frame="w"):
    """dro_sfm/geometry/camera.py:111-147.  ``Tcw`` is the camera's own pose (identity if None)."""
    B, C, H, W = depth.shape
    assert C == 1
    grid = pixel_grid(B, H, W, depth.dtype, depth.device).view(B, 3, -1)
    rays = K_inverse(K).bmm(grid).view(B, 3, H, W)
    Xc = rays * depth
    if frame == "c":
        return Xc
    if frame == "w":
        Tcw = _eye_T(B, depth.dtype, depth.device) if Tcw is None else Tcw
        return _transform(invert_T(Tcw), Xc)
    raise ValueError("Unknown reference frame {}".format(frame))


def project(X, K, Tcw=None, frame="w", normalize=True):
    """dro_sfm/geometry/camera.py:149-194."""
    B, C, H, W = X.shape
    assert C == 3
    if frame == "c":
        Xc = K.bmm(X.reshape(B, 3, -1))
    elif frame == "w":
        Tcw = _eye_T(B, X.dtype, X.device) if Tcw is None else Tcw
        Xc = K.bmm(_transform(Tcw, X).view(B, 3, -1))
    else:
        raise ValueError("Unknown reference frame {}".format(frame))
    Xp, Yp = Xc[:, 0], Xc[:, 1]
    Zp = Xc[:, 2].clamp(min=1e-5)
    if normalize:
        u = 2 * (Xp / Zp) / (W - 1) - 1.
        v = 2 * (Yp / Zp) / (H - 1) - 1.
    else:
        u, v = Xp / Zp, Yp / Zp
    return torch.stack([u, v], dim=-1).view(B, H, W, 2)


def warp_coords(depth, K, ref_K, T, scale=1.0, normalize=True):
    """reconstruct(target cam, identity pose) -> project(reference cam with pose T).

    The composition used by camera_utils.py:50-52, DepthPoseNet.py:83-90 and
    supervised_loss.py:283-289.  ``K.float()`` there is ``K.to(depth.dtype)`` here so that the
    float64 comparator stays in float64.
    """
    Kt = scale_K(K.to(depth.dtype), scale)
    Kr = scale_K(ref_K.to(depth.dtype), scale)
    world = reconstruct(depth, Kt, None, "w")
    return project(world, Kr, T, "w", normalize)


# ----------------------------------------------------------------------------------------------
# gather / view synthesis
# ----------------------------------------------------------------------------------------------
def grid_gather(src, coords, padding_mode="zeros"):
    """The F.grid_sample call of camera_utils.py:55 / DepthPoseNet.py:92 (bilinear, align_corners)."""
    return F.grid_sample(src, coords, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def view_synthesis(src, depth, K, ref_K, T, scale=1.0, padding_mode="zeros"):
    """dro_sfm/geometry/camera_utils.py:23-56."""
    assert depth.size(1) == 1
    return grid_gather(src, warp_coords(depth, K, ref_K, T, scale), padding_mode)


# ----------------------------------------------------------------------------------------------
# feature-metric cost (recurrent optimiser)
# ----------------------------------------------------------------------------------------------
def _as_T(pose):
    """[B,6] euler vector -> [B,4,4] (Pose.from_vec); a [B,4,4] matrix passes through."""
    return pose if pose.dim() == 3 else pose_vec_to_T(pose)


def feat_cost_each(pose_vec, fmap, fmap_ref, depth, K, ref_K, scale):
    """dro_sfm/networks/depth_pose/DepthPoseNet.py:76-96 -- per-channel squared difference."""
    coords = warp_coords(depth, K, ref_K, _as_T(pose_vec), scale)
    return (fmap - grid_gather(fmap_ref, coords, "zeros")) ** 2


def depth_cost(inv_depth, fmap, fmaps_ref, pose_list, K, ref_K, scale):
    """dro_sfm/networks/depth_pose/DepthPoseNet.py:98-105 -- mean over views."""
    costs = [feat_cost_each(p, fmap, fr, inv2depth(inv_depth), K, ref_K, scale)
             for p, fr in zip(pose_list, fmaps_ref)]
    return torch.stack(costs, dim=1).mean(dim=1)


# ----------------------------------------------------------------------------------------------
# photometric loss
# ----------------------------------------------------------------------------------------------
def ssim(x, y, C1=1e-4, C2=9e-4):
    """dro_sfm/losses/multiview_photometric_loss_mf.py:15-54 (3x3, stride 1, reflection pad 1)."""
    x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    mu_xy, mu_xx, mu_yy = mu_x * mu_y, mu_x.pow(2), mu_y.pow(2)
    sig_x = F.avg_pool2d(x.pow(2), 3, 1) - mu_xx
    sig_y = F.avg_pool2d(y.pow(2), 3, 1) - mu_yy
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_xy
    num = (2 * mu_xy + C1) * (2 * sig_xy + C2)
    den = (mu_xx + mu_yy + C1) * (sig_x + sig_y + C2)
    return num / den


def photometric_map(est, image, ssim_w=0.85, C1=1e-4, C2=9e-4, clip=0.0):
    """multiview_photometric_loss_mf.py:175-229 for one (estimate, image) pair -> [B,1,H,W]."""
    l1 = torch.abs(est - image)
    if ssim_w > 0.0:
        s = torch.clamp((1. - ssim(est, image, C1, C2)) / 2., 0., 1.)
        pm = ssim_w * s.mean(1, True) + (1 - ssim_w) * l1.mean(1, True)
    else:
        pm = l1
    if clip > 0.0:
        mean, std = pm.mean(), pm.std()
        pm = torch.clamp(pm, max=float(mean + clip * std))
    return pm


def photometric_loss(image, context, inv_depths, K, ref_K, poses, ssim_w=0.85, C1=1e-4, C2=9e-4,
                     reduce_op="min", clip=0.0, padding_mode="zeros", automask=True, gamma=0.85,
                     forced_sel=None, maps_out=None):
    """multiview_photometric_loss_mf.py:132-171,231-269,333-353.

    ``poses[v][i]`` is the [B,4,4] target->source transform of view v at prediction i.
    Map order per prediction is [warp_0, unwarp_0, warp_1, unwarp_1, ...] (lines 343-351).

    Checker-only extras ('min' reduce): ``maps_out`` (a list) receives the detached [B,2V,H,W] map stack of every
    prediction; ``forced_sel[i]`` ([B,H,W] long: source view v >= 0, -1 = an un-warped / auto-mask map) replaces the
    per-pixel arg-min, so that gradients of two fp32 evaluations can be compared for the SAME selection (the per-pixel
    min is discontinuous: at a near-tie the winner -- and with it a whole gradient term -- may differ).
    """
    n = len(inv_depths)
    per_pred = [[] for _ in range(n)]
    for ref_image, pose in zip(context, poses):
        for i in range(n):
            warped = view_synthesis(ref_image, inv2depth(inv_depths[i]), K, ref_K, pose[i], 1.0,
                                    padding_mode)
            per_pred[i].append(photometric_map(warped, image, ssim_w, C1, C2, clip))
        if automask:
            for i in range(n):
                per_pred[i].append(photometric_map(ref_image, image, ssim_w, C1, C2, clip))
    total = 0.0
    for i in range(n):
        if reduce_op == "mean":
            li = sum([m.mean() for m in per_pred[i]]) / len(per_pred[i])
        elif reduce_op == "min":
            stack = torch.cat(per_pred[i], 1)
            if maps_out is not None:
                maps_out.append(stack.detach())
            if forced_sel is None:
                li = stack.min(1, True)[0].mean()
            else:
                idx = forced_sel[i].long()
                if automask:
                    auto = stack[:, 1::2].argmin(1) * 2 + 1
                    idx = torch.where(idx < 0, auto, idx * 2)
                li = stack.gather(1, idx.unsqueeze(1)).mean()
        else:
            raise NotImplementedError("Unknown photometric_reduce_op: {}".format(reduce_op))
        total = total + gamma ** (n - i - 1) * li
    return total


def smoothness_loss(inv_depths, image, weight=0.001):
    """multiview_photometric_loss_mf.py:273-299 + utils/depth.py:147-199 + utils/image.py:134-162."""
    n = len(inv_depths)
    total = 0.0
    for i, d in enumerate(inv_depths):
        dn = d / d.mean(2, True).mean(3, True).clamp(min=1e-6)
        dgx = dn[:, :, :, :-1] - dn[:, :, :, 1:]
        dgy = dn[:, :, :-1, :] - dn[:, :, 1:, :]
        igx = image[:, :, :, :-1] - image[:, :, :, 1:]
        igy = image[:, :, :-1, :] - image[:, :, 1:, :]
        wx = torch.exp(-torch.mean(torch.abs(igx), 1, keepdim=True))
        wy = torch.exp(-torch.mean(torch.abs(igy), 1, keepdim=True))
        total = total + ((dgx * wx).abs().mean() + (dgy * wy).abs().mean()) / 2 ** i
    return weight * (total / n)


def multiview_photometric_decay_loss(image, context, inv_depths, K, ref_K, poses, smooth_w=0.001, **kw):
    """MultiViewPhotometricDecayLoss.forward, multiview_photometric_loss_mf.py:303-361 -> (loss[1], metrics)."""
    photo = photometric_loss(image, context, inv_depths, K, ref_K, poses, **kw)
    metrics = {}
    loss = photo
    if smooth_w > 0.0:
        sm = smoothness_loss(inv_depths, image, smooth_w)
        metrics["smoothness_loss"] = sm.detach()
        loss = loss + sm
    # Reference quirk: add_metric stores ``photometric_loss.detach()`` (line 268, shares storage) and
    # line 356 then does ``loss += smoothness`` IN PLACE on that same tensor, so the logged
    # 'photometric_loss' metric equals the total loss whenever the smoothness term is enabled.
    metrics["photometric_loss"] = loss.detach()
    return loss.unsqueeze(0), metrics


# ----------------------------------------------------------------------------------------------
# supervised losses
# ----------------------------------------------------------------------------------------------
def reproj_coords(T, K, ref_K, depth):
    """dro_sfm/losses/supervised_loss.py:279-291 -> (coords [B,H,W,2], valid mask)."""
    c = warp_coords(depth, K, ref_K, T, 1)
    return c, (c >= -1) & (c <= 1)


def reproj_pose_loss(pred_poses, gt_poses, gt_depth, K, ref_K, min_depth, max_depth, gamma=0.85):
    """dro_sfm/losses/supervised_loss.py:293-325.  pred_poses[v][i], gt_poses[v] are [B,4,4]."""
    n = len(pred_poses[0])
    dmask = ((gt_depth > min_depth) & (gt_depth < max_depth / 4.0)).permute(0, 2, 3, 1)
    total, wsum = 0, 0
    for i in range(n):
        w = gamma ** (n - i - 1)
        wsum += w
        li = 0
        for v, T_gt in enumerate(gt_poses):
            c_gt, m_gt = reproj_coords(T_gt, K, ref_K, gt_depth)
            c_pr, m_pr = reproj_coords(pred_poses[v][i], K, ref_K, gt_depth)
            valid = m_gt * m_pr * dmask
            li = li + torch.mean(valid * torch.abs(c_pr - c_gt).clamp(-1, 1))
        total = total + (li / len(gt_poses)) * w
    return total / wsum


def supervised_depth_loss(inv_depths, gt_inv_depth, min_depth, max_depth, gamma=0.85):
    """dro_sfm/losses/supervised_loss.py:244-277 (all predictions at the GT resolution)."""
    n = len(inv_depths)
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    total, wsum = 0, 0
    for i in range(n):
        w = gamma ** (n - i - 1)
        wsum += w
        valid = ((gt_inv_depth > lo) & (gt_inv_depth < hi)).detach().squeeze(1)
        total = total + w * torch.mean(valid * torch.abs(gt_inv_depth - inv_depths[i]).squeeze(1))
    return total / wsum


# ----------------------------------------------------------------------------------------------
# convex up-sampling
# ----------------------------------------------------------------------------------------------
def upsample_depth(depth, mask, ratio=8):
    """dro_sfm/networks/depth_pose/DepthPoseNet.py:63-74."""
    N, _, H, W = depth.shape
    m = torch.softmax(mask.view(N, 1, 9, ratio, ratio, H, W), dim=2)
    nb = F.unfold(depth, [3, 3], padding=1).view(N, 1, 9, 1, 1, H, W)
    up = torch.sum(m * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 1, ratio * H, ratio * W)
