"""torch.autograd bindings of the drosfm_b200 kernels (include/drosfm_b200.h).

Each Function allocates its outputs with torch, launches the kernel on the current stream through
the C ABI and implements backward with the explicit backward kernel.  No host synchronisation, no
CPU path.  Poses are passed either as [B,4,4] matrices (``Pose.mat``) or as [B,6] euler vectors
(then ``Pose.from_vec(vec, 'euler')``, pose.py:38-45, is evaluated inside the kernel).
"""
import torch

from . import _lib as L

__all__ = ["reconstruct", "project", "warp_coords", "grid_gather", "view_synthesis", "feat_cost",
           "photometric_loss", "smoothness_loss", "reproj_pose_loss"]


def _pose_kind(pose):
    if pose is None:
        return L.POSE_IDENTITY
    if pose.dim() == 3 and tuple(pose.shape[-2:]) == (4, 4):
        return L.POSE_MAT4
    if pose.dim() == 2 and pose.shape[-1] == 6:
        return L.POSE_EULER6
    raise ValueError("pose must be [B,4,4] or [B,6], got {}".format(tuple(pose.shape)))


def _padding(mode):
    if mode == "zeros":
        return L.PAD_ZEROS
    if mode == "border":
        return L.PAD_BORDER
    raise NotImplementedError("padding_mode {!r} is not supported (zeros | border)".format(mode))


def _no_grad_for(name, t):
    if t is not None and torch.is_tensor(t) and t.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("dro_sfm_b200: gradient w.r.t. {} is not implemented".format(name))


# ------------------------------------------------------------------------------------------------
# Camera.reconstruct / Camera.project
# ------------------------------------------------------------------------------------------------
class _Reconstruct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, K, Twc):
        L.require_cuda(depth, K, Twc)
        B, C, H, W = depth.shape
        depth = L.f32c(depth)
        K, kd = L.k_arg(K)
        Twc = L.f32c(Twc)
        out = torch.empty(B, 3, H, W, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            L.check(L.lib().drosfm_reconstruct_fwd(L.ptr(depth), L.ptr(K), kd, L.ptr(Twc), L.ptr(out), B, H, W, L.stream()),
                    "reconstruct_fwd")
        ctx.save_for_backward(K, Twc)
        ctx.kd = kd
        return out

    @staticmethod
    def backward(ctx, g):
        K, Twc = ctx.saved_tensors
        g = L.f32c(g)
        B, _, H, W = g.shape
        gd = torch.empty(B, 1, H, W, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            L.check(L.lib().drosfm_reconstruct_bwd(L.ptr(g), L.ptr(K), ctx.kd, L.ptr(Twc), L.ptr(gd), B, H, W, L.stream()),
                    "reconstruct_bwd")
        return gd, None, None


def reconstruct(depth, K, Twc=None):
    """Camera.reconstruct (camera.py:111-147).  Twc=None is frame 'c'; otherwise [B,4,4] world<-camera."""
    if depth.shape[1] != 1:
        raise AssertionError("depth must be [B,1,H,W]")
    _no_grad_for("K", K)
    _no_grad_for("the reconstructing camera's pose", Twc)
    return _Reconstruct.apply(depth, K, Twc)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, K, Tcw, normalize):
        L.require_cuda(X, K, Tcw)
        B, C, H, W = X.shape
        X = L.f32c(X)
        K, kd = L.k_arg(K)
        Tcw = L.f32c(Tcw)
        uv = torch.empty(B, H, W, 2, device=X.device, dtype=torch.float32)
        with torch.cuda.device(X.device):
            L.check(L.lib().drosfm_project_fwd(L.ptr(X), L.ptr(K), kd, L.ptr(Tcw), L.ptr(uv), B, H, W, int(normalize),
                                               L.stream()), "project_fwd")
        ctx.save_for_backward(X, K, Tcw)
        ctx.kd, ctx.normalize = kd, int(normalize)
        return uv

    @staticmethod
    def backward(ctx, g):
        X, K, Tcw = ctx.saved_tensors
        g = L.f32c(g)
        B, _, H, W = X.shape
        gX = torch.empty_like(X) if ctx.needs_input_grad[0] else None
        gT = torch.empty(B, 4, 4, device=X.device, dtype=torch.float32) if (Tcw is not None and ctx.needs_input_grad[2]) else None
        with torch.cuda.device(X.device):
            ws = L.workspace(X.device, B) if gT is not None else None
            L.check(L.lib().drosfm_project_bwd(L.ptr(g), L.ptr(X), L.ptr(K), ctx.kd, L.ptr(Tcw), L.ptr(gX), L.ptr(gT), L.ptr(ws),
                                               B, H, W, ctx.normalize, L.stream()), "project_bwd")
        return gX, None, gT, None


def project(X, K, Tcw=None, normalize=True):
    """Camera.project (camera.py:149-194).  Tcw=None is frame 'c'."""
    if X.shape[1] != 3:
        raise AssertionError("points must be [B,3,H,W]")
    _no_grad_for("K", K)
    return _Project.apply(X, K, Tcw, bool(normalize))


# ------------------------------------------------------------------------------------------------
# fused reconstruct -> project
# ------------------------------------------------------------------------------------------------
class _WarpCoords(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, pose, K, Kref, scale, normalize, depth_kind, want_mask):
        L.require_cuda(depth, pose, K, Kref)
        B, _, H, W = depth.shape
        depth, pose = L.f32c(depth), L.f32c(pose)
        kind = _pose_kind(pose)
        cams, keep = L.make_cams(K, Kref, scale, None, None, pose, kind)
        uv = torch.empty(B, H, W, 2, device=depth.device, dtype=torch.float32)
        mask = torch.empty(B, H, W, 2, device=depth.device, dtype=torch.uint8) if want_mask else None
        with torch.cuda.device(depth.device):
            L.check(L.lib().drosfm_warp_coords_fwd(L.ptr(depth), depth_kind, cams, L.ptr(uv), L.ptr(mask), B, H, W,
                                                   int(normalize), L.stream()), "warp_coords_fwd")
        ctx.save_for_backward(depth, pose, keep[0], keep[1])
        ctx.cfg = (float(scale), int(normalize), depth_kind, kind)
        if want_mask:
            ctx.mark_non_differentiable(mask)
            return uv, mask
        return uv

    @staticmethod
    def backward(ctx, g, *unused):
        depth, pose, K, Kref = ctx.saved_tensors
        scale, normalize, depth_kind, kind = ctx.cfg
        g = L.f32c(g)
        B, _, H, W = depth.shape
        cams, _ = L.make_cams(K, Kref, scale, None, None, pose, kind)
        gd = torch.empty_like(depth) if ctx.needs_input_grad[0] else None
        gp = torch.empty_like(pose) if (pose is not None and ctx.needs_input_grad[1]) else None
        with torch.cuda.device(depth.device):
            ws = L.workspace(depth.device, B) if gp is not None else None
            L.check(L.lib().drosfm_warp_coords_bwd(L.ptr(g), L.ptr(depth), depth_kind, cams, L.ptr(gd), L.ptr(gp), L.ptr(ws),
                                                   B, H, W, normalize, L.stream()), "warp_coords_bwd")
        return gd, gp, None, None, None, None, None, None


def warp_coords(depth, pose, K, Kref=None, scale=1.0, normalize=True, inverse_depth=False, want_mask=False):
    """reconstruct(target camera at identity) -> project(source camera with `pose`), fused.

    Replaces camera_utils.py:50-52 / DepthPoseNet.py:83-90 / supervised_loss.py:283-290.
    Returns uv [B,H,W,2] (and the (uv>=-1)&(uv<=1) mask when want_mask)."""
    Kref = K if Kref is None else Kref
    _no_grad_for("K", K)
    out = _WarpCoords.apply(depth, pose, K, Kref, float(scale), bool(normalize),
                            L.INV_DEPTH if inverse_depth else L.DEPTH, bool(want_mask))
    if want_mask:
        return out[0], out[1].view(torch.bool)
    return out


# ------------------------------------------------------------------------------------------------
# grid_sample / view_synthesis
# ------------------------------------------------------------------------------------------------
class _GridGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, uv, padding):
        L.require_cuda(src, uv)
        src, uv = L.f32c(src), L.f32c(uv)
        B, C, Hs, Ws = src.shape
        _, H, W, _ = uv.shape
        out = torch.empty(B, C, H, W, device=src.device, dtype=torch.float32)
        with torch.cuda.device(src.device):
            L.check(L.lib().drosfm_grid_gather_fwd(L.ptr(src), L.ptr(uv), L.ptr(out), B, C, Hs, Ws, H, W, padding, L.stream()),
                    "grid_gather_fwd")
        ctx.save_for_backward(src, uv)
        ctx.padding = padding
        return out

    @staticmethod
    def backward(ctx, g):
        src, uv = ctx.saved_tensors
        g = L.f32c(g)
        B, C, Hs, Ws = src.shape
        _, H, W, _ = uv.shape
        gs = torch.zeros_like(src) if ctx.needs_input_grad[0] else None
        gu = torch.empty_like(uv) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(src.device):
            L.check(L.lib().drosfm_grid_gather_bwd(L.ptr(g), L.ptr(src), L.ptr(uv), L.ptr(gs), L.ptr(gu), B, C, Hs, Ws, H, W,
                                                   ctx.padding, L.stream()), "grid_gather_bwd")
        return gs, gu, None


def grid_gather(src, uv, padding_mode="zeros"):
    """F.grid_sample(src, uv, mode='bilinear', padding_mode, align_corners=True)."""
    return _GridGather.apply(src, uv, _padding(padding_mode))


class _ViewSynthesis(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, depth, pose, K, Kref, scale, padding, depth_kind):
        L.require_cuda(src, depth, pose, K, Kref)
        src, depth, pose = L.f32c(src), L.f32c(depth), L.f32c(pose)
        B, C, Hs, Ws = src.shape
        _, _, H, W = depth.shape
        kind = _pose_kind(pose)
        cams, keep = L.make_cams(K, Kref, scale, None, None, pose, kind)
        out = torch.empty(B, C, H, W, device=src.device, dtype=torch.float32)
        with torch.cuda.device(src.device):
            L.check(L.lib().drosfm_view_synthesis_fwd(L.ptr(src), L.ptr(depth), depth_kind, cams, L.ptr(out), B, C, Hs, Ws, H, W,
                                                      padding, L.stream()), "view_synthesis_fwd")
        ctx.save_for_backward(src, depth, pose, keep[0], keep[1])
        ctx.cfg = (float(scale), padding, depth_kind, kind)
        return out

    @staticmethod
    def backward(ctx, g):
        src, depth, pose, K, Kref = ctx.saved_tensors
        scale, padding, depth_kind, kind = ctx.cfg
        g = L.f32c(g)
        B, C, Hs, Ws = src.shape
        _, _, H, W = depth.shape
        cams, _ = L.make_cams(K, Kref, scale, None, None, pose, kind)
        gs = torch.zeros_like(src) if ctx.needs_input_grad[0] else None
        gd = torch.empty_like(depth) if ctx.needs_input_grad[1] else None
        gp = torch.empty_like(pose) if (pose is not None and ctx.needs_input_grad[2]) else None
        with torch.cuda.device(src.device):
            ws = L.workspace(src.device, B) if gp is not None else None
            L.check(L.lib().drosfm_view_synthesis_bwd(L.ptr(g), L.ptr(src), L.ptr(depth), depth_kind, cams, L.ptr(gs), L.ptr(gd),
                                                      L.ptr(gp), L.ptr(ws), B, C, Hs, Ws, H, W, padding, L.stream()),
                    "view_synthesis_bwd")
        return gs, gd, gp, None, None, None, None, None


def view_synthesis(src, depth, pose, K, Kref=None, scale=1.0, padding_mode="zeros", inverse_depth=False):
    """Fused view_synthesis (camera_utils.py:23-56) for a target camera at the identity."""
    if depth.shape[1] != 1:
        raise AssertionError("depth must be [B,1,H,W]")
    Kref = K if Kref is None else Kref
    _no_grad_for("K", K)
    return _ViewSynthesis.apply(src, depth, pose, K, Kref, float(scale), _padding(padding_mode),
                                L.INV_DEPTH if inverse_depth else L.DEPTH)


# ------------------------------------------------------------------------------------------------
# feature-metric cost
# ------------------------------------------------------------------------------------------------
def _layout_of(t):
    """NHWC when the tensor is stored channels_last (and not also plain-contiguous), else NCHW."""
    if t.dim() == 4 and t.shape[1] > 1 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous():
        return L.NHWC
    return L.NCHW


def _as_layout(t, layout):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous(memory_format=torch.channels_last) if layout == L.NHWC else t.contiguous()


class _FeatCost(torch.autograd.Function):
    """inputs: depth, fmap, K, Kref, scale, depth_kind, V, ref_0..ref_{V-1}, pose_0..pose_{V-1}"""

    @staticmethod
    def forward(ctx, depth, fmap, K, Kref, scale, depth_kind, V, *views):
        refs, poses = views[:V], views[V:]
        L.require_cuda(depth, fmap, K, Kref, *views)
        layout = _layout_of(fmap)
        if layout == L.NHWC and fmap.shape[1] % 4 != 0:
            layout = L.NCHW
        fmap = _as_layout(fmap, layout)
        refs = [_as_layout(r, layout) for r in refs]
        poses = [L.f32c(p) for p in poses]
        depth = L.f32c(depth)
        B, C, h, w = fmap.shape
        kind = _pose_kind(poses[0])
        cams, keep = L.make_cams(K, Kref, scale, None, None, None, kind)
        cost = torch.empty_like(fmap)   # preserves the storage layout
        with torch.cuda.device(fmap.device):
            L.check(L.lib().drosfm_feat_cost_fwd(L.ptr(fmap), L.ptr_array(refs), L.ptr(depth), depth_kind, cams,
                                                 L.ptr_array(poses), V, L.ptr(cost), B, C, h, w, layout, L.stream()),
                    "feat_cost_fwd")
        ctx.save_for_backward(depth, fmap, keep[0], keep[1], *refs, *poses)
        ctx.cfg = (float(scale), depth_kind, V, kind, layout)
        return cost

    @staticmethod
    def backward(ctx, g):
        scale, depth_kind, V, kind, layout = ctx.cfg
        depth, fmap, K, Kref = ctx.saved_tensors[:4]
        refs = ctx.saved_tensors[4:4 + V]
        poses = ctx.saved_tensors[4 + V:4 + 2 * V]
        g = _as_layout(g, layout)
        B, C, h, w = fmap.shape
        need = ctx.needs_input_grad
        cams, _ = L.make_cams(K, Kref, scale, None, None, None, kind)
        g_fmap = torch.empty_like(fmap) if need[1] else None
        # one zero-filled slab for everything the kernel accumulates into (a single memset)
        n_ref = sum(1 for v in range(V) if need[7 + v])
        slab = torch.zeros(n_ref * fmap.numel() + (depth.numel() if need[0] else 0), device=fmap.device, dtype=torch.float32)
        g_refs, off = [], 0
        for v in range(V):
            if need[7 + v]:
                flat = slab[off:off + fmap.numel()]
                off += fmap.numel()
                g_refs.append(flat.view(B, h, w, C).permute(0, 3, 1, 2) if layout == L.NHWC else flat.view(B, C, h, w))
            else:
                g_refs.append(None)
        g_depth = slab[off:off + depth.numel()].view_as(depth) if need[0] else None
        g_poses = [torch.empty_like(poses[v]) if need[7 + V + v] else None for v in range(V)]
        with torch.cuda.device(fmap.device):
            ws = L.workspace(fmap.device, V * B) if any(p is not None for p in g_poses) else None
            L.check(L.lib().drosfm_feat_cost_bwd(L.ptr(g), L.ptr(fmap), L.ptr_array(refs), L.ptr(depth), depth_kind, cams,
                                                 L.ptr_array(poses), V, L.ptr(g_fmap), L.ptr_array(g_refs), L.ptr(g_depth),
                                                 L.ptr_array(g_poses), L.ptr(ws), B, C, h, w, layout, L.stream()),
                    "feat_cost_bwd")
        return (g_depth, g_fmap, None, None, None, None, None, *g_refs, *g_poses)


def feat_cost(depth, fmap, fmaps_ref, poses, K, Kref=None, scale=1.0, inverse_depth=False):
    """(1/V) sum_v (fmap - warp_v(fmap_ref_v))^2  -> [B,C,h,w]  (DepthPoseNet.py:76-105).

    fmaps_ref / poses: sequences of V tensors; poses are [B,6] euler vectors or [B,4,4] matrices.
    The result keeps fmap's storage layout (NCHW or channels_last)."""
    V = len(fmaps_ref)
    if V < 1 or V > L.MAX_VIEWS or len(poses) != V:
        raise ValueError("feat_cost needs 1..{} views with one pose each".format(L.MAX_VIEWS))
    Kref = K if Kref is None else Kref
    _no_grad_for("K", K)
    return _FeatCost.apply(depth, fmap, K, Kref, float(scale), L.INV_DEPTH if inverse_depth else L.DEPTH, V,
                           *fmaps_ref, *poses)


def photometric_loss(*args, **kwargs):
    raise NotImplementedError


def smoothness_loss(*args, **kwargs):
    raise NotImplementedError


def reproj_pose_loss(*args, **kwargs):
    raise NotImplementedError
