"""torch.autograd bindings of the drosfm_b200 kernels (include/drosfm_b200.h).

Each Function allocates its outputs with torch, launches the kernel on the current stream through
the C ABI and implements backward with the explicit backward kernel.  No host synchronisation, no
CPU path.  Poses are passed either as [B,4,4] matrices (``Pose.mat``) or as [B,6] euler vectors
(then ``Pose.from_vec(vec, 'euler')``, pose.py:38-45, is evaluated inside the kernel).
"""
import os

import torch

from . import _lib as L

__all__ = ["pose_vec2mat", "reconstruct", "project", "warp_coords", "grid_gather", "view_synthesis", "feat_cost", "feat_cost_batch",
           "photometric_loss", "reproj_pose_loss", "sup_depth_loss", "upsample_depth", "images_u8_to_f32", "post_process_inv_depth", "depth_metrics", "split_channels_last_sink"]


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
# Pose.from_vec
# ------------------------------------------------------------------------------------------------
class _PoseVec2Mat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec):
        L.require_cuda(vec)
        vec = L.f32c(vec)
        N = vec.shape[0]
        mat = torch.empty(N, 4, 4, device=vec.device, dtype=torch.float32)
        with torch.cuda.device(vec.device):
            L.check(L.lib().drosfm_pose_vec2mat_fwd(L.ptr(vec), L.ptr(mat), N, L.stream()), "pose_vec2mat_fwd")
        ctx.save_for_backward(vec)
        return mat

    @staticmethod
    def backward(ctx, g):
        (vec,) = ctx.saved_tensors
        g = L.f32c(g)
        gv = torch.empty_like(vec)
        with torch.cuda.device(vec.device):
            L.check(L.lib().drosfm_pose_vec2mat_bwd(L.ptr(g), L.ptr(vec), L.ptr(gv), vec.shape[0], L.stream()),
                    "pose_vec2mat_bwd")
        return gv


def pose_vec2mat(vec):
    """Pose.from_vec(vec, 'euler').mat (pose.py:38-45): [N,6] -> [N,4,4], R = Rx Ry Rz."""
    if vec.dim() != 2 or vec.shape[1] != 6:
        raise ValueError("pose vector must be [N,6]")
    return _PoseVec2Mat.apply(vec)


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


_ZEROS = {}


def _zero_scalar(device):
    """One fp32 zero per device (the stride-0 placeholder gradients of the sinks expand it; never written)."""
    z = _ZEROS.get(device)
    if z is None:
        z = _ZEROS[device] = torch.zeros((), device=device, dtype=torch.float32)
    return z


class GradSink:
    """Shared gradient accumulator of one feature map over all cost calls of a step.

    ``to_channels_last_sink`` (below) creates one per converted map.  Each cost backward adds its
    contribution into ``buffer`` inside the kernel (the scatter into the source maps is atomic anyway, the
    target-map gradient becomes a read-modify-write), instead of returning 2*V*T separate tensors that
    autograd would have to sum with as many element-wise kernels."""
    __slots__ = ("buffer", "dummy", "consumed", "subs")

    def __init__(self):
        self.buffer = None
        self.dummy = None
        self.consumed = False        # set once the backward pass went through: the copy belongs to a finished step
        self.subs = []               # SubSinks of a stacked map (split_channels_last_sink)


def relayout(t, layout):
    """The same logical [B,C,H,W] tensor in the other dense storage layout (L.NHWC = channels_last, L.NCHW):
    drosfm_relayout when `t` is a dense fp32 CUDA tensor in the opposite layout, torch's strided copy otherwise."""
    fmt = torch.channels_last if layout == L.NHWC else torch.contiguous_format
    if t.is_contiguous(memory_format=fmt):
        return t
    other = torch.contiguous_format if layout == L.NHWC else torch.channels_last
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.is_contiguous(memory_format=other)):
        return t.contiguous(memory_format=fmt)
    B, C, H, W = t.shape
    out = torch.empty((B, C, H, W), device=t.device, dtype=t.dtype, memory_format=fmt)
    with torch.cuda.device(t.device):
        L.check(L.lib().drosfm_relayout(L.ptr(t), L.ptr(out), B, C, H, W, layout, L.stream()), "relayout")
    return out


class _ToChannelsLastSink(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, sink):
        ctx.sink = sink
        ctx.nchw_input = t.is_contiguous()
        return relayout(t, L.NHWC)

    @staticmethod
    def backward(ctx, g):
        buf, ctx.sink.buffer = ctx.sink.buffer, None
        ctx.sink.consumed = True
        if buf is None:
            return g, None
        # g is the zero placeholder returned by the first sink-aware consumer plus whatever ordinary
        # consumers of the converted tensor contributed; the bare placeholder (stride-0 view of one zero) adds nothing
        dummy = ctx.sink.dummy
        if not (dummy is not None and g.data_ptr() == dummy.data_ptr() and all(st == 0 for st in g.stride())):
            buf = buf.add_(g)
        # hand the gradient back in the layout of the input: AccumulateGrad / the producer's backward would
        # otherwise re-layout it with a generic strided copy
        return (relayout(buf, L.NCHW) if ctx.nchw_input else buf), None


def to_channels_last_sink(t):
    """channels_last copy of an NCHW feature map whose cost gradients are summed in-kernel (see GradSink)."""
    sink = GradSink()
    out = _ToChannelsLastSink.apply(t, sink)
    out._drosfm_sink = sink if (t.requires_grad and torch.is_grad_enabled()) else None
    out._drosfm_sink_state = sink
    return out


class SubSink:
    """The part of a GradSink that belongs to one batch-slice of a stacked map (see split_channels_last_sink)."""
    __slots__ = ("parent", "start", "stop", "shape", "device", "asked")

    def __init__(self, parent, start, stop, like):
        # shape and device only: a reference to the stacked tensor would close a cycle through its autograd graph and
        # keep a finished step's AccumulateGrad nodes (and their streams) alive until the garbage collector runs
        self.parent, self.start, self.stop, self.asked = parent, start, stop, False
        self.shape, self.device = tuple(like.shape), like.device

    @property
    def consumed(self):
        return self.parent.consumed


class _SplitSink(torch.autograd.Function):
    """torch.split along the batch axis of a stacked channels_last map whose pieces carry sub-sinks: the backward hands
    ONE placeholder to the stacked tensor's sink node instead of materialising a full-size gradient per piece."""

    @staticmethod
    def forward(ctx, stacked, sink, sizes):
        ctx.sink, ctx.sizes = sink, sizes
        ctx.meta = (stacked.shape, stacked.device)
        return tuple(torch.split(stacked.detach(), sizes, dim=0))

    @staticmethod
    def backward(ctx, *gs):
        sink, (shape, device) = ctx.sink, ctx.meta
        dummy = sink.dummy
        real = []
        start = 0
        for g, n in zip(gs, ctx.sizes):
            if g is not None and not (dummy is not None and g.data_ptr() == dummy.data_ptr() and all(st == 0 for st in g.stride())):
                real.append((start, n, g))
            start += n
        if real:                                       # an ordinary consumer of a piece: its gradient joins the sums
            if sink.buffer is None:
                sink.buffer = torch.empty(shape, device=device, dtype=torch.float32, memory_format=torch.channels_last).zero_()
            for start, n, g in real:
                sink.buffer[start:start + n].add_(g)
        for sub in sink.subs:
            sub.asked = False
        if sink.buffer is None:
            return None, None, None
        if sink.dummy is None:
            sink.dummy = _zero_scalar(device)
        return sink.dummy.expand(shape), None, None


def split_channels_last_sink(stacked, sizes):
    """Converts a stacked NCHW map (e.g. the encoder's output for [target, source_1..V]) to channels_last ONCE and splits
    it along the batch axis; every piece adds its cost gradients into its slice of ONE buffer, which returns to NCHW
    through one launch (instead of one conversion, one zero-fill and one back-conversion per piece)."""
    sizes = [int(n) for n in sizes]
    conv = to_channels_last_sink(stacked)
    sink = conv._drosfm_sink_state
    pieces = _SplitSink.apply(conv, sink, sizes)
    start = 0
    for piece, n in zip(pieces, sizes):
        sub = SubSink(sink, start, start + n, conv)
        sink.subs.append(sub)
        piece._drosfm_sink = sub if conv._drosfm_sink is not None else None
        piece._drosfm_sink_state = sub
        start += n
    return list(pieces)


def sink_buffer(sink, like):
    """(buffer, value to return to autograd): the first consumer of a backward pass allocates the zeroed
    buffer and returns a zero placeholder so that the sink node runs; later ones return None.  The buffer
    belongs to ONE backward pass: a callback at the end of the pass drops it if the sink node was not reached
    (torch.autograd.grad(cost, inputs=[pose]) or an aborted pass), so partial sums never leak into the next."""
    if isinstance(sink, SubSink):
        parent = sink.parent
        if parent.buffer is None:
            parent.buffer = torch.empty(sink.shape, device=sink.device, dtype=torch.float32, memory_format=torch.channels_last).zero_()
            buf = parent.buffer

            def end_of_pass(parent=parent, buf=buf):
                if parent.buffer is buf:
                    parent.buffer = None
                for sub in parent.subs:
                    sub.asked = False
            torch.autograd.Variable._execution_engine.queue_callback(end_of_pass)
        if parent.dummy is None:
            parent.dummy = _zero_scalar(like.device)
        piece = parent.buffer[sink.start:sink.stop]
        # as for a plain sink: only the first consumer of a piece returns the placeholder (autograd would otherwise
        # materialise the sum of two stride-0 tensors)
        first, sink.asked = not sink.asked, True
        return piece, (parent.dummy.expand(like.shape) if first else None)
    if sink.buffer is None:
        sink.buffer = torch.zeros_like(like)
        if sink.dummy is None or sink.dummy.shape != like.shape:
            sink.dummy = _zero_scalar(like.device).expand(like.shape)
        buf = sink.buffer

        def end_of_pass(sink=sink, buf=buf):
            if sink.buffer is buf:
                sink.buffer = None
        torch.autograd.Variable._execution_engine.queue_callback(end_of_pass)
        return sink.buffer, sink.dummy
    return sink.buffer, None


class _FeatCost(torch.autograd.Function):
    """inputs: depth, fmap, K, Kref, scale, depth_kind, V, ref_0..ref_{V-1}, pose_0..pose_{V-1}"""

    @staticmethod
    def forward(ctx, depth, fmap, K, Kref, scale, depth_kind, V, *views):
        refs, poses = views[:V], views[V:]
        L.require_cuda(depth, fmap, K, Kref, *views)
        layout = _layout_of(fmap)
        if layout == L.NHWC and fmap.shape[1] % 4 != 0:
            layout = L.NCHW
        sinks = [getattr(x, "_drosfm_sink", None) for x in (fmap, *refs)]
        fmap_l = _as_layout(fmap, layout)
        refs_l = [_as_layout(r, layout) for r in refs]
        # a sink only applies when the tensor is used as handed over (no re-layout copy in between)
        sinks = [s if a is b else None for s, a, b in zip(sinks, (fmap, *refs), (fmap_l, *refs_l))]
        fmap, refs = fmap_l, refs_l
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
        ctx.sinks = sinks
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
        sinks = ctx.sinks

        flags = 0
        g_fmap = ret_fmap = None
        if need[1]:
            if sinks[0] is not None:
                g_fmap, ret_fmap = sink_buffer(sinks[0], fmap)
                flags |= L.ACCUMULATE_FMAP
            else:
                g_fmap = ret_fmap = torch.empty_like(fmap)
        # one zero-filled slab for everything else the kernel accumulates into (a single memset)
        # (the NHWC kernel writes every depth gradient itself; the NCHW one adds partial sums per channel group)
        plain = [v for v in range(V) if need[7 + v] and sinks[1 + v] is None]
        depth_in_slab = need[0] and layout != L.NHWC
        slab = torch.zeros(len(plain) * fmap.numel() + (depth.numel() if depth_in_slab else 0), device=fmap.device,
                           dtype=torch.float32)
        g_refs, ret_refs, off = [], [], 0
        for v in range(V):
            if not need[7 + v]:
                g_refs.append(None)
                ret_refs.append(None)
            elif sinks[1 + v] is not None:
                buf, ret = sink_buffer(sinks[1 + v], refs[v])
                g_refs.append(buf)
                ret_refs.append(ret)
            else:
                flat = slab[off:off + fmap.numel()]
                off += fmap.numel()
                t = flat.view(B, h, w, C).permute(0, 3, 1, 2) if layout == L.NHWC else flat.view(B, C, h, w)
                g_refs.append(t)
                ret_refs.append(t)
        g_depth = None
        if need[0]:
            g_depth = slab[off:off + depth.numel()].view_as(depth) if depth_in_slab else torch.empty_like(depth)
        g_poses = [torch.empty_like(poses[v]) if need[7 + V + v] else None for v in range(V)]
        with torch.cuda.device(fmap.device):
            ws = L.workspace(fmap.device, V * B) if any(p is not None for p in g_poses) else None
            L.check(L.lib().drosfm_feat_cost_bwd(L.ptr(g), L.ptr(fmap), L.ptr_array(refs), L.ptr(depth), depth_kind, cams,
                                                 L.ptr_array(poses), V, L.ptr(g_fmap), L.ptr_array(g_refs), L.ptr(g_depth),
                                                 L.ptr_array(g_poses), L.ptr(ws), B, C, h, w, layout, flags, L.stream()),
                    "feat_cost_bwd")
        return (g_depth, ret_fmap, None, None, None, None, None, *ret_refs, *g_poses)


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


class _FeatCostBatch(torch.autograd.Function):
    """Several independent cost calls in one launch (drosfm_feat_cost_batch_*; channels_last tensors only).
    inputs: K, Kref, scale, spec = ((V, depth_kind), ...), then per job: depth, fmap, ref_0..ref_{V-1}, pose_0..pose_{V-1}.
    outputs: one cost map per job."""

    @staticmethod
    def _split(spec, tensors):
        jobs, pos = [], 0
        for V, dk, dmin, drange in spec:
            jobs.append((tensors[pos], tensors[pos + 1], tensors[pos + 2:pos + 2 + V], tensors[pos + 2 + V:pos + 2 + 2 * V],
                         (dk, dmin, drange)))
            pos += 2 + 2 * V
        return jobs

    @staticmethod
    def forward(ctx, K, Kref, scale, spec, *tensors):
        L.require_cuda(K, Kref, *tensors)
        jobs = _FeatCostBatch._split(spec, tensors)
        B, C, h, w = jobs[0][1].shape
        kind = _pose_kind(jobs[0][3][0])
        cams, keep = L.make_cams(K, Kref, scale, None, None, None, kind)
        table = (L.CostJob * len(jobs))()
        saved, sinks, costs, hold = [], [], [], []
        for k, (depth, fmap, refs, poses, dk) in enumerate(jobs):
            if tuple(fmap.shape) != (B, C, h, w) or _pose_kind(poses[0]) != kind:
                raise ValueError("feat_cost_batch: all jobs must share the feature-map shape and the pose encoding")
            sinks.append([getattr(x, "_drosfm_sink", None) for x in (fmap, *refs)])
            depth, poses = L.f32c(depth), [L.f32c(p) for p in poses]
            cost = torch.empty_like(fmap)
            ra, pa = L.ptr_array(list(refs)), L.ptr_array(poses)
            hold += [ra, pa]
            j = table[k]
            j.fmap, j.fmap_ref, j.depth, j.depth_kind = fmap.data_ptr(), ra, depth.data_ptr(), dk[0]
            j.poses, j.n_views, j.cost, j.disp_min, j.disp_range = pa, len(refs), cost.data_ptr(), dk[1], dk[2]
            saved += [depth, fmap, *refs, *poses]
            costs.append(cost)
        with torch.cuda.device(jobs[0][1].device):
            L.check(L.lib().drosfm_feat_cost_batch_fwd(table, len(jobs), cams, B, C, h, w, L.NHWC, L.stream()), "feat_cost_batch_fwd")
        ctx.save_for_backward(keep[0], keep[1], *saved)
        ctx.cfg = (float(scale), spec, kind)
        ctx.sinks = sinks
        ctx.set_materialize_grads(False)
        return tuple(costs)

    @staticmethod
    def backward(ctx, *g_costs):
        scale, spec, kind = ctx.cfg
        K, Kref = ctx.saved_tensors[:2]
        jobs = _FeatCostBatch._split(spec, ctx.saved_tensors[2:])
        B, C, h, w = jobs[0][1].shape
        dev = jobs[0][1].device
        cams, _ = L.make_cams(K, Kref, scale, None, None, None, kind)
        need = ctx.needs_input_grad[4:]
        live = [k for k in range(len(jobs)) if g_costs[k] is not None]
        rets = [None] * len(need)
        if not live:
            return (None, None, None, None, *rets)
        jt, gt = (L.CostJob * len(live))(), (L.CostJobGrads * len(live))()
        hold, pos_of, pos, slots = [], [], 0, 0
        for V, *_ in spec:
            pos_of.append(pos)
            pos += 2 + 2 * V
        # every plain (non-sink) source-map gradient of the launch lives in one zero-filled slab: a single memset
        n_plain = sum(1 for k in live for v in range(spec[k][0]) if need[pos_of[k] + 2 + v] and ctx.sinks[k][1 + v] is None)
        slab = torch.zeros(n_plain * B * C * h * w, device=dev, dtype=torch.float32) if n_plain else None
        off = 0
        for row, k in enumerate(live):
            depth, fmap, refs, poses, dk = jobs[k]
            V, p0, sinks = spec[k][0], pos_of[k], ctx.sinks[k]
            g = g_costs[k]
            g = g if g.is_contiguous(memory_format=torch.channels_last) and g.dtype == torch.float32 else _as_layout(g, L.NHWC)
            flags, g_fmap = 0, None
            if need[p0 + 1]:
                if sinks[0] is not None:
                    g_fmap, rets[p0 + 1] = sink_buffer(sinks[0], fmap)
                    flags |= L.ACCUMULATE_FMAP
                else:
                    g_fmap = rets[p0 + 1] = torch.empty_like(fmap)
            g_refs = []
            for v in range(V):
                if not need[p0 + 2 + v]:
                    g_refs.append(None)
                elif sinks[1 + v] is not None:
                    buf, rets[p0 + 2 + v] = sink_buffer(sinks[1 + v], refs[v])
                    g_refs.append(buf)
                else:
                    t = slab[off:off + fmap.numel()].view(B, h, w, C).permute(0, 3, 1, 2)
                    off += fmap.numel()
                    g_refs.append(t)
                    rets[p0 + 2 + v] = t
            g_depth = rets[p0] = torch.empty_like(depth) if need[p0] else None
            g_poses = [torch.empty_like(poses[v]) if need[p0 + 2 + V + v] else None for v in range(V)]
            for v in range(V):
                rets[p0 + 2 + V + v] = g_poses[v]
            ra, pa, gra, gpa = L.ptr_array(list(refs)), L.ptr_array(list(poses)), L.ptr_array(g_refs), L.ptr_array(g_poses)
            hold += [ra, pa, gra, gpa, g]
            j, q = jt[row], gt[row]
            j.fmap, j.fmap_ref, j.depth, j.depth_kind, j.poses, j.n_views, j.cost = fmap.data_ptr(), ra, depth.data_ptr(), dk[0], pa, V, None
            j.disp_min, j.disp_range = dk[1], dk[2]
            q.g_cost, q.g_fmap, q.g_fmap_ref = g.data_ptr(), (None if g_fmap is None else g_fmap.data_ptr()), gra
            q.g_depth, q.g_poses, q.flags = (None if g_depth is None else g_depth.data_ptr()), gpa, flags
            slots += V * B
        with torch.cuda.device(dev):
            ws = L.workspace(dev, slots)
            L.check(L.lib().drosfm_feat_cost_batch_bwd(jt, gt, len(live), cams, L.ptr(ws), B, C, h, w, L.NHWC, L.stream()),
                    "feat_cost_batch_bwd")
        return (None, None, None, None, *rets)


def feat_cost_batch(jobs, K, Kref=None, scale=1.0):
    """Several independent feature-metric cost calls in ONE launch.

    jobs: sequence of (depth, fmap, fmaps_ref, poses, kind) -- the arguments of `feat_cost` -- sharing K, Kref, scale, the
    feature-map shape and the pose encoding.  kind: False = depth, True = inverse depth, ("disp", min_depth, max_depth) =
    the network's raw disparity, scaled by disp_to_depth (layers.py:11-20) inside the kernel, gradient w.r.t. the raw map.
    Returns one cost map per job.  The batched kernels run on channels_last maps; any other input falls back to one
    `feat_cost` call per job (same results; the disparity scaling then runs as torch ops)."""
    def kind_of(k):
        if isinstance(k, (tuple, list)):
            if k[0] != "disp":
                raise ValueError("unknown depth kind {!r}".format(k))
            lo, hi = 1.0 / float(k[2]), 1.0 / float(k[1])                  # min_disp, max_disp as Python doubles (layers.py:14-15)
            return (L.DISP, float(torch.tensor(lo, dtype=torch.float32)), float(torch.tensor(hi - lo, dtype=torch.float32)))
        return (L.INV_DEPTH if k else L.DEPTH, 0.0, 1.0)
    jobs = [(d, f, list(fr), list(ps), kind_of(k)) for d, f, fr, ps, k in jobs]
    Kref = K if Kref is None else Kref
    _no_grad_for("K", K)

    def nhwc(t):
        return t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0 and _layout_of(t) == L.NHWC

    ok = 1 <= len(jobs) <= L.MAX_COST_JOBS and all(
        nhwc(f) and all(nhwc(r) for r in fr) and 1 <= len(fr) <= L.MAX_VIEWS and len(ps) == len(fr) for _, f, fr, ps, _ in jobs)
    if not ok:
        out = []
        for d, f, fr, ps, (dk, dmin, drange) in jobs:
            if dk == L.DISP:                                            # disp_to_depth with torch ops, as the reference
                d, dk = dmin + drange * d, L.INV_DEPTH
            out.append(feat_cost(d, f, fr, ps, K, Kref, scale, inverse_depth=(dk == L.INV_DEPTH)))
        return out
    spec = tuple((len(fr), dk, dmin, drange) for _, _, fr, _, (dk, dmin, drange) in jobs)
    flat = [t for d, f, fr, ps, _ in jobs for t in (d, f, *fr, *ps)]
    return list(_FeatCostBatch.apply(K, Kref, float(scale), spec, *flat))


# ------------------------------------------------------------------------------------------------
# photometric + smoothness loss
# ------------------------------------------------------------------------------------------------
# DROSFM_PHOTO_SAVE_WARP=0 makes the photometric backward re-warp instead of re-reading (saves 12 B per pixel,
# view and prediction of activation memory at ~15 % more backward time)
SAVE_WARP = os.environ.get("DROSFM_PHOTO_SAVE_WARP", "1") != "0"
# DROSFM_PHOTO_OVERLAP=0 keeps the whole loss on the caller's stream (no second stream for auto-mask / smoothness).
# OVERLAP may also be set to "serial" (bench.py's per-kernel timing pass): the same split calls, on one stream.
OVERLAP = os.environ.get("DROSFM_PHOTO_OVERLAP", "1") != "0"
# DROSFM_PHOTO_FUSE_BWD=0: the window gradients of the photometric loss are computed in the backward pass (their own
# kernel) instead of by the training forward (two-view losses on the staged path)
FUSE_BWD = os.environ.get("DROSFM_PHOTO_FUSE_BWD", "1") != "0"
FUSE_BWD_VIEWS = (2, 4)      # view counts that take the fused training forward (the library supports every even V)
# DROSFM_PHOTO_RGBX=0 keeps the flat warp on the caller's planar pictures (no RGBx texel copy of the sources)
RGBX = os.environ.get("DROSFM_PHOTO_RGBX", "1") != "0"


def _reduce_op(name):
    if name == "min":
        return L.REDUCE_MIN
    if name == "mean":
        return L.REDUCE_MEAN
    raise NotImplementedError("Unknown photometric_reduce_op: {}".format(name))


class _PhotoLoss(torch.autograd.Function):
    """inputs: image, K, Kref, cfg, V, n, context_0..V-1, inv_depth_0..n-1, pose_{v,i} (v-major).
    outputs: total loss [1]; [photometric term, smoothness term] (non-differentiable metrics)."""

    @staticmethod
    def forward(ctx, image, K, Kref, cfg, V, n, *tensors):
        ssim_w, C1, C2, padding, reduce_op, automask, gamma, smooth_w, depth_kind, clip = cfg
        context = [L.f32c(x) for x in tensors[:V]]
        invs = [L.f32c(x) for x in tensors[V:V + n]]
        poses = [L.f32c(x) for x in tensors[V + n:]]
        L.require_cuda(image, K, Kref, *tensors)
        image = L.f32c(image)
        B, _, H, W = image.shape
        dev = image.device
        kind = _pose_kind(poses[0])
        cams, keep = L.make_cams(K, Kref, 1.0, None, None, None, kind)
        # clip_loss > 0: per-map thresholds from a statistics pass (fused path only); the scratch holds one 8-float slot
        # per photometric map and must be zero on entry
        clip_scratch = torch.zeros(8 * (V + n * V), device=dev, dtype=torch.float32) if clip > 0.0 else None
        opts = L.PhotoOpts(ssim_w, C1, C2, padding, reduce_op, int(automask), gamma, clip,
                           None if clip_scratch is None else clip_scratch.data_ptr())
        # [photometric, smoothness]: each written by its kernel's finisher (zero only if there is no smoothness term)
        losses = (torch.empty if smooth_w > 0.0 else torch.zeros)(2, device=dev, dtype=torch.float32)
        sel = torch.empty(n, B, H, W, device=dev, dtype=torch.uint8) if (reduce_op == L.REDUCE_MIN or clip > 0.0) else None
        # staged path (12 bytes per pixel, view and prediction): the sources are warped once by a flat kernel and
        # the SSIM kernels of both passes read the result; without it everything runs fused and keeps nothing
        keep_warp = SAVE_WARP and any(ctx.needs_input_grad[6 + V:]) and not clip > 0.0 and ssim_w > 0.0
        wsave = torch.empty(n, V, B, 3, H, W, device=dev, dtype=torch.float32) if keep_warp else None
        stats = torch.empty(n, B, 4, device=dev, dtype=torch.float32) if smooth_w > 0.0 else None
        # edge weights of the smoothness term, kept for its backward pass (8 bytes per pixel)
        edge_w = torch.empty(B, 2, H, W, device=dev, dtype=torch.float32) if smooth_w > 0.0 and any(ctx.needs_input_grad[6 + V:6 + V + n]) else None
        # source pictures as RGBx texels for the flat warp and its adjoint (one 128-bit gather per tap)
        rgbx = torch.empty(V, B, H, W, 4, device=dev, dtype=torch.float32) if keep_warp and RGBX else None
        lib = L.lib()
        staged = wsave is not None and bool(OVERLAP)        # split calls
        two_streams = staged and OVERLAP is True            # OVERLAP == "serial": split calls, one stream
        # two views in training: the forward's SSIM pass also produces d loss / d warped (unscaled); the backward of the
        # loss is then the warp adjoint alone, and the warped copy is not kept
        # (the library takes any even V; at 6-8 views its one-block-per-SM instantiation is slower than the separate
        # backward stage -- sweep at 384x1280, V = 8: 3.33 vs 3.10 ms per loss fwd + bwd -- so those stay unfused)
        fused_bwd = staged and FUSE_BWD and V in FUSE_BWD_VIEWS and B * 3 * H * W < 2 ** 31 and B * n <= 65535
        g_warped = torch.empty_like(wsave) if fused_bwd else None
        with torch.cuda.device(dev):
            ws = L.workspace(dev, max(n * B + 1, V * n * B))
            st = L.stream()
            amask = torch.empty(B, H, W, device=dev, dtype=torch.float32) if automask else None
            pa, pc, pi, pp_ = L.ptr_array(context), L.ptr(image), L.ptr_array(invs), L.ptr_array(poses)

            def side_work():
                # independent of the warp: the un-warped auto-mask map and the smoothness term
                if automask:
                    L.check(lib.drosfm_automask_fwd(pc, pa, V, opts, L.ptr(amask), B, H, W, L.stream()), "automask_fwd")
                ev = None
                if two_streams:
                    ev = torch.cuda.Event()
                    ev.record()
                if smooth_w > 0.0:
                    L.check(lib.drosfm_smoothness_fwd(pc, pi, n, smooth_w, L.ptr(stats), L.ptr(losses[1:]),
                                                      L.ptr(L.workspace(dev, n * B + 1)), L.ptr(edge_w), B, H, W, L.stream()), "smoothness_fwd")
                return ev

            if staged:
                # the flat warp of all sources runs on the caller's stream while a second stream produces the
                # auto-mask map (needed by the SSIM stage) and the smoothness term (needed at the end)
                if two_streams:
                    main, side = torch.cuda.current_stream(dev), L.side_stream(dev)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        ev_mask = side_work()
                else:
                    side_work()
                L.check(lib.drosfm_warp_sources_fwd(pa, V, pi, depth_kind, n, cams, pp_, padding, L.ptr(rgbx), L.ptr(wsave), B, H, W, st),
                        "warp_sources_fwd")
                if two_streams:
                    main.wait_event(ev_mask)
                flags = L.PHOTO_WARPED_READY | (L.PHOTO_FUSE_BWD if fused_bwd else 0)
            else:
                side_work()
                flags = 0
            L.check(lib.drosfm_photometric_fwd(pc, pa, V, pi, depth_kind, n, cams, pp_, L.ptr(amask), opts, L.ptr(sel),
                                               L.ptr(losses), L.ptr(ws), L.ptr(wsave), L.ptr(g_warped), flags, B, H, W, st),
                    "photometric_fwd")
            if two_streams:
                main.wait_stream(side)
        total = losses.sum().reshape(1)
        ctx.save_for_backward(image, keep[0], keep[1], sel, stats, None if fused_bwd else wsave, rgbx if staged else None, g_warped,
                              edge_w, *context, *invs, *poses)
        ctx.cfg, ctx.V, ctx.n, ctx.kind = cfg, V, n, kind
        ctx.mark_non_differentiable(losses)
        if sel is None:
            return total, losses
        ctx.mark_non_differentiable(sel)
        return total, losses, sel

    @staticmethod
    def backward(ctx, g_total, *unused):
        ssim_w, C1, C2, padding, reduce_op, automask, gamma, smooth_w, depth_kind, clip = ctx.cfg
        V, n, kind = ctx.V, ctx.n, ctx.kind
        image, K, Kref, sel, stats, wsave, rgbx, g_fused, edge_w = ctx.saved_tensors[:9]
        rest = ctx.saved_tensors[9:]
        context, invs, poses = rest[:V], rest[V:V + n], rest[V + n:]
        B, _, H, W = image.shape
        dev = image.device
        need = ctx.needs_input_grad
        if any(need[6 + v] for v in range(V)) or need[0]:
            raise NotImplementedError("dro_sfm_b200: the fused photometric loss has no gradient w.r.t. the images")
        need_inv = [need[6 + V + i] for i in range(n)]
        need_pose = [need[6 + V + n + k] for k in range(V * n)]
        g = L.f32c(g_total.reshape(1))
        cams, _ = L.make_cams(K, Kref, 1.0, None, None, None, kind)
        opts = L.PhotoOpts(ssim_w, C1, C2, padding, reduce_op, int(automask), gamma, clip, None)
        smooth = smooth_w > 0.0 and any(need_inv)
        # fused backward with a smoothness term on a second stream: both kernels ADD into the same (zero-filled) maps with
        # atomic reductions, so neither waits for the other
        concurrent = g_fused is not None and smooth and OVERLAP is True
        g_inv_slab = (torch.zeros if concurrent else torch.empty)(n, *invs[0].shape, device=dev, dtype=torch.float32) if any(need_inv) else None
        g_invs = [g_inv_slab[i] if need_inv[i] else None for i in range(n)]
        g_pose_slab = torch.empty(V * n, *poses[0].shape, device=dev, dtype=torch.float32) if any(need_pose) else None
        g_poses = [g_pose_slab[k] if need_pose[k] else None for k in range(V * n)]
        g_warped = torch.empty_like(wsave) if wsave is not None else None     # scratch between the two backward stages
        lib = L.lib()
        if g_fused is not None:
            # the forward left d loss / d warped behind (unscaled): the backward is the warp adjoint, scaled by g
            with torch.cuda.device(dev):
                ws = L.workspace(dev, max(n * B + 1, V * n * B))
                st = L.stream()
                pa, pi, pp_ = L.ptr_array(context), L.ptr_array(invs), L.ptr_array(poses)
                two_streams = OVERLAP is True and smooth
                main = torch.cuda.current_stream(dev)
                side = L.side_stream(dev) if two_streams else main
                if smooth:
                    if two_streams:
                        side.wait_stream(main)
                    with torch.cuda.stream(side):
                        L.check(lib.drosfm_smoothness_bwd(L.ptr(g), L.ptr(image), pi, n, smooth_w, L.ptr(stats), L.ptr_array(g_invs),
                                                          2 if concurrent else 0, L.ptr(edge_w), B, H, W, L.stream()), "smoothness_bwd")
                L.check(lib.drosfm_warp_sources_bwd(L.ptr(g_fused), pa, V, pi, depth_kind, n, cams, pp_, padding, L.ptr(rgbx), L.ptr(g),
                                                    L.ptr_array(g_invs), L.ptr_array(g_poses), L.ptr(ws), 1 if smooth else 0, B, H, W, st),
                        "warp_sources_bwd")
                if smooth and two_streams:
                    main.wait_stream(side)
            return (None, None, None, None, None, None, *([None] * V), *g_invs, *g_poses)
        with torch.cuda.device(dev):
            ws = L.workspace(dev, max(n * B + 1, V * n * B))
            st = L.stream()
            pa, pc, pi, pp_ = L.ptr_array(context), L.ptr(image), L.ptr_array(invs), L.ptr_array(poses)
            if wsave is not None and OVERLAP and smooth:
                # the smoothness gradient is written by a second stream while the window-gradient stage runs; the warp
                # adjoint then adds its share to the same buffers
                two_streams = OVERLAP is True
                main, side = torch.cuda.current_stream(dev), (L.side_stream(dev) if two_streams else torch.cuda.current_stream(dev))
                if two_streams:
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    L.check(lib.drosfm_smoothness_bwd(L.ptr(g), pc, pi, n, smooth_w, L.ptr(stats), L.ptr_array(g_invs), 0, L.ptr(edge_w), B, H, W,
                                                      L.stream()), "smoothness_bwd")
                L.check(lib.drosfm_photometric_bwd(L.ptr(g), pc, pa, V, pi, depth_kind, n, cams, pp_, L.ptr(sel), opts, None, None,
                                                   L.ptr(ws), L.ptr(wsave), L.ptr(g_warped), L.PHOTO_NO_ADJOINT, B, H, W, st),
                        "photometric_bwd")
                if two_streams:
                    main.wait_stream(side)
                L.check(lib.drosfm_warp_sources_bwd(L.ptr(g_warped), pa, V, pi, depth_kind, n, cams, pp_, padding, L.ptr(rgbx), None,
                                                    L.ptr_array(g_invs), L.ptr_array(g_poses), L.ptr(ws), 1, B, H, W, st),
                        "warp_sources_bwd")
            else:
                L.check(lib.drosfm_photometric_bwd(L.ptr(g), pc, pa, V, pi, depth_kind, n, cams, pp_, L.ptr(sel), opts,
                                                   L.ptr_array(g_invs), L.ptr_array(g_poses), L.ptr(ws), L.ptr(wsave),
                                                   L.ptr(g_warped), 0, B, H, W, st), "photometric_bwd")
                if smooth:
                    L.check(lib.drosfm_smoothness_bwd(L.ptr(g), pc, pi, n, smooth_w, L.ptr(stats), L.ptr_array(g_invs), 1, L.ptr(edge_w), B, H, W,
                                                      st), "smoothness_bwd")
        return (None, None, None, None, None, None, *([None] * V), *g_invs, *g_poses)


def photometric_loss(image, context, inv_depths, K, ref_K, poses, ssim_w=0.85, C1=1e-4, C2=9e-4, reduce_op="min",
                     padding_mode="zeros", automask=True, smooth_w=0.001, gamma=0.85, inverse_depth=True, want_selection=False,
                     clip=0.0):
    """MultiViewPhotometricDecayLoss.forward (multiview_photometric_loss_mf.py:303-361), fused.

    context: V source images; inv_depths: n predictions [B,1,H,W]; poses[v][i]: [B,4,4] or [B,6].
    Returns (total [1], terms [2]) with terms = detached [photometric, smoothness] values; with want_selection also
    the arg-min view per pixel of every prediction ([n,B,H,W] uint8; 255 = an un-warped / auto-mask map won; None for the
    'mean' reduce op).  clip > 0: every photometric map is clamped at its mean + clip * std (lines 220-227 of the
    reference; thresholds from a statistics pass on the device, no host synchronisation; the selection then reads 254
    where the winning map was clipped, and for 'mean' it is the bit mask of the un-clipped views)."""
    V, n = len(context), len(inv_depths)
    if not (1 <= V <= L.MAX_VIEWS and 1 <= n <= L.MAX_PREDS):
        raise ValueError("photometric_loss supports 1..{} views and 1..{} predictions".format(L.MAX_VIEWS, L.MAX_PREDS))
    if len(poses) != V or any(len(p) != n for p in poses):
        raise ValueError("poses must be a list of V lists of n transforms")
    if automask and reduce_op != "min":
        raise AssertionError("For automasking only the min photometric_reduce_op is supported.")
    if ssim_w < 0.0:
        raise ValueError("ssim_loss_weight must be >= 0")
    if ssim_w == 0.0 and clip > 0.0 and reduce_op != "min":
        raise NotImplementedError("dro_sfm_b200: ssim_loss_weight == 0 with clip_loss > 0 needs photometric_reduce_op='min'")
    for d in inv_depths:
        if tuple(d.shape[-2:]) != tuple(image.shape[-2:]):
            raise NotImplementedError("dro_sfm_b200: predictions must be at the image resolution")
    _no_grad_for("K", K)
    cfg = (float(ssim_w), float(C1), float(C2), _padding(padding_mode), _reduce_op(reduce_op), bool(automask), float(gamma),
           float(smooth_w), L.INV_DEPTH if inverse_depth else L.DEPTH, float(clip))
    flat = [p for pv in poses for p in pv]
    out = _PhotoLoss.apply(image, K, ref_K, cfg, V, n, *context, *inv_depths, *flat)
    if want_selection:
        return out[0], out[1], (out[2] if len(out) > 2 else None)
    return out[0], out[1]


# ------------------------------------------------------------------------------------------------
# reprojection pose loss
# ------------------------------------------------------------------------------------------------
class _ReprojLoss(torch.autograd.Function):
    """inputs: depth, K, Kref, cfg, V, n, gt_0..V-1, pred_{v,i} (v-major)"""

    @staticmethod
    def forward(ctx, depth, K, Kref, cfg, V, n, *tensors):
        min_depth, max_depth, gamma, depth_kind = cfg
        L.require_cuda(depth, K, Kref, *tensors)
        depth = L.f32c(depth)
        gts = [L.f32c(x) for x in tensors[:V]]
        preds = [L.f32c(x) for x in tensors[V:]]
        B, _, H, W = depth.shape
        kind = _pose_kind(preds[0])
        if _pose_kind(gts[0]) != kind:
            raise ValueError("GT and predicted poses must use the same encoding ([B,4,4] or [B,6])")
        cams, keep = L.make_cams(K, Kref, 1.0, None, None, None, kind)
        loss = torch.empty(1, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            ws = L.workspace(depth.device, max(n + 1, V * n * B))
            L.check(L.lib().drosfm_reproj_loss_fwd(L.ptr(depth), depth_kind, cams, L.ptr_array(gts), L.ptr_array(preds), V, n,
                                                   min_depth, max_depth, gamma, L.ptr(loss), L.ptr(ws), B, H, W, L.stream()),
                    "reproj_loss_fwd")
        ctx.save_for_backward(depth, keep[0], keep[1], *gts, *preds)
        ctx.cfg, ctx.V, ctx.n, ctx.kind = cfg, V, n, kind
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        min_depth, max_depth, gamma, depth_kind = ctx.cfg
        V, n, kind = ctx.V, ctx.n, ctx.kind
        depth, K, Kref = ctx.saved_tensors[:3]
        gts, preds = ctx.saved_tensors[3:3 + V], ctx.saved_tensors[3 + V:]
        B, _, H, W = depth.shape
        need = ctx.needs_input_grad
        if need[0] or any(need[6 + v] for v in range(V)):
            raise NotImplementedError("dro_sfm_b200: the reprojection loss has gradients w.r.t. the predicted poses only")
        need_pose = [need[6 + V + k] for k in range(V * n)]
        g = L.f32c(g.reshape(1))
        cams, _ = L.make_cams(K, Kref, 1.0, None, None, None, kind)
        slab = torch.empty(V * n, *preds[0].shape, device=depth.device, dtype=torch.float32)
        g_preds = [slab[k] if need_pose[k] else None for k in range(V * n)]
        with torch.cuda.device(depth.device):
            ws = L.workspace(depth.device, max(n + 1, V * n * B))
            L.check(L.lib().drosfm_reproj_loss_bwd(L.ptr(g), L.ptr(depth), depth_kind, cams, L.ptr_array(gts), L.ptr_array(preds),
                                                   V, n, min_depth, max_depth, gamma, L.ptr_array(g_preds), L.ptr(ws), B, H, W,
                                                   L.stream()), "reproj_loss_bwd")
        return (None, None, None, None, None, None, *([None] * V), *g_preds)


def reproj_pose_loss(pred_poses, gt_poses, gt_depth, K, ref_K, min_depth, max_depth, gamma=0.85, inverse_depth=False):
    """SupervisedDepthPoseLoss.calc_pose_loss (supervised_loss.py:293-325), fused.

    pred_poses[v][i], gt_poses[v]: [B,4,4] (or [B,6]); gt_depth [B,1,H,W] (GT inverse depth when
    inverse_depth=True: inv2depth is applied inside the kernel)."""
    V, n = len(gt_poses), len(pred_poses[0])
    if not (1 <= V <= L.MAX_VIEWS and 1 <= n <= L.MAX_PREDS) or len(pred_poses) != V:
        raise ValueError("reproj_pose_loss supports 1..{} views and 1..{} predictions".format(L.MAX_VIEWS, L.MAX_PREDS))
    _no_grad_for("K", K)
    cfg = (float(min_depth), float(max_depth), float(gamma), L.INV_DEPTH if inverse_depth else L.DEPTH)
    flat = [p for pv in pred_poses for p in pv]
    return _ReprojLoss.apply(gt_depth, K, ref_K, cfg, V, n, *gt_poses, *flat)


# ------------------------------------------------------------------------------------------------
# supervised depth loss
# ------------------------------------------------------------------------------------------------
class _SupDepthLoss(torch.autograd.Function):
    """inputs: gt_inv_depth, cfg, n, inv_depth_0..n-1"""

    @staticmethod
    def forward(ctx, gt, cfg, n, *invs):
        min_depth, max_depth, gamma = cfg
        L.require_cuda(gt, *invs)
        gt = L.f32c(gt)
        invs = [L.f32c(x) for x in invs]
        B, _, H, W = gt.shape
        loss = torch.empty(1, device=gt.device, dtype=torch.float32)
        with torch.cuda.device(gt.device):
            ws = L.workspace(gt.device, n + 1)
            L.check(L.lib().drosfm_sup_depth_loss_fwd(L.ptr(gt), L.ptr_array(invs), n, min_depth, max_depth, gamma, L.ptr(loss),
                                                      L.ptr(ws), B, H, W, L.stream()), "sup_depth_loss_fwd")
        ctx.save_for_backward(gt, *invs)
        ctx.cfg, ctx.n = cfg, n
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        min_depth, max_depth, gamma = ctx.cfg
        n = ctx.n
        gt, invs = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("dro_sfm_b200: the supervised depth loss has no gradient w.r.t. the ground truth")
        B, _, H, W = gt.shape
        need = [ctx.needs_input_grad[3 + i] for i in range(n)]
        slab = torch.empty(n, *invs[0].shape, device=gt.device, dtype=torch.float32)
        gs = [slab[i] if need[i] else None for i in range(n)]
        g = L.f32c(g.reshape(1))
        with torch.cuda.device(gt.device):
            L.check(L.lib().drosfm_sup_depth_loss_bwd(L.ptr(g), L.ptr(gt), L.ptr_array(invs), n, min_depth, max_depth, gamma,
                                                      L.ptr_array(gs), B, H, W, L.stream()), "sup_depth_loss_bwd")
        return (None, None, None, *gs)


def sup_depth_loss(inv_depths, gt_inv_depth, min_depth, max_depth, gamma=0.85):
    """SupervisedDepthPoseLoss.calculate_loss (supervised_loss.py:244-277) for predictions at the GT resolution."""
    n = len(inv_depths)
    if not 1 <= n <= L.MAX_PREDS:
        raise ValueError("sup_depth_loss supports 1..{} predictions".format(L.MAX_PREDS))
    for d in inv_depths:
        if tuple(d.shape) != tuple(gt_inv_depth.shape):
            raise NotImplementedError("dro_sfm_b200: predictions must be at the ground-truth resolution")
    return _SupDepthLoss.apply(gt_inv_depth, (float(min_depth), float(max_depth), float(gamma)), n, *inv_depths)


# ------------------------------------------------------------------------------------------------
# evaluation path
# ------------------------------------------------------------------------------------------------
def post_process_inv_depth(inv_depth, inv_depth_flipped, method="mean"):
    """post_process_inv_depth (utils/depth.py:230-258), one launch.  No gradient (evaluation only)."""
    methods = {"mean": 0, "max": 1, "min": 2}
    if method not in methods:
        raise ValueError("Unknown post-process method {}".format(method))
    L.require_cuda(inv_depth, inv_depth_flipped)
    a, b = L.f32c(inv_depth.detach()), L.f32c(inv_depth_flipped.detach())
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError("inverse depth maps must be [B,C,H,W] tensors of the same shape")
    B, C, H, W = a.shape
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        L.check(L.lib().drosfm_post_process_inv_depth(L.ptr(a), L.ptr(b), L.ptr(out), B * C, H, W, methods[method], L.stream()),
                "post_process_inv_depth")
    return out


def depth_metrics(gt, pred, min_depth, max_depth, crop="", use_gt_scale=True):
    """compute_depth_metrics (utils/depth.py:261-340) -> float32 [9] on the device: abs_rel, sq_rel, rmse, rmse_log, a1,
    a2, a3, SILog, iabs_diff.  gt [B,1,H,W]; pred [B,1,h,w] (interpolated to the ground-truth resolution inside)."""
    L.require_cuda(gt, pred)
    g, p = L.f32c(gt.detach()), L.f32c(pred.detach())
    if g.dim() != 4 or p.dim() != 4 or g.shape[1] != 1 or p.shape[1] != 1 or g.shape[0] != p.shape[0]:
        raise ValueError("gt and pred must be [B,1,H,W] / [B,1,h,w]")
    B, _, H, W = g.shape
    code = {"garg": 1, "eigen_nyu": 2}.get(crop, 0)
    out = torch.empty(9, device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        ws = L.eval_workspace(g.device, B)
        L.check(L.lib().drosfm_depth_metrics(L.ptr(g), L.ptr(p), B, H, W, p.shape[2], p.shape[3], float(min_depth), float(max_depth),
                                             code, int(bool(use_gt_scale)), L.ptr(out), L.ptr(ws), L.stream()), "depth_metrics")
    return out


# ------------------------------------------------------------------------------------------------
# 8-bit pictures
# ------------------------------------------------------------------------------------------------
def images_u8_to_f32(src, out=None):
    """uint8 pictures -> float32 / 255 (ToTensor, datasets/augmentations.py:149-152) on the device, one launch.
    `out`: optional float32 tensor with the same number of elements (written in place)."""
    L.require_cuda(src)
    if src.dtype != torch.uint8:
        raise ValueError("images_u8_to_f32 expects a uint8 tensor")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or out.numel() != src.numel() or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 tensor with as many elements as src")
    with torch.cuda.device(src.device):
        L.check(L.lib().drosfm_images_u8_to_f32(L.ptr(src), L.ptr(out), src.numel(), L.stream()), "images_u8_to_f32")
    return out


# ------------------------------------------------------------------------------------------------
# convex up-sampling
# ------------------------------------------------------------------------------------------------
class _UpsampleDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, mask, ratio, disp_min, disp_range):
        L.require_cuda(depth, mask)
        depth, mask = L.f32c(depth), L.f32c(mask)
        N, _, H, W = depth.shape
        out = torch.empty(N, 1, ratio * H, ratio * W, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            L.check(L.lib().drosfm_upsample_depth_fwd(L.ptr(depth), L.ptr(mask), L.ptr(out), N, H, W, ratio, disp_min, disp_range,
                                                      L.stream()), "upsample_depth_fwd")
        ctx.save_for_backward(depth, mask)
        ctx.ratio, ctx.disp_range = ratio, disp_range
        return out

    @staticmethod
    def backward(ctx, g):
        depth, mask = ctx.saved_tensors
        g = L.f32c(g)
        N, _, H, W = depth.shape
        gd = torch.zeros_like(depth) if ctx.needs_input_grad[0] else None
        gm = torch.empty_like(mask) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(depth.device):
            L.check(L.lib().drosfm_upsample_depth_bwd(L.ptr(g), L.ptr(depth), L.ptr(mask), L.ptr(gd), L.ptr(gm), N, H, W,
                                                      ctx.ratio, ctx.disp_range, L.stream()), "upsample_depth_bwd")
        return gd, gm, None, None, None


def upsample_depth(depth, mask, ratio=8, disp_range=None):
    """DepthPoseNet.upsample_depth (DepthPoseNet.py:63-74): [N,1,H,W] x [N,9*ratio^2,H,W] -> [N,1,ratio*H,ratio*W].
    disp_range = (min_depth, max_depth): additionally applies disp_to_depth's scaling (layers.py:11-20, what
    DepthPoseNet.forward does to every up-sampled map) as the kernel's epilogue."""
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise AssertionError("depth must be [N,1,H,W]")
    if ratio != 8:
        raise NotImplementedError("dro_sfm_b200: upsample_depth supports ratio 8 (DepthPoseNet.feat_ratio)")
    if tuple(mask.shape) != (depth.shape[0], 9 * ratio * ratio, depth.shape[2], depth.shape[3]):
        raise ValueError("mask must be [N,{},H,W], got {}".format(9 * ratio * ratio, tuple(mask.shape)))
    dmin, drange = 0.0, 1.0
    if disp_range is not None:
        lo, hi = 1.0 / float(disp_range[1]), 1.0 / float(disp_range[0])
        dmin, drange = float(torch.tensor(lo, dtype=torch.float32)), float(torch.tensor(hi - lo, dtype=torch.float32))
    return _UpsampleDepth.apply(depth, mask, int(ratio), dmin, drange)
