"""One training step's worth of hot-path work, driven through the reference-facing operator surface.

This is what ``bench.py``, ``__graft_entry__.smoke()`` and the end-to-end tests execute: for a
workload of ``synthetic.WORKLOADS`` it issues exactly the calls the unchanged training loop makes
into the replaced subsystems:

  * ``DepthPoseNet.forward`` (DepthPoseNet.py:154-192): for each of the T GRU steps one
    ``depth_cost_calc`` (V views, gradients to inverse depth and features) and V ``get_cost_each``
    calls (gradients to the pose vector and features) -- 2*V*T cost evaluations forward and backward.
    With the lock-step schedule that ``patch.install()`` grafts onto DepthPoseNet (networks/lockstep.py; default) the
    1 + V evaluations of a GRU step are one ``cost_batch`` call; ``lockstep=False`` issues them one by one as the
    reference's own schedule does;
  * ``SelfSupModelMF`` / ``SupModelMF``: ``Pose.from_vec`` for every (view, prediction), then
    ``MultiViewPhotometricDecayLoss.forward`` or ``SupervisedDepthPoseLoss.forward`` and their backward.

The upstream gradients of the cost maps (what the update blocks' convolutions would send back) are
fixed random tensors.  A *frame* is one batch sample through all of this.
"""
import torch

from . import _lib as L
from . import synthetic as syn
from .geometry import Pose
from .losses import MultiViewPhotometricDecayLoss, SupervisedDepthPoseLoss
from .networks import cost_batch, depth_cost_calc, get_cost_each


def _bytes(t):
    return t.numel() * t.element_size()


class HotPathStep:
    """Static device buffers + the step function (eager or CUDA-graph replay)."""

    def __init__(self, wl, device, B=None, C=128, seed=1234, channels_last=False, lockstep=True):
        self.wl, self.device, self.C = wl, torch.device(device), C
        self.lockstep = lockstep
        self.B = wl.B if B is None else B
        self.channels_last = channels_last
        self.host = syn.hot_path_batch(wl, seed=seed, C=C, B=self.B)
        self.graph = None
        self._alloc()
        if wl.supervised:
            self.loss_mod = SupervisedDepthPoseLoss(min_depth=wl.min_depth, max_depth=wl.max_depth)
        else:
            # configs/default_config.py:90-111 + the YAML's automask_loss / photometric_reduce_op
            self.loss_mod = MultiViewPhotometricDecayLoss(
                ssim_loss_weight=0.85, C1=1e-4, C2=9e-4, photometric_reduce_op='min', clip_loss=0.0,
                padding_mode='zeros', automask_loss=True, smooth_loss_weight=0.001)

    # -- buffers ---------------------------------------------------------------------------------
    def _feat(self, t):
        t = t.to(self.device)
        return t.contiguous(memory_format=torch.channels_last) if self.channels_last else t

    def _alloc(self):
        wl, h, dev = self.wl, self.host, self.device
        seq_iters = wl.T // wl.seq_len
        # Every forward input of a step lives in ONE flat device buffer: pictures first, GT tensors last (the two
        # regions the end-to-end path refreshes from the host every step, see below).
        host_list = [h["image"], *h["context"], h["fmap"], *h["fmaps_ref"], *h["inv_depth_lr"], *h["inv_depths"]]
        host_list += [p for r in h["pose_lr"] for p in r] + [p for r in h["poses"] for p in r]
        if wl.supervised:
            host_list += [h["gt_inv_depth"], *h["gt_poses"]]
        host_list = [t.contiguous().float() for t in host_list]
        sizes = [t.numel() for t in host_list]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 3) // 4 * 4                       # keep every tensor 16-byte aligned
        k64 = h["K"].contiguous()                           # float64 intrinsics travel separately (tiny)
        self.host_flat = torch.zeros(total, dtype=torch.float32)
        for t, o in zip(host_list, offs):
            self.host_flat[o:o + t.numel()] = t.reshape(-1)
        # What a data loader delivers per step (the end-to-end leg of bench.py moves exactly this, every step):
        #   * the target and source pictures as uint8 (they are 8-bit images, synthetic.quantise8; the device turns them
        #     into the float32 tensors the loss reads with drosfm_images_u8_to_f32 == ToTensor's x / 255),
        #   * the float64 intrinsics, and for supervised workloads the GT inverse depth and GT poses (float32).
        # Feature maps, inverse depths and pose vectors are produced ON the device by the networks in a training step
        # (DepthPoseNet.py:113-205); they stay resident and are not part of the host batch.
        self.n_img = (1 + wl.V) * self.B * 3 * wl.H * wl.W                       # image + contexts lead the flat buffer
        assert offs[1 + wl.V] == self.n_img
        self.host_u8 = torch.round(self.host_flat[:self.n_img] * 255.0).to(torch.uint8)
        self.extra_lo = offs[len(host_list) - (1 + wl.V)] if wl.supervised else total     # GT tensors trail the flat buffer
        self.host_extra = self.host_flat[self.extra_lo:].clone()
        if dev.type == "cuda":
            self.host_u8, self.host_extra, self.host_K = self.host_u8.pin_memory(), self.host_extra.pin_memory(), k64.pin_memory()
        else:
            self.host_K = k64
        self.flat = self.host_flat.to(dev)
        self.staging_u8 = torch.empty(self.host_u8.shape, dtype=torch.uint8, device=dev)
        self.staging_extra = torch.empty(self.host_extra.shape, dtype=torch.float32, device=dev)
        self.K = k64.to(dev)                                 # float64, as numpy collation delivers it
        self.K_staging = torch.empty_like(self.K)
        with torch.no_grad():
            views = [self.flat[o:o + t.numel()].view(t.shape) for t, o in zip(host_list, offs)]
        it = iter(views)
        V, T, n = wl.V, wl.T, wl.n

        def take(k):
            return [next(it) for _ in range(k)]

        self.image = next(it)
        self.context = take(V)
        # the feature maps of [target, source_1..V] are ONE tensor, as the encoder delivers them (DepthPoseNet.py:113-118:
        # fnet(torch.cat([target] + refs)) followed by torch.split); they are adjacent in the flat buffer
        parts = [next(it)] + take(V)
        o0 = offs[1 + V]
        with torch.no_grad():
            stacked = self.flat[o0:o0 + sum(p.numel() for p in parts)].view((1 + V) * self.B, *parts[0].shape[1:])
        assert stacked.data_ptr() == parts[0].data_ptr() and parts[0].shape[0] == self.B
        if self.channels_last:                               # a network that already runs channels_last
            stacked = stacked.detach().contiguous(memory_format=torch.channels_last)
        self.fmaps_all = stacked.requires_grad_(True)
        self.inv_lr = [x.requires_grad_(True) for x in take(T)]
        self.inv_depths = [x.requires_grad_(True) for x in take(n)]
        flat_pose_lr = take(T * V)
        self.pose_lr = [[flat_pose_lr[t * V + v].requires_grad_(True) for v in range(V)] for t in range(T)]
        flat_poses = take(V * n)
        self.poses = [[flat_poses[v * n + i].requires_grad_(True) for i in range(n)] for v in range(V)]
        if wl.supervised:
            self.gt_inv_depth = next(it)
            self.gt_poses = take(V)
        # the pose cost sees depth = inv2depth(scale(inv_depth)) built once per outer iteration by the caller
        self.depth_lr = [(1.0 / h["inv_depth_lr"][k * wl.seq_len].clamp(min=1e-6)).to(dev) for k in range(seq_iters)]
        g = syn.gen(99)
        hh, ww = wl.H // 8, wl.W // 8
        n_cost = wl.T * (1 + wl.V)
        # the consumer of the cost map (a 1x1 convolution) returns its gradient in the layout of the cost map
        from .networks import cost as _cost_mod
        cl = self.channels_last or _cost_mod._LAYOUT == "nhwc"
        self.g_costs = [torch.randn(self.B, self.C, hh, ww, generator=g).to(dev) for _ in range(n_cost)]
        if cl:
            self.g_costs = [t.contiguous(memory_format=torch.channels_last) for t in self.g_costs]
        self.one = torch.ones(1, device=dev)
        self.h2d_bytes = self.host_u8.numel() + self.host_extra.numel() * 4 + self.host_K.numel() * 8

    def leaves(self):
        out = [self.fmaps_all] + self.inv_lr + [p for r in self.pose_lr for p in r]
        return out + self.inv_depths + [p for r in self.poses for p in r]

    def grads(self):
        """Gradients in the order bench.cpu_step returns them: fmap, fmaps_ref[0..V-1], inv_depth_lr, pose_lr, inv_depths,
        poses (the stacked feature-map gradient split per map)."""
        ls = self.leaves()
        return list(torch.split(ls[0].grad, self.B, dim=0)) + [t.grad for t in ls[1:]]

    def feature_maps(self):
        """(fmap, fmaps_ref) of this step: batch slices of the stacked tensor -- converted to channels_last once when the
        cost kernels want that layout (networks/cost.py: split_feature_maps)."""
        from .networks import cost as _cost_mod
        pieces = _cost_mod.split_feature_maps(self.fmaps_all, [self.B] * (1 + self.wl.V))
        return pieces[0], list(pieces[1:])

    def prefetch(self, stream):
        """Host -> device copy of the NEXT step's batch (uint8 pictures, intrinsics, GT tensors; pinned memory) into the
        staging buffers on `stream` (overlaps the running step); returns the event that marks its completion."""
        with torch.no_grad(), torch.cuda.stream(stream):
            self.staging_u8.copy_(self.host_u8, non_blocking=True)
            if self.staging_extra.numel():
                self.staging_extra.copy_(self.host_extra, non_blocking=True)
            self.K_staging.copy_(self.host_K, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        return ev

    def commit_staging(self):
        """The staged batch becomes the tensors the step reads (current stream): the pictures are converted to float32
        by one kernel, the small tensors are copied device-to-device."""
        from . import ops
        with torch.no_grad():
            ops.images_u8_to_f32(self.staging_u8, out=self.flat[:self.n_img])
            if self.staging_extra.numel():
                self.flat[self.extra_lo:].copy_(self.staging_extra, non_blocking=True)
            self.K.copy_(self.K_staging, non_blocking=True)

    # -- the step --------------------------------------------------------------------------------
    def forward_backward(self):
        wl = self.wl
        costs = []
        fmap, frefs = self.feature_maps()
        for t in range(wl.T):
            poses_t = [p.detach() for p in self.pose_lr[t]]          # DepthPoseNet.py:156
            depth = self.depth_lr[t // wl.seq_len]
            if self.lockstep:
                jobs = [(self.inv_lr[t], fmap, frefs, poses_t, True)]
                jobs += [(depth, fmap, [frefs[v]], [self.pose_lr[t][v]], False) for v in range(wl.V)]
                costs += cost_batch(jobs, self.K, self.K, 1.0 / 8)
                continue
            costs.append(depth_cost_calc(self.inv_lr[t], fmap, frefs, poses_t, self.K, self.K, 1.0 / 8))
            for v in range(wl.V):
                costs.append(get_cost_each(self.pose_lr[t][v], fmap, frefs[v], depth, self.K, self.K, 1.0 / 8))
        poses = [[Pose.from_vec(p, 'euler') for p in row] for row in self.poses]   # SfmModelMF.py:169-182
        if wl.supervised:
            out = self.loss_mod(self.image, self.context, self.inv_depths, self.gt_inv_depth,
                                [Pose.from_vec(p, 'euler').mat for p in self.gt_poses], self.K, self.K, poses)
        else:
            out = self.loss_mod(self.image, self.context, self.inv_depths, self.K, self.K, poses)
        loss = out['loss']
        torch.autograd.backward([loss] + costs, [self.one] + self.g_costs)
        return loss

    def zero_grads(self):
        for t in self.leaves():
            t.grad = None

    def capture(self, warmup=3):
        """Capture forward_backward into a CUDA graph (side-stream warm-up first, as torch requires)."""
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.zero_grads()
                self.forward_backward()
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.zero_grads()
        before = L.lib().drosfm_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the warm-up stream: autograd ties every leaf's AccumulateGrad node to the stream of its
        # first use, and a different capture stream would put cross-stream joins into the graph
        with torch.cuda.graph(self.graph, stream=s):
            self.loss = self.forward_backward()
        self.launches_per_step = int(L.lib().drosfm_launch_count() - before)
        return self.graph

    def step(self):
        """One step; returns the loss tensor (device)."""
        if self.graph is not None:
            self.graph.replay()
            return self.loss
        self.zero_grads()
        return self.forward_backward()

    # -- bookkeeping for the roofline ---------------------------------------------------------------
    def algorithmic_bytes(self):
        """Algorithmic bytes per launch unit, SURVEY.md section 8(d): every distinct input read once and every output
        written once per operator call, fp32.  Keys with a SURVEY figure:
          feat_cost_*            (V+2)*4C+4 fwd, (2V+3)*4C+8 bwd per feature pixel; a batched launch is the sum of its jobs;
          photometric_loss_fwd   (16+12V) B per pixel and prediction -- the loss forward as ONE fused operator (target,
                                 depth, V sources; auto-mask and min included), however many kernels implement it;
          photometric_loss_bwd   (20+12V) B per pixel and prediction;
          smoothness_fwd / _bwd  the operands of the pass over all n predictions: image once + n maps (+ n gradients);
          reproj_loss_*          4 B per pixel.
        Keys under "stage_operands" are the operand bytes of the individual stages of the staged photometric path
        (warped copy, g_warped: traffic of THIS design, no SURVEY figure) -- reported per call for orientation only."""
        wl, B, C = self.wl, self.B, self.C
        p = (wl.H // 8) * (wl.W // 8) * B
        P = wl.H * wl.W * B
        V, T, n = wl.V, wl.T, wl.n
        out = {
            "feat_cost_fwd_v1": (C * 4 * 3 + 4) * p, "feat_cost_bwd_v1": (C * 4 * 5 + 8) * p,
            "feat_cost_fwd_vN": (C * 4 * (V + 2) + 4) * p, "feat_cost_bwd_vN": (C * 4 * (2 * V + 3) + 8) * p,
            "photometric_loss_fwd": (16 + 12 * V) * P * n, "photometric_loss_bwd": (20 + 12 * V) * P * n,
            "smoothness_fwd": (12 + 4 * n) * P, "smoothness_bwd": (12 + 8 * n) * P,
            "reproj_loss_fwd": 4 * P, "reproj_loss_bwd": 4 * P,
        }
        out["feat_cost_batch_fwd"] = out["feat_cost_fwd_vN"] + V * out["feat_cost_fwd_v1"]
        out["feat_cost_batch_bwd"] = out["feat_cost_bwd_vN"] + V * out["feat_cost_bwd_v1"]
        out["stage_operands"] = {
            "warp_sources_fwd": (4 + 24 * V) * P * n, "photometric_fwd": (17 + 12 * V) * P * n,
            "photometric_bwd": (13 + 24 * V) * P * n, "warp_sources_bwd": (8 + 24 * V) * P * n,
            "automask_fwd": (12 + 12 * V + 4) * P,
        }
        cost = T * (out["feat_cost_fwd_vN"] + out["feat_cost_bwd_vN"]) + V * T * (out["feat_cost_fwd_v1"] + out["feat_cost_bwd_v1"])
        if wl.supervised:
            total = cost + out["reproj_loss_fwd"] + out["reproj_loss_bwd"]
        else:
            # SURVEY 8(d) per prediction: photometric 16+12V / 20+12V, smoothness 16 / 20 bytes per pixel
            total = cost + out["photometric_loss_fwd"] + out["photometric_loss_bwd"] + 36 * P * n
        out["step_total"] = total
        return out
