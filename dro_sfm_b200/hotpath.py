"""One training step's worth of hot-path work, driven through the reference-facing operator surface.

This is what ``bench.py``, ``__graft_entry__.smoke()`` and the end-to-end tests execute: for a
workload of ``synthetic.WORKLOADS`` it issues exactly the calls the unchanged training loop makes
into the replaced subsystems:

  * ``DepthPoseNet.forward`` (DepthPoseNet.py:154-192): for each of the T GRU steps one
    ``depth_cost_calc`` (V views, gradients to inverse depth and features) and V ``get_cost_each``
    calls (gradients to the pose vector and features) -- 2*V*T cost evaluations forward and backward;
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
from .networks import depth_cost_calc, get_cost_each


def _bytes(t):
    return t.numel() * t.element_size()


class HotPathStep:
    """Static device buffers + the step function (eager or CUDA-graph replay)."""

    def __init__(self, wl, device, B=None, C=128, seed=1234, channels_last=False):
        self.wl, self.device, self.C = wl, torch.device(device), C
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
        self.image = h["image"].to(dev)
        self.context = [c.to(dev) for c in h["context"]]
        self.K = h["K"].to(dev)                                    # float64, as numpy collation delivers it
        self.fmap = self._feat(h["fmap"]).requires_grad_(True)
        self.frefs = [self._feat(f).requires_grad_(True) for f in h["fmaps_ref"]]
        self.inv_lr = [x.to(dev).requires_grad_(True) for x in h["inv_depth_lr"]]
        # the pose cost sees depth = inv2depth(scale(inv_depth)) built once per outer iteration by the caller
        self.depth_lr = [(1.0 / h["inv_depth_lr"][k * wl.seq_len].clamp(min=1e-6)).to(dev) for k in range(seq_iters)]
        self.pose_lr = [[p.to(dev).requires_grad_(True) for p in row] for row in h["pose_lr"]]
        self.inv_depths = [x.to(dev).requires_grad_(True) for x in h["inv_depths"]]
        self.poses = [[p.to(dev).requires_grad_(True) for p in row] for row in h["poses"]]
        if wl.supervised:
            self.gt_inv_depth = h["gt_inv_depth"].to(dev)
            self.gt_poses = [p.to(dev) for p in h["gt_poses"]]
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
        # host-resident inputs of a step (pinned) and their device destinations, for the end-to-end timing
        self._h2d = []
        pairs = [(h["image"], self.image), (h["K"], self.K), (h["fmap"], self.fmap)]
        pairs += list(zip(h["context"], self.context)) + list(zip(h["fmaps_ref"], self.frefs))
        pairs += list(zip(h["inv_depth_lr"], self.inv_lr)) + list(zip(h["inv_depths"], self.inv_depths))
        pairs += [(p, q) for r, s in zip(h["pose_lr"], self.pose_lr) for p, q in zip(r, s)]
        pairs += [(p, q) for r, s in zip(h["poses"], self.poses) for p, q in zip(r, s)]
        if wl.supervised:
            pairs += [(h["gt_inv_depth"], self.gt_inv_depth)] + list(zip(h["gt_poses"], self.gt_poses))
        for src, dst in pairs:
            self._h2d.append((src.contiguous().pin_memory() if self.device.type == "cuda" else src, dst))
        self.h2d_bytes = sum(_bytes(s) for s, _ in self._h2d)

    def leaves(self):
        out = [self.fmap] + self.frefs + self.inv_lr + [p for r in self.pose_lr for p in r]
        return out + self.inv_depths + [p for r in self.poses for p in r]

    def upload(self):
        """Host -> device copy of every forward input of the step (pinned memory, current stream)."""
        with torch.no_grad():
            for src, dst in self._h2d:
                dst.copy_(src, non_blocking=True)      # channels_last destinations are handled by copy_

    # -- the step --------------------------------------------------------------------------------
    def forward_backward(self):
        wl = self.wl
        costs = []
        for t in range(wl.T):
            poses_t = [p.detach() for p in self.pose_lr[t]]          # DepthPoseNet.py:156
            costs.append(depth_cost_calc(self.inv_lr[t], self.fmap, self.frefs, poses_t, self.K, self.K, 1.0 / 8))
            depth = self.depth_lr[t // wl.seq_len]
            for v in range(wl.V):
                costs.append(get_cost_each(self.pose_lr[t][v], self.fmap, self.frefs[v], depth, self.K, self.K, 1.0 / 8))
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
        with torch.cuda.graph(self.graph):
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
        """every distinct input read once, every output written once per call."""
        wl, B, C = self.wl, self.B, self.C
        p = (wl.H // 8) * (wl.W // 8) * B
        P = wl.H * wl.W * B
        V, T, n = wl.V, wl.T, wl.n
        out = {
            "feat_cost_fwd_v1": (C * 4 * 3 + 4) * p, "feat_cost_bwd_v1": (C * 4 * 5 + 8) * p,
            "feat_cost_fwd_vN": (C * 4 * (V + 2) + 4) * p, "feat_cost_bwd_vN": (C * 4 * (2 * V + 3) + 8) * p,
            "photometric_fwd": (16 + 12 * V) * P * n, "photometric_bwd": (20 + 12 * V) * P * n,
            "automask_fwd": (12 + 12 * V + 4) * P,
            "smoothness_fwd": (12 + 8 * n) * P, "smoothness_bwd": (12 + 8 * n) * P,
            "reproj_loss_fwd": 4 * P, "reproj_loss_bwd": 4 * P,
        }
        calls = {"feat_cost_fwd_v1": V * T, "feat_cost_bwd_v1": V * T, "feat_cost_fwd_vN": T, "feat_cost_bwd_vN": T}
        if wl.supervised:
            total = sum(out[k] * calls[k] for k in calls) + out["reproj_loss_fwd"] + out["reproj_loss_bwd"]
        else:
            total = sum(out[k] * calls[k] for k in calls) + sum(out[k] for k in (
                "photometric_fwd", "photometric_bwd", "automask_fwd", "smoothness_fwd", "smoothness_bwd"))
        out["step_total"] = total
        return out
