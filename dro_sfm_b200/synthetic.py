"""Seeded synthetic KITTI-/ScanNet-shaped inputs for the warping hot path.

Everything is generated on the CPU with an explicit ``torch.Generator`` so that the same tensors
can be fed to the CUDA kernels, to the CPU oracle and (in the build container) to the reference.
Shapes and value ranges follow the reference's batch schema (kitti_dataset.py:348-406,
scannet_dataset.py:189-231) and its configs (configs/train_kitti_mf_selfsup.yaml etc.).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration, reduced to what the hot path sees."""
    name: str
    H: int
    W: int
    B: int            # per-GPU batch (YAML batch_size)
    V: int            # source views (forward_context + back_context)
    T: int            # total GRU steps (DepthPoseNet version 'it<T>')
    seq_len: int
    n: int            # predictions seen by the loss
    min_depth: float
    max_depth: float
    supervised: bool
    dataset: str      # 'kitti' | 'scannet'


WORKLOADS = {
    # configs/overfit_kitti_mf_gt.yaml  (it12-h-out, 192x640, V=2) -- BASELINE config 0, B=1 on CPU
    "overfit_kitti_mf_gt": Workload("overfit_kitti_mf_gt", 192, 640, 1, 2, 12, 4, 4, 0.2, 80.0, True, "kitti"),
    # configs/train_kitti_mf_selfsup.yaml (it8-seq4-inter-out, 320x960, B=2, V=2) -- BASELINE config 1
    "train_kitti_mf_selfsup": Workload("train_kitti_mf_selfsup", 320, 960, 2, 2, 8, 4, 9, 0.5, 80.0, False, "kitti"),
    # same network on the "KITTI-shaped" 192x640 frame named by BASELINE.json
    "train_kitti_mf_selfsup_192x640": Workload("train_kitti_mf_selfsup_192x640", 192, 640, 2, 2, 8, 4, 9, 0.5, 80.0, False, "kitti"),
    # configs/train_scannet_mf_gt_view3.yaml (it12-h-out, 240x320, B=8, V=2)
    "train_scannet_mf_gt_view3": Workload("train_scannet_mf_gt_view3", 240, 320, 8, 2, 12, 4, 4, 0.2, 10.0, True, "scannet"),
    # configs/train_scannet_mf_selfsup_view5.yaml (it12-h-out, 240x320, B=4, V=4)
    "train_scannet_mf_selfsup_view5": Workload("train_scannet_mf_selfsup_view5", 240, 320, 4, 4, 12, 4, 4, 0.2, 10.0, False, "scannet"),
}


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def intrinsics(dataset, B, H, W, flip=False, dtype=torch.float64):
    """[B,3,3] pinhole intrinsics resized to HxW (augmentations.py:96-99); optional LR-flip
    (utils/image.py:79-80: fx -> -fx, cx -> W - cx)."""
    if dataset == "kitti":
        fx, fy, cx, cy, W0, H0 = 721.5377, 721.5377, 609.5593, 172.854, 1242.0, 375.0
    else:
        fx, fy, cx, cy, W0, H0 = 1169.62, 1167.11, 646.30, 489.93, 1296.0, 968.0
    K = torch.tensor([[fx * W / W0, 0.0, cx * W / W0],
                      [0.0, fy * H / H0, cy * H / H0],
                      [0.0, 0.0, 1.0]], dtype=dtype)
    if flip:
        K[0, 0] = -K[0, 0]
        K[0, 2] = W - K[0, 2]
    return K.unsqueeze(0).repeat(B, 1, 1).contiguous()


def smooth_noise(g, B, C, H, W, cell=8):
    """Band-limited noise in [0,1): bilinear up-sampling of a coarse uniform field."""
    h, w = max(2, math.ceil(H / cell)), max(2, math.ceil(W / cell))
    coarse = torch.rand(B, C, h, w, generator=g)
    return F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True)


def images(g, B, H, W):
    """[B,3,H,W] in [0,1): smooth field + 5% white noise (non-degenerate SSIM statistics)."""
    return (0.95 * smooth_noise(g, B, 3, H, W) + 0.05 * torch.rand(B, 3, H, W, generator=g)).contiguous()


def quantise8(x):
    """The same image as a decoded 8-bit picture: k / 255 (what ToTensor delivers, datasets/augmentations.py:149-152)."""
    return torch.round(x.clamp(0, 1) * 255.0).div(255.0)


def inv_depth(g, B, H, W, min_depth, max_depth, frac_nonpos=0.0):
    """[B,1,H,W] inverse depth in (1/max_depth, 1/min_depth) (layers.py:11-20 range).
    ``frac_nonpos`` injects values <= 0 to exercise the inv2depth mask (utils/depth.py:119-121)."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    s = torch.sigmoid(4.0 * (smooth_noise(g, B, 1, H, W) - 0.5) - 2.0)
    d = lo + (hi - lo) * s
    if frac_nonpos > 0:
        hole = torch.rand(B, 1, H, W, generator=g) < frac_nonpos
        d = torch.where(hole, -torch.rand(B, 1, H, W, generator=g) * (torch.rand(B, 1, H, W, generator=g) > 0.5), d)
    return d.contiguous()


def pose_vec(g, B, dataset, direction=1.0):
    """[B,6] = (tx,ty,tz,rx,ry,rz): small ego-motion around +-1 m forward (KITTI) or a hand-held
    wobble (ScanNet)."""
    if dataset == "kitti":
        t = torch.randn(B, 3, generator=g) * torch.tensor([0.05, 0.02, 0.5]) + torch.tensor([0.0, 0.0, direction])
        r = torch.randn(B, 3, generator=g) * 0.01
    else:
        t = torch.randn(B, 3, generator=g) * 0.1
        r = torch.randn(B, 3, generator=g) * 0.05
    return torch.cat([t, r], dim=1).contiguous()


def features(g, B, C, h, w):
    """[B,C,h,w] ~ N(0,1) feature maps (the encoder output the cost kernels consume)."""
    return torch.randn(B, C, h, w, generator=g)


def hot_path_batch(wl, seed=1234, C=128, B=None):
    """All hot-path inputs of one training step of workload ``wl`` (CPU tensors).

    Returns a dict: image [B,3,H,W]; context (V x [B,3,H,W]); K [B,3,3] float64; fmap and
    fmaps_ref at 1/8 resolution; per GRU step low-res inverse depths / pose vectors; the n
    full-resolution inverse-depth predictions and V x n pose vectors seen by the loss; for
    supervised workloads GT inverse depth (30% holes on KITTI) and V GT pose vectors.
    """
    B = wl.B if B is None else B
    g = gen(seed)
    h, w = wl.H // 8, wl.W // 8
    out = {
        # 8-bit pictures (k / 255): the end-to-end path ships them as uint8, the way a dataset stores them
        "image": quantise8(images(g, B, wl.H, wl.W)),
        "context": [quantise8(images(g, B, wl.H, wl.W)) for _ in range(wl.V)],
        "K": intrinsics(wl.dataset, B, wl.H, wl.W),
        "fmap": features(g, B, C, h, w),
        "fmaps_ref": [features(g, B, C, h, w) for _ in range(wl.V)],
        "inv_depth_lr": [inv_depth(g, B, h, w, wl.min_depth, wl.max_depth) for _ in range(wl.T)],
        "pose_lr": [[pose_vec(g, B, wl.dataset, 1.0 if v % 2 == 0 else -1.0) for v in range(wl.V)]
                    for _ in range(wl.T)],
        "inv_depths": [inv_depth(g, B, wl.H, wl.W, wl.min_depth, wl.max_depth) for _ in range(wl.n)],
        "poses": [[pose_vec(g, B, wl.dataset, 1.0 if v % 2 == 0 else -1.0) for _ in range(wl.n)]
                  for v in range(wl.V)],
    }
    if wl.supervised:
        gt = inv_depth(g, B, wl.H, wl.W, wl.min_depth, wl.max_depth)
        if wl.dataset == "kitti":
            gt = gt * (torch.rand(B, 1, wl.H, wl.W, generator=g) > 0.3)
        out["gt_inv_depth"] = gt.contiguous()
        out["gt_poses"] = [pose_vec(g, B, wl.dataset, 1.0 if v % 2 == 0 else -1.0) for v in range(wl.V)]
    return out
