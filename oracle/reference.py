"""Harness around the UNMODIFIED reference (xyang9527/dro-sfm) -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it means staging its package where it can be imported:
``/root/reference`` in the build container, or the git-ignored copy ``baseline/_ref`` that
``oracle/stage_reference.py`` makes there (it travels to the GPU box with the repo snapshot; no reference source is
ever committed).  Nothing under ``dro_sfm_b200/`` imports this module.  Users:

  * ``tests/golden/make_golden.py``       fixtures from the reference on the CPU,
  * ``tests/test_reference_e2e_gpu.py``   the reference's DepthPoseNet + loss on the B200, stock vs ``patch.install()``,
  * ``bench.py``                          ``--impl reference`` / ``cpu_baseline`` (kind "reference") and the
                                          ``gpu_aten_reference`` comparator.

Harness-side shims (the reference's files are not touched):
  * stub modules for yacs / matplotlib.cm / numpy.lib.type_check / termcolor, imported by the reference but never used
    on the hot path;
  * ``Tensor.get_device`` returns the device object on CPU (``warp_ref_image``,
    multiview_photometric_loss_mf.py:156, feeds it to ``.to()``);
  * ``torchvision.models.resnet.model_urls`` / ``model_zoo.load_url`` -> a seeded random ResNet-18 state dict
    (extractor.py:56-65 downloads ImageNet weights; there is no network).
"""
import logging
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
STAGED = os.path.join(_REPO, "baseline", "_ref")


def root():
    """Directory that holds the reference's ``dro_sfm`` package, or None."""
    for cand in (os.environ.get("DROSFM_REFERENCE"), "/root/reference", STAGED):
        if cand and os.path.isdir(os.path.join(cand, "dro_sfm")):
            return cand
    return None


def available():
    return root() is not None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _shims():
    import numpy as np
    import torch

    logging.disable(logging.WARNING)  # every hot-path ctor logs at WARNING

    class _CfgNode(dict):
        pass

    yacs = _stub("yacs")
    yacs.config = _stub("yacs.config", CfgNode=_CfgNode)
    try:
        import matplotlib.cm  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.cm = _stub("matplotlib.cm", get_cmap=lambda *a, **k: None)
    try:
        from numpy.lib.type_check import imag  # noqa: F401
    except Exception:
        _stub("numpy.lib.type_check", imag=np.imag)
    try:
        import termcolor  # noqa: F401
    except Exception:
        _stub("termcolor", colored=lambda s, *a, **k: s)

    if not getattr(torch.Tensor.get_device, "_drosfm_patched", False):
        orig = torch.Tensor.get_device

        def get_device(t):
            return orig(t) if t.is_cuda else t.device
        get_device._drosfm_patched = True
        torch.Tensor.get_device = get_device


_loaded = None


def load():
    """Namespace with the reference's hot-path symbols (imports the reference on first use)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    ref_root = root()
    if ref_root is None:
        raise RuntimeError("reference tree not found (tried $DROSFM_REFERENCE, /root/reference, %s)" % STAGED)
    _shims()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)

    from dro_sfm.geometry.camera import Camera
    from dro_sfm.geometry.pose import Pose
    from dro_sfm.geometry.camera_utils import view_synthesis, scale_intrinsics
    from dro_sfm.utils.depth import inv2depth, calc_smoothness
    from dro_sfm.utils import depth as _depth_utils
    from dro_sfm.losses.multiview_photometric_loss_mf import MultiViewPhotometricDecayLoss, SSIM
    from dro_sfm.losses.supervised_loss import SupervisedDepthPoseLoss
    from dro_sfm.networks.depth_pose.DepthPoseNet import DepthPoseNet

    net = object.__new__(DepthPoseNet)  # get_cost_each / depth_cost_calc use no module state
    # the stock functions are captured HERE, before any patch.install(): callers that want the stock behaviour after
    # the reference tree has been patched keep using these
    stock = dict(get_cost_each=DepthPoseNet.get_cost_each, depth_cost_calc=DepthPoseNet.depth_cost_calc,
                 upsample_depth=DepthPoseNet.upsample_depth, post_process_inv_depth=_depth_utils.post_process_inv_depth,
                 compute_depth_metrics=_depth_utils.compute_depth_metrics)
    _loaded = types.SimpleNamespace(
        root=ref_root, Camera=Camera, Pose=Pose, view_synthesis=view_synthesis, scale_intrinsics=scale_intrinsics,
        inv2depth=inv2depth, calc_smoothness=calc_smoothness, SSIM=SSIM,
        MultiViewPhotometricDecayLoss=MultiViewPhotometricDecayLoss,
        SupervisedDepthPoseLoss=SupervisedDepthPoseLoss, DepthPoseNet=DepthPoseNet,
        get_cost_each=lambda *a, **k: stock["get_cost_each"](net, *a, **k),
        depth_cost_calc=lambda *a, **k: stock["depth_cost_calc"](net, *a, **k),
        upsample_depth=lambda *a, **k: stock["upsample_depth"](net, *a, **k),
        post_process_inv_depth=stock["post_process_inv_depth"], compute_depth_metrics=stock["compute_depth_metrics"],
    )
    return _loaded


def build_depth_pose_net(version, min_depth, max_depth, seed=0):
    """The reference's DepthPoseNet with seeded random weights (no ImageNet download)."""
    import torch
    import torchvision.models as models
    from torch.utils import model_zoo
    ref = load()
    if not hasattr(models.resnet, "model_urls"):
        models.resnet.model_urls = {"resnet18": "seeded://resnet18"}
    state = {}

    def load_url(url, *a, **k):
        if "sd" not in state:
            g = torch.random.get_rng_state()
            torch.manual_seed(seed + 1000)
            state["sd"] = models.resnet18().state_dict()
            torch.random.set_rng_state(g)
        return state["sd"]
    orig = model_zoo.load_url
    model_zoo.load_url = load_url
    try:
        torch.manual_seed(seed)
        net = ref.DepthPoseNet(version=version, min_depth=min_depth, max_depth=max_depth)
    finally:
        model_zoo.load_url = orig
    return net


def hot_path_step(wl, batch, device="cpu", dtype=None, return_grads=False):
    """The hot-path work of one training step (what bench.py times) executed by the REFERENCE's own functions:
    2*V*T ``DepthPoseNet.get_cost_each`` / ``depth_cost_calc`` evaluations fwd+bwd and
    ``MultiViewPhotometricDecayLoss`` / ``SupervisedDepthPoseLoss`` fwd+bwd, on `device`.  Same inputs, call order and
    upstream gradients as ``bench.cpu_step`` / ``HotPathStep``.  Returns the loss (and every leaf gradient)."""
    import torch
    ref = load()
    dev = torch.device(device)
    dt = dtype or torch.float32

    def c(x):
        return x.to(device=dev, dtype=dt)
    K = batch["K"].to(dev)                                  # float64, as numpy collation delivers it (callers do K.float())
    fmap = c(batch["fmap"]).clone().requires_grad_(True)
    frefs = [c(f).clone().requires_grad_(True) for f in batch["fmaps_ref"]]
    B, C, h, w = fmap.shape
    g = torch.Generator().manual_seed(99)
    n_cost = wl.T * (1 + wl.V)
    gouts = [c(torch.randn(B, C, h, w, generator=g)) for _ in range(n_cost)] if return_grads else \
        [c(torch.randn(B, C, h, w, generator=g))] * n_cost
    outs, inv_lr, pose_lr = [], [], []
    for t in range(wl.T):
        inv = c(batch["inv_depth_lr"][t]).clone().requires_grad_(True)
        inv_lr.append(inv)
        outs.append(ref.depth_cost_calc(inv, fmap, frefs, [c(p) for p in batch["pose_lr"][t]], K, K, 0.125))
        depth = ref.inv2depth(c(batch["inv_depth_lr"][(t // wl.seq_len) * wl.seq_len]))
        for v in range(wl.V):
            pose = c(batch["pose_lr"][t][v]).clone().requires_grad_(True)
            pose_lr.append(pose)
            outs.append(ref.get_cost_each(pose, fmap, frefs[v], depth, K, K, 0.125))
    invs = [c(x).clone().requires_grad_(True) for x in batch["inv_depths"]]
    pvec = [[c(p).clone().requires_grad_(True) for p in row] for row in batch["poses"]]
    poses = [[ref.Pose.from_vec(p, "euler") for p in row] for row in pvec]
    image, context = c(batch["image"]), [c(x) for x in batch["context"]]
    if wl.supervised:
        mod = ref.SupervisedDepthPoseLoss(min_depth=wl.min_depth, max_depth=wl.max_depth)
        gts = [ref.Pose.from_vec(c(p), "euler").mat for p in batch["gt_poses"]]
        out = mod(image, context, invs, c(batch["gt_inv_depth"]), gts, K, K, poses)
    else:
        mod = ref.MultiViewPhotometricDecayLoss(ssim_loss_weight=0.85, C1=1e-4, C2=9e-4, photometric_reduce_op="min",
                                                clip_loss=0.0, padding_mode="zeros", automask_loss=True,
                                                smooth_loss_weight=0.001)
        out = mod(image, context, invs, K, K, poses)
    loss = out["loss"].sum()
    torch.autograd.backward([loss] + outs, [torch.ones((), device=dev, dtype=loss.dtype)] + gouts)
    if not return_grads:
        return float(loss.detach())
    leaves = [fmap] + frefs + inv_lr + pose_lr + invs + [p for row in pvec for p in row]
    return float(loss.detach()), [x.grad for x in leaves]
