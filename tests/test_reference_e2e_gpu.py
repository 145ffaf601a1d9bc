"""The UNMODIFIED reference model on top of the B200 kernels, on the GPU.

`SelfSupModelMF` / `SupModelMF` (dro_sfm/models/*.py) with the reference's own `DepthPoseNet`
(`it8-seq4-inter-out`, DepthPoseNet.py:107-205; seeded random weights) run one training step -- forward, loss, backward
-- on cuda:0 twice: stock (every op is ATen), then after `dro_sfm_b200.patch.install()` (cost calls, convex
up-sampling, Pose.from_vec and the loss go through libdrosfm_b200.so; encoders, GRUs and heads are untouched).  The
network parameters are shared, so predictions, loss and parameter gradients must agree up to the rounding the
recurrent network amplifies.  Skipped when the staged reference (baseline/_ref, oracle/stage_reference.py) is absent.

The measured differences and the step times are appended to gpurun_out/reference_e2e.log (committed copy:
profiles/r2_reference_e2e.txt).
"""
import importlib
import os
import time

import pytest
import torch

from oracle import reference
from dro_sfm_b200 import synthetic as syn

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference.available(), reason="reference tree not staged")]
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOSS_CFG = dict(num_scales=4, progressive_scaling=0.0, rotation_mode="euler", upsample_depth_maps=True, ssim_loss_weight=0.85,
                occ_reg_weight=0.1, smooth_loss_weight=0.001, C1=1e-4, C2=9e-4, photometric_reduce_op="min", disp_norm=True,
                clip_loss=0.0, padding_mode="zeros", automask_loss=True, supervised_method="sparse-l1",
                supervised_num_scales=4, supervised_loss_weight=0.9)      # configs/default_config.py:88-113


def _log(line):
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "reference_e2e.log"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
    print(line)


def _batch(wl, B, supervised):
    g = syn.gen(7)
    image = syn.images(g, B, wl.H, wl.W)
    context = [0.7 * torch.roll(image, 3 * (v + 1), 3) + 0.3 * syn.images(g, B, wl.H, wl.W) for v in range(wl.V)]
    batch = {"rgb": image, "rgb_context": context, "rgb_original": image, "rgb_context_original": context,
             "intrinsics": syn.intrinsics(wl.dataset, B, wl.H, wl.W)}            # float64, as numpy collation delivers it
    if supervised:
        inv = syn.inv_depth(g, B, wl.H, wl.W, wl.min_depth, wl.max_depth)
        depth = (1.0 / inv) * (torch.rand(B, 1, wl.H, wl.W, generator=g) > 0.3)
        from oracle import pose_vec_to_T
        batch["depth"] = depth
        batch["pose_context"] = [pose_vec_to_T(syn.pose_vec(g, B, wl.dataset) * 0.3) for _ in range(wl.V)]
    return {k: ([x.to(DEV) for x in v] if isinstance(v, list) else v.to(DEV)) for k, v in batch.items()}


def _fresh(batch):
    """flip_lr_intr (utils/image.py:60-79) negates fx IN PLACE in batch['intrinsics']: every step gets its own copy."""
    return dict(batch, intrinsics=batch["intrinsics"].clone())


def _step(model, batch):
    model.zero_grad(set_to_none=True)
    out = model(_fresh(batch))
    out["loss"].sum().backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    poses = [[p.mat.detach().clone() for p in pv] for pv in out["poses"]]
    return out["loss"].detach().clone(), [d.detach().clone() for d in out["inv_depths"]], poses, grads


def _timed(model, batch, reps=3):
    _step(model, batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        model.zero_grad(set_to_none=True)
        model(_fresh(batch))["loss"].sum().backward()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def _step_float64(build_model, net, batch):
    """The stock reference in float64 on the GPU: what both fp32 passes are measured against.  The reference casts the
    intrinsics with K.float() (DepthPoseNet.py:84-85, multiview_photometric_loss_mf.py:162-163); for this pass only,
    Tensor.float is a cast to double.  Returns None if the reference cannot run in float64."""
    import copy
    import dro_sfm.geometry.pose as ref_pose
    orig, orig_identity = torch.Tensor.float, ref_pose.Pose.__dict__["identity"]
    try:
        net64 = copy.deepcopy(net).double()
        torch.Tensor.float = lambda t, *a, **k: t.double()
        # Pose.identity's dtype default is torch.float (pose.py:28); the identity target pose must be double here
        ref_pose.Pose.identity = classmethod(lambda cls, N=1, device=None, dtype=torch.float64:
                                             orig_identity.__func__(cls, N, device, torch.float64))
        model = build_model(net64).double()
        b64 = {k: ([x.double() for x in v] if isinstance(v, list) else v.double()) for k, v in batch.items()}
        return _step(model, b64)
    except Exception as e:           # noqa: BLE001
        _log("   float64 comparator unavailable: %r" % (e,))
        return None
    finally:
        torch.Tensor.float = orig
        ref_pose.Pose.identity = orig_identity


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("kind,wl_name,B,flip,lockstep", [("selfsup", "train_kitti_mf_selfsup_192x640", 1, 0.0, True),
                                                         ("selfsup", "train_kitti_mf_selfsup_192x640", 1, 0.0, False),
                                                         ("selfsup", "train_kitti_mf_selfsup", 2, 1.0, True),
                                                         ("sup", "train_scannet_mf_gt_view3", 2, 0.0, True)])
def test_reference_model_stock_vs_patched(kind, wl_name, B, flip, lockstep):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    wl = syn.WORKLOADS[wl_name]
    reference.load()
    version = "it8-seq4-inter-out" if kind == "selfsup" else "it12-h-out"
    net = reference.build_depth_pose_net(version, wl.min_depth, wl.max_depth, seed=3).to(DEV).train()
    batch = _batch(wl, B, kind == "sup")
    mod_name, cls_name = ("dro_sfm.models.SelfSupModelMF", "SelfSupModelMF") if kind == "selfsup" else \
        ("dro_sfm.models.SupModelMF", "SupModelMF")
    kwargs = dict(LOSS_CFG, flip_lr_prob=flip, min_depth=wl.min_depth, max_depth=wl.max_depth)

    def build(depth_net=None):
        model = getattr(importlib.import_module(mod_name), cls_name)(**kwargs)
        model.add_depth_net(net if depth_net is None else depth_net)
        return model.to(DEV).train()

    # the stock pass must see the stock reference: this test has to run before anything calls patch.install()
    import dro_sfm.geometry.pose as ref_pose
    from dro_sfm_b200.geometry import Pose as MyPose
    if ref_pose.Pose is MyPose:
        pytest.skip("the reference tree is already patched in this process (run this file on its own)")
    stock = build()
    loss0, inv0, poses0, grads0 = _step(stock, batch)
    ms_stock = _timed(stock, batch)
    truth = _step_float64(build, net, batch)

    state = _install(lockstep)
    try:
        patched = build()
        from dro_sfm_b200 import _lib as L
        before = L.lib().drosfm_launch_count()
        loss1, inv1, poses1, grads1 = _step(patched, batch)
        launched = int(L.lib().drosfm_launch_count() - before)
        ms_patched = _timed(patched, batch)
    finally:
        _uninstall(state)
    assert launched > 0, "the patched model did not reach the CUDA library"
    assert type(patched._photometric_loss if kind == "selfsup" else patched._loss).__module__.startswith("dro_sfm_b200")

    d_loss = abs(float(loss1.sum()) - float(loss0.sum())) / abs(float(loss0.sum()))
    d_inv = max(_rel(a, b) for a, b in zip(inv1, inv0))
    d_pose = max(_rel(a, b) for pa, pb in zip(poses1, poses0) for a, b in zip(pa, pb))
    assert set(grads0) == set(grads1)
    g0 = torch.cat([grads0[k].flatten() for k in sorted(grads0)])
    g1 = torch.cat([grads1[k].flatten() for k in sorted(grads1)])
    d_grad = _rel(g1, g0)
    worst = max(((_rel(grads1[k], grads0[k]), k) for k in grads0 if float(grads0[k].norm()) > 1e-6 * float(g0.norm())))
    _log("%s %s B=%d flip=%.0f %s: %d predictions, %d C-ABI launches/step | rel. diff patched vs stock: loss %.2e, inv_depths %.2e, "
         "poses %.2e, all parameter gradients %.2e (worst tensor %.2e %s) | whole training step fwd+bwd: stock %.1f ms, "
         "patched %.1f ms (x%.2f)" % (cls_name, wl_name, B, flip, "lock-step" if lockstep else "reference schedule", len(inv0), launched, d_loss, d_inv, d_pose, d_grad, worst[0],
                                      worst[1], ms_stock, ms_patched, ms_stock / ms_patched))
    # the recurrent network (8-12 GRU steps feeding the cost back) amplifies rounding-level differences of the cost maps;
    # fp32 convolutions re-associate at the 1e-6 level themselves
    assert d_loss <= 2e-4 and d_inv <= 2e-4 and d_pose <= 2e-4, (d_loss, d_inv, d_pose)
    if truth is not None:
        # parameter gradients: both fp32 passes against the float64 pass of the stock reference.  The loss gradient
        # w.r.t. the inverse depths is ill-conditioned in fp32, so stock fp32 is itself ~1e-3 away
        g64 = torch.cat([truth[3][k].flatten() for k in sorted(grads0)])
        e_stock, e_patched = _rel(g0, g64), _rel(g1, g64)
        _log("   parameter gradients vs the float64 pass of the stock reference: stock fp32 %.2e, patched %.2e (loss %.2e / %.2e)"
             % (e_stock, e_patched, abs(float(loss0.sum()) - float(truth[0].sum())) / abs(float(truth[0].sum())),
                abs(float(loss1.sum()) - float(truth[0].sum())) / abs(float(truth[0].sum()))))
        assert e_patched <= 2.0 * e_stock + 1e-4, (e_patched, e_stock)
    else:
        assert d_grad <= 5e-3, d_grad


@pytest.mark.parametrize("kind,wl_name,B", [("selfsup", "train_kitti_mf_selfsup", 2), ("sup", "train_scannet_mf_gt_view3", 2)])
def test_reference_model_training_step_as_cuda_graph(kind, wl_name, B):
    """the patched reference model's whole training step (encoder, GRU iterations, cost calls, loss, backward)
    recorded in ONE CUDA graph (dro_sfm_b200.graphs.GraphedStep) -- same loss and gradients as the eager step, on a second
    batch as well (the graph must read the static buffers, not constants baked in at capture)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    from dro_sfm_b200.graphs import GraphedStep
    wl = syn.WORKLOADS[wl_name]
    reference.load()
    version = "it8-seq4-inter-out" if kind == "selfsup" else "it12-h-out"
    net = reference.build_depth_pose_net(version, wl.min_depth, wl.max_depth, seed=3).to(DEV).train()
    batch = _batch(wl, B, kind == "sup")
    batch2 = dict(batch, rgb=torch.roll(batch["rgb"], 5, 2), rgb_original=torch.roll(batch["rgb"], 5, 2))
    mod_name, cls_name = ("dro_sfm.models.SelfSupModelMF", "SelfSupModelMF") if kind == "selfsup" else \
        ("dro_sfm.models.SupModelMF", "SupModelMF")
    kwargs = dict(LOSS_CFG, flip_lr_prob=0.0, min_depth=wl.min_depth, max_depth=wl.max_depth)
    state = _install(True)
    try:
        model = getattr(importlib.import_module(mod_name), cls_name)(**kwargs)
        model.add_depth_net(net)
        model = model.to(DEV).train()
        for m in model.modules():                      # BatchNorm running statistics would differ between the two arms
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        eager = [_step(model, b) for b in (batch, batch2)]
        ms_eager = _timed(model, batch)
        step = GraphedStep(lambda b: model(b), batch, model.parameters())
        got = []
        for b in (batch, batch2, batch):
            out = step(b)
            torch.cuda.synchronize()
            got.append((out["loss"].detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step(batch)
        torch.cuda.synchronize()
        ms_graph = (time.perf_counter() - t0) / 5 * 1e3
        # the same pair with PyTorch's default convolution precision (TF32 allowed), as a user trains: timing only
        torch.backends.cudnn.allow_tf32 = True
        ms_eager_tf32 = _timed(model, batch)
        step_tf32 = GraphedStep(lambda b: model(b), batch, model.parameters())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step_tf32(batch)
        torch.cuda.synchronize()
        ms_graph_tf32 = (time.perf_counter() - t0) / 5 * 1e3
    finally:
        torch.backends.cudnn.allow_tf32 = False
        _uninstall(state)
    for i, j in ((0, 0), (1, 1), (2, 0)):
        loss_e, _, _, grads_e = eager[j]
        loss_g, grads_g = got[i]
        d_loss = abs(float(loss_g.sum()) - float(loss_e.sum())) / abs(float(loss_e.sum()))
        ge = torch.cat([grads_e[k].flatten() for k in sorted(grads_e)])
        gg = torch.cat([grads_g[k].flatten() for k in sorted(grads_e)])
        d_grad = _rel(gg, ge)
        _log("%s %s B=%d CUDA-graph replay %d vs eager: loss %.2e, parameter gradients %.2e" % (cls_name, wl_name, B, i, d_loss, d_grad))
        # same kernels, same inputs: only the order of the atomic accumulations differs between two runs
        assert d_loss <= 1e-5 and d_grad <= 2e-3, (i, d_loss, d_grad)
    assert float((got[0][0] - got[1][0]).abs().sum()) > 0.0, "the second batch did not reach the graph"
    _log("%s %s B=%d whole training step fwd+bwd: eager (patched) %.1f ms, one CUDA graph %.1f ms (x%.2f)"
         % (cls_name, wl_name, B, ms_eager, ms_graph, ms_eager / ms_graph))
    _log("   with cuDNN TF32 convolutions (PyTorch default): eager %.1f ms, one CUDA graph %.1f ms (x%.2f)"
         % (ms_eager_tf32, ms_graph_tf32, ms_eager_tf32 / ms_graph_tf32))


def _install(lockstep=True):
    """patch.install() with a record of what it replaced, so that the next parametrisation starts from the stock tree."""
    import sys
    import dro_sfm_b200.patch as patch
    names = ["dro_sfm.geometry.pose", "dro_sfm.geometry.camera", "dro_sfm.geometry.camera_utils",
             "dro_sfm.losses.multiview_photometric_loss_mf", "dro_sfm.losses.supervised_loss", "dro_sfm.models.SelfSupModelMF",
             "dro_sfm.models.SupModelMF", "dro_sfm.models.SemiSupModelMF", "dro_sfm.models.SfmModelMF",
             "dro_sfm.networks.depth_pose.DepthPoseNet", "dro_sfm.utils.depth"]
    saved = {}
    for n in names:
        try:
            m = importlib.import_module(n)
        except Exception:
            continue
        saved[n] = dict(vars(m))
    net_cls = sys.modules["dro_sfm.networks.depth_pose.DepthPoseNet"].DepthPoseNet
    saved_methods = {k: net_cls.__dict__[k] for k in ("get_cost_each", "depth_cost_calc", "upsample_depth", "forward")
                     if k in net_cls.__dict__}
    patch.install(lockstep=lockstep)
    return saved, net_cls, saved_methods


def _uninstall(state):
    import sys
    saved, net_cls, saved_methods = state
    for n, d in saved.items():
        m = sys.modules[n]
        for k, v in d.items():
            if m.__dict__.get(k) is not v:
                setattr(m, k, v)
    for k, v in saved_methods.items():
        setattr(net_cls, k, v)
