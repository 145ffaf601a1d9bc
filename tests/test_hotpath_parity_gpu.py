"""Parity of the BENCHMARKED configuration: one whole hot-path step (every cost call fwd+bwd + the loss fwd+bwd) of
`train_kitti_mf_selfsup` (320x960, V=2, T=8, n=9 -- BASELINE.json configs[1]) on the B200, eager and as CUDA-graph replay,
against `bench.py:cpu_step` (the CPU arm of the benchmark, i.e. the oracle) in float32 and float64: the loss and the
gradient of EVERY leaf (feature maps, low-resolution inverse depths and pose vectors of all GRU steps, the n
full-resolution predictions and the V x n loss poses).

Poses enter as [B,6] euler vectors on both sides (the entry the training loop uses).  The oracle converts them with
`euler_T_as_on_gpu`: the reference's euler2mat evaluated by torch on this GPU (bit-identical to the kernels' prologue,
see test_pose_vec2mat), so that both sides see the same matrices.
"""
import numpy as np
import pytest
import torch

import bench
from conftest import assert_close, assert_close_or_better, euler_T_as_on_gpu, RTOL, ATOL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _names(wl):
    names = ["g_fmap"] + [f"g_fref{v}" for v in range(wl.V)] + [f"g_inv_lr{t}" for t in range(wl.T)]
    names += [f"g_pose_lr{t}_{v}" for t in range(wl.T) for v in range(wl.V)]
    names += [f"g_inv{i}" for i in range(wl.n)] + [f"g_pose{v}_{i}" for v in range(wl.V) for i in range(wl.n)]
    return names


@pytest.mark.parametrize("wl_name,B", [("train_kitti_mf_selfsup", 1), ("train_scannet_mf_gt_view3", 2)])
def test_full_step_matches_the_cpu_arm(wl_name, B):
    from dro_sfm_b200 import synthetic as syn
    from dro_sfm_b200.hotpath import HotPathStep
    wl = syn.WORKLOADS[wl_name]
    batch = syn.hot_path_batch(wl, seed=1234, C=128, B=B)
    threads = torch.get_num_threads()
    names = _names(wl)

    step = HotPathStep(wl, DEV, B=B, seed=1234)
    loss_eager = step.step()
    torch.cuda.synchronize()
    grads_eager = [g.detach().cpu().clone() for g in step.grads()]
    sel = getattr(step.loss_mod, "last_selection", None)

    # (1) the loss against the CPU arm as bench.py runs it (its own per-pixel arg-min)
    maps = []
    loss32 = bench.cpu_step(wl, batch, threads, torch.float32, pose_to_T=euler_T_as_on_gpu, maps_out=maps)
    assert_close(loss_eager.detach().cpu().reshape(()), np.float32(loss32), what="loss vs CPU arm fp32")

    # (2) self-supervised: the per-pixel min over the 2V photometric maps is discontinuous -- at a near-tie two fp32
    # evaluations may pick different winners, which swaps a whole gradient term.  The selection of the kernels must
    # equal the oracle's except at near-ties; the gradients are then compared for the SAME selection (the kernels').
    forced = None
    if sel is not None:
        sel = sel.cpu()
        forced = [torch.where(s == 255, torch.full_like(s, -1, dtype=torch.long), s.long()) for s in sel]
        flips = 0
        for i, stack in enumerate(maps):
            ref = stack.argmin(1)
            same = torch.where(sel[i] == 255, ref % 2 == 1, sel[i].long() * 2 == ref)
            srt = torch.sort(stack, 1)[0]
            margin = (srt[:, 1] - srt[:, 0]) / srt[:, 0].clamp(min=1e-6)
            assert (margin[~same] < 1e-3).all(), "arg-min differs from the oracle away from a tie (prediction %d)" % i
            flips += int((~same).sum())
        print("per-pixel arg-min: %d of %d pixel-predictions differ from the fp32 oracle (all at relative margins < 1e-3)"
              % (flips, sel.numel()))
        assert flips <= 1e-4 * sel.numel()
    loss32f, g32 = bench.cpu_step(wl, batch, threads, torch.float32, return_grads=True, pose_to_T=euler_T_as_on_gpu, forced_sel=forced)
    loss64f, g64 = bench.cpu_step(wl, batch, threads, torch.float64, return_grads=True, forced_sel=forced)
    assert len(names) == len(g32)
    assert abs(float(loss_eager.detach()) - loss64f) <= ATOL + RTOL * abs(loss64f), (float(loss_eager.detach()), loss64f)

    def check(grads, tag):
        for name, g, r32, r64 in zip(names, grads, g32, g64):
            assert g is not None, name
            assert_close_or_better(g, r32, r64, what=f"{name}[{tag}]")

    check(grads_eager, "eager")
    # (3) the same step captured in a CUDA graph (what bench.py times), replayed twice: a replay leaves no state behind
    step2 = HotPathStep(wl, DEV, B=B, seed=1234)        # fresh leaves for the capture stream
    step2.capture(warmup=1)
    for tag in ("graph", "replay2"):
        loss = step2.step()
        torch.cuda.synchronize()
        assert abs(float(loss.detach()) - float(loss_eager.detach())) <= 1e-6 * abs(float(loss_eager.detach())), tag
        if sel is not None:
            assert torch.equal(step2.loss_mod.last_selection.cpu(), sel), tag
        check([g.detach().cpu() for g in step2.grads()], tag)
