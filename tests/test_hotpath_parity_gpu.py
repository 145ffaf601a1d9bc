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
    loss32, g32 = bench.cpu_step(wl, batch, threads, torch.float32, return_grads=True, pose_to_T=euler_T_as_on_gpu)
    loss64, g64 = bench.cpu_step(wl, batch, threads, torch.float64, return_grads=True)
    names = _names(wl)
    assert len(names) == len(g32)

    def check(step, tag):
        loss = step.step()
        torch.cuda.synchronize()
        assert_close(loss.detach().cpu().reshape(()), np.float32(loss32), what=f"loss ({tag}) vs CPU arm fp32")
        assert abs(float(loss) - loss64) <= ATOL + RTOL * abs(loss64), (tag, float(loss), loss64)
        grads = [t.grad for t in step.leaves()]
        for name, g, r32, r64 in zip(names, grads, g32, g64):
            assert g is not None, name
            assert_close_or_better(g.detach().cpu(), r32, r64, what=f"{name}[{tag}]")

    step = HotPathStep(wl, DEV, B=B, seed=1234)
    check(step, "eager")
    step2 = HotPathStep(wl, DEV, B=B, seed=1234)        # fresh leaves for the capture stream
    step2.capture(warmup=1)
    check(step2, "graph")
    step2.step()
    check(step2, "replay2")                             # a replay leaves no state behind (workspace, sinks)
