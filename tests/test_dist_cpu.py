"""world_size-2 gloo test of the data-parallel plumbing (runs on CPU; the N>1 GPU path uses the same
helpers with the nccl backend)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from dro_sfm_b200 import dist_utils as du, synthetic as syn
    r, w = du.init("gloo")
    assert (r, w) == (rank, world)
    # every rank draws its own shard: different seeds, same shapes
    wl = syn.WORKLOADS["train_scannet_mf_selfsup_view5"]
    g = syn.gen(du.shard_seed(1234, rank))
    B, H, W = 1, 24, 32
    K = syn.intrinsics(wl.dataset, B, H, W)
    image = syn.images(g, B, H, W)
    context = [syn.images(g, B, H, W) for _ in range(2)]
    invs = [syn.inv_depth(g, B, H, W, wl.min_depth, wl.max_depth)]
    Ts = [[oracle.pose_vec_to_T(syn.pose_vec(g, B, wl.dataset) * 0.1)] for _ in range(2)]
    loss, _ = oracle.multiview_photometric_decay_loss(image, context, invs, K.float(), K.float(), Ts)
    local = float(loss)
    avg = du.average_loss(loss.detach().clone())
    times = du.max_over_ranks([1.0 + rank, 5.0 - rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    if rank == 0:
        out.put((gathered, float(avg), times))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_reductions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, avg, times = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered[0] != gathered[1], "ranks must work on different shards"
    assert abs(avg - sum(gathered) / world) < 1e-6
    assert times == [2.0, 5.0]


def test_whole_job_rate_is_weak_scaling():
    from dro_sfm_b200 import dist_utils as du
    assert du.whole_job_rate(units_per_rank=20, world=8, seconds=0.5) == 320.0
    assert du.shard_seed(1234, 3, 2) == 1234 + 3000 + 2


def test_bind_host_to_gpu_is_best_effort():
    """Without NVML / a GPU the NUMA binding is a no-op that reports None and leaves the affinity alone."""
    import os
    from dro_sfm_b200 import dist_utils as du
    before = os.sched_getaffinity(0)
    cpus = du.bind_host_to_gpu(0)
    assert cpus is None or set(cpus) <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
