"""Cost phase only (all cost calls of a step, fwd+bwd), CUDA-graph replay, L2 flushed: lock-step on/off x pixels per warp."""
import os, sys, json
import torch
sys.path.insert(0, '.')
from dro_sfm_b200 import synthetic as syn
from dro_sfm_b200.hotpath import HotPathStep

def cost_only(step):
    wl = step.wl
    from dro_sfm_b200.networks import cost_batch, depth_cost_calc, get_cost_each
    costs = []
    fmap_, frefs_ = step.feature_maps()
    for t in range(wl.T):
        poses_t = [p.detach() for p in step.pose_lr[t]]
        depth = step.depth_lr[t // wl.seq_len]
        if step.lockstep:
            jobs = [(step.inv_lr[t], fmap_, frefs_, poses_t, True)]
            jobs += [(depth, fmap_, [frefs_[v]], [step.pose_lr[t][v]], False) for v in range(wl.V)]
            costs += cost_batch(jobs, step.K, step.K, 1.0 / 8)
        else:
            costs.append(depth_cost_calc(step.inv_lr[t], fmap_, frefs_, poses_t, step.K, step.K, 1.0 / 8))
            for v in range(wl.V):
                costs.append(get_cost_each(step.pose_lr[t][v], fmap_, frefs_[v], depth, step.K, step.K, 1.0 / 8))
    torch.autograd.backward(costs, step.g_costs)

def bench(wl_name, lockstep, ppw, B=None, reps=30):
    if ppw: os.environ["DROSFM_PPW"] = str(ppw)
    else: os.environ.pop("DROSFM_PPW", None)
    wl = syn.WORKLOADS[wl_name]
    step = HotPathStep(wl, "cuda:0", B=B, lockstep=lockstep)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step.zero_grads(); cost_only(step)
    torch.cuda.current_stream().wait_stream(s)
    step.zero_grads()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        cost_only(step)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3

