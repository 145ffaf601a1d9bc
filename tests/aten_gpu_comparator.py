"""Stock-ATen comparator on the same GPU ("also time the reference on the same GPU").

Not a test and not part of bench.py's contract: it runs the oracle -- the functional PyTorch restatement of the
reference's hot path, i.e. the same chains of stock ATen ops the reference issues -- on cuda:0 for the benchmark
workload and prints one JSON line, so that DESIGN.md can put the hand-written kernels beside what PyTorch itself
achieves on a B200.  /root/reference does not exist on the GPU box; the oracle is the travelling stand-in.

    python tests/aten_gpu_comparator.py [--workload train_kitti_mf_selfsup] [--batch 2] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def to_dev(x, dev):
    if torch.is_tensor(x):
        return x.to(dev)
    if isinstance(x, (list, tuple)):
        return [to_dev(y, dev) for y in x]
    return x


def step(wl, batch, dev):
    """bench.cpu_step on `dev`: 2*V*T cost evaluations fwd+bwd + the loss fwd+bwd with the oracle's ATen chains."""
    import oracle
    K = batch["K"]
    fmap = batch["fmap"].clone().requires_grad_(True)
    frefs = [f.clone().requires_grad_(True) for f in batch["fmaps_ref"]]
    gout = batch["_gout"]
    outs = []
    for t in range(wl.T):
        inv = batch["inv_depth_lr"][t].clone().requires_grad_(True)
        outs.append(oracle.depth_cost(inv, fmap, frefs, [p for p in batch["pose_lr"][t]], K, K, 0.125))
        depth = oracle.inv2depth(batch["inv_depth_lr"][(t // wl.seq_len) * wl.seq_len])
        for v in range(wl.V):
            pose = batch["pose_lr"][t][v].clone().requires_grad_(True)
            outs.append(oracle.feat_cost_each(pose, fmap, frefs[v], depth, K, K, 0.125))
    invs = [x.clone().requires_grad_(True) for x in batch["inv_depths"]]
    pvec = [[p.clone().requires_grad_(True) for p in row] for row in batch["poses"]]
    Ts = [[oracle.pose_vec_to_T(p) for p in row] for row in pvec]
    if wl.supervised:
        loss = oracle.reproj_pose_loss(Ts, [oracle.pose_vec_to_T(p) for p in batch["gt_poses"]],
                                       oracle.inv2depth(batch["gt_inv_depth"]), K, K, wl.min_depth, wl.max_depth) \
            + oracle.supervised_depth_loss(invs, batch["gt_inv_depth"], wl.min_depth, wl.max_depth)
    else:
        loss, _ = oracle.multiview_photometric_decay_loss(batch["image"], batch["context"], invs, K, K, Ts, smooth_w=0.001,
                                                          automask=True, reduce_op="min")
    torch.autograd.backward([loss.sum()] + outs, [torch.ones((), device=dev)] + [gout] * len(outs))
    return float(loss.sum())


def main():
    import bench
    from dro_sfm_b200 import synthetic as syn
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="train_kitti_mf_selfsup")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    wl = syn.WORKLOADS[args.workload]
    B = args.batch or wl.B
    dev = torch.device("cuda:0")
    batch = {k: to_dev(v, dev) for k, v in syn.hot_path_batch(wl, seed=1234, C=128, B=B).items()}
    batch["_gout"] = torch.randn(B, 128, wl.H // 8, wl.W // 8, generator=torch.Generator().manual_seed(99)).to(dev)
    for _ in range(2):
        step(wl, batch, dev)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step(wl, batch, dev)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    print(json.dumps({"impl": "stock ATen ops on cuda:0 (oracle port of the reference)", "metric": bench.METRIC,
                      "value": B / (ms * 1e-3), "unit": bench.UNIT, "ms_per_step": ms, "batch": B, "workload": wl.name,
                      "steps": args.steps, "note": "eager PyTorch, launch-bound; includes the float(loss) sync per step"}))


if __name__ == "__main__":
    main()
