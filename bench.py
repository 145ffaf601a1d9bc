#!/usr/bin/env python
"""Hot-path benchmark of dro_sfm_b200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--layout nchw|nhwc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # CPU arm: the reference's algorithm on the host cores

A step = one pass of the dense depth-pose warping hot path over one batch (dro_sfm_b200/hotpath.py):
2*V*T feature-metric cost evaluations fwd+bwd plus the multi-view photometric (or supervised
reprojection) loss fwd+bwd, issued through the reference's operator surface.  Metric: frames/s
(frame = one batch sample).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "warp+cost+photometric fwd+bwd frames/sec"
UNIT = "frames/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# C-ABI call -> the CUDA kernels it launches (names as in the ncu reports); the staged photometric calls are two each
NCU_NAMES = {"photometric_bwd": ["ssim_bwd_stream"], "photometric_fwd": ["ssim_train_stream2_kernel"],
             "feat_cost_batch_fwd": ["feat_cost_fwd_nhwc<2>"], "feat_cost_batch_bwd": ["feat_cost_bwd_nhwc<2>"],
             "photometric_loss_fwd": ["ssim_fwd_stream2_kernel<1>", "pack_rgbx_kernel", "warp_sources_kernel<1>",
                                      "ssim_train_stream2_kernel"],
             "photometric_loss_bwd": ["warp_sources_adjoint_kernel<1, 1>"],
             "warp_sources_fwd": ["pack_rgbx_kernel", "warp_sources_kernel<1>"], "warp_sources_bwd": ["warp_sources_adjoint_kernel<1, 1>"],
             "feat_cost_fwd_v1": ["feat_cost_fwd_nhwc<1>"], "feat_cost_bwd_v1": ["feat_cost_bwd_nhwc<1>"],
             "feat_cost_fwd_vN": ["feat_cost_fwd_nhwc<2>"], "feat_cost_bwd_vN": ["feat_cost_bwd_nhwc<2>"],
             "automask_fwd": ["ssim_fwd_stream2_kernel<1>"], "smoothness_fwd": ["smooth_fwd_kernel"],
             "smoothness_bwd": ["smooth_bwd_kernel"]}


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the kernel from the latest committed `ncu --set full` capture (profiles/*_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files or kernel_key not in NCU_NAMES:
        return None, None
    with open(files[-1]) as f:
        table = json.load(f)
    total, found = 0.0, 0
    for want in NCU_NAMES[kernel_key]:
        for name, t in table.items():
            if name.startswith(want):
                total += t["dram_read_MB"] + t["dram_write_MB"]
                found += 1
                break
    if found != len(NCU_NAMES[kernel_key]):
        return None, None
    return int(total * 1e6), os.path.basename(files[-1])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms from the warm-up through the timed regions."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port, PyTorch CPU) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step(wl, batch, threads, dtype=torch.float32, return_grads=False, pose_to_T=None, forced_sel=None, maps_out=None):
    """One frame-batch of the same hot-path work with the CPU oracle (test infrastructure used here only
    as the reported baseline / checker): 2*V*T cost evaluations fwd+bwd + the loss fwd+bwd.

    return_grads: also return the gradient of every leaf in HotPathStep.leaves() order.  pose_to_T: how a [B,6] pose
    vector becomes a [B,4,4] matrix (default: the oracle's Pose.from_vec on the CPU).  forced_sel / maps_out: see
    oracle.photometric_loss (checker-only: the per-pixel arg-min of the photometric loss taken from another evaluation)."""
    import oracle
    torch.set_num_threads(threads)
    to_T = pose_to_T or oracle.pose_vec_to_T
    c = lambda x: x.to(dtype)                                                    # noqa: E731
    K = batch["K"].float().to(dtype)                                             # the reference casts K.float() first
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
        outs.append(oracle.depth_cost(inv, fmap, frefs, [to_T(c(p)) for p in batch["pose_lr"][t]], K, K, 0.125))
        depth = oracle.inv2depth(c(batch["inv_depth_lr"][(t // wl.seq_len) * wl.seq_len]))
        for v in range(wl.V):
            pose = c(batch["pose_lr"][t][v]).clone().requires_grad_(True)
            pose_lr.append(pose)
            outs.append(oracle.feat_cost_each(to_T(pose), fmap, frefs[v], depth, K, K, 0.125))
    invs = [c(x).clone().requires_grad_(True) for x in batch["inv_depths"]]
    pvec = [[c(p).clone().requires_grad_(True) for p in row] for row in batch["poses"]]
    Ts = [[to_T(p) for p in row] for row in pvec]
    if wl.supervised:
        loss = oracle.reproj_pose_loss(Ts, [to_T(c(p)) for p in batch["gt_poses"]],
                                       oracle.inv2depth(c(batch["gt_inv_depth"])), K, K, wl.min_depth, wl.max_depth) \
            + oracle.supervised_depth_loss(invs, c(batch["gt_inv_depth"]), wl.min_depth, wl.max_depth)
    else:
        loss, _ = oracle.multiview_photometric_decay_loss(c(batch["image"]), [c(x) for x in batch["context"]], invs, K, K, Ts,
                                                          smooth_w=0.001, automask=True, reduce_op="min",
                                                          forced_sel=forced_sel, maps_out=maps_out)
    torch.autograd.backward([loss.sum()] + outs, [torch.ones((), dtype=dtype)] + gouts)
    if not return_grads:
        return float(loss.detach().sum())
    leaves = [fmap] + frefs + inv_lr + pose_lr + invs + [p for row in pvec for p in row]
    return float(loss.detach().sum()), [x.grad for x in leaves]


def cpu_arm():
    """(kind, step function) of the CPU arm: the UNMODIFIED reference's own functions when its package is staged
    (baseline/_ref, oracle/stage_reference.py; kind "reference"), else the oracle port (kind "port"; bit-equal to the
    reference on the CPU, tests/test_oracle_vs_reference.py)."""
    from oracle import reference
    if reference.available():
        def ref_step(wl, batch, threads):
            torch.set_num_threads(threads)
            return reference.hot_path_step(wl, batch, "cpu")
        return "reference", ref_step
    return "port", cpu_step


def time_cpu(wl, B, reps, threads):
    """Returns (frames/s, seconds per step, loss, kind)."""
    from dro_sfm_b200 import synthetic as syn
    kind, fn = cpu_arm()
    batch = syn.hot_path_batch(wl, seed=1234, C=128, B=B)
    loss = fn(wl, batch, threads)                    # warm-up
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(wl, batch, threads)
        best = min(best, time.perf_counter() - t0)
    return B / best, best, loss, kind


def time_gpu_aten_reference(wl, B, dev, steps=3):
    """The REFERENCE's own cost + loss calls (stock ATen ops, TF32 off) on the same GPU: the like-for-like comparator of
    SURVEY.md section 8(d).  None when the reference package is not staged."""
    from oracle import reference
    if not reference.available():
        return None
    from dro_sfm_b200 import synthetic as syn
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        batch = syn.hot_path_batch(wl, seed=1234, C=128, B=B)
        for _ in range(2):
            loss = reference.hot_path_step(wl, batch, dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            reference.hot_path_step(wl, batch, dev)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return {"value": B / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "batch": B, "steps": steps, "loss": loss,
            "what": "the unmodified reference's get_cost_each / depth_cost_calc / loss modules (stock ATen, eager, TF32 off) "
                    "on the same B200, same inputs and call sequence, inputs resident on the device"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from dro_sfm_b200 import synthetic as syn
    B = 1
    kind, fn = cpu_arm()
    batch = syn.hot_path_batch(wl, seed=1234, C=128, B=B)
    for _ in range(min(args.warmup, 1)):
        fn(wl, batch, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(wl, batch, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = B / dt
    sample = "B=1 frame of %s per step (all %d cost calls + loss, fwd+bwd), %s" % (
        wl.name, 2 * wl.V * wl.T, "the unmodified reference's own functions on PyTorch-CPU (baseline/_ref)" if kind == "reference"
        else "PyTorch-CPU oracle port of the reference")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "H": wl.H, "W": wl.W, "views": wl.V, "gru_steps": wl.T, "predictions": wl.n,
                   "batch_per_step": B, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
KERNEL_KEYS = {
    "drosfm_feat_cost_fwd": lambda a: "feat_cost_fwd_v1" if a[6] == 1 else "feat_cost_fwd_vN",
    "drosfm_feat_cost_bwd": lambda a: "feat_cost_bwd_v1" if a[7] == 1 else "feat_cost_bwd_vN",
    "drosfm_photometric_fwd": lambda a: "photometric_fwd", "drosfm_photometric_bwd": lambda a: "photometric_bwd",
    "drosfm_automask_fwd": lambda a: "automask_fwd", "drosfm_smoothness_fwd": lambda a: "smoothness_fwd",
    "drosfm_smoothness_bwd": lambda a: "smoothness_bwd", "drosfm_reproj_loss_fwd": lambda a: "reproj_loss_fwd",
    "drosfm_reproj_loss_bwd": lambda a: "reproj_loss_bwd", "drosfm_pose_vec2mat_fwd": lambda a: "pose_vec2mat_fwd",
    "drosfm_pose_vec2mat_bwd": lambda a: "pose_vec2mat_bwd",
    "drosfm_warp_sources_fwd": lambda a: "warp_sources_fwd", "drosfm_warp_sources_bwd": lambda a: "warp_sources_bwd",
    "drosfm_feat_cost_batch_fwd": lambda a: "feat_cost_batch_fwd", "drosfm_feat_cost_batch_bwd": lambda a: "feat_cost_batch_bwd",
}
# The photometric loss is ONE operator per direction in SURVEY.md section 8(d); the staged path implements it with several
# launches.  The roofline treats each direction as a unit: SURVEY bytes over the summed duration of its launches.
UNITS = {"photometric_loss_fwd": ["automask_fwd", "warp_sources_fwd", "photometric_fwd"],
         "photometric_loss_bwd": ["photometric_bwd", "warp_sources_bwd"]}


def run_gpu(args, wl):
    import torch.distributed as dist
    from dro_sfm_b200 import _lib as L, dist_utils as du
    from dro_sfm_b200.hotpath import HotPathStep

    rank, world, local = du.env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_cpus = du.bind_host_to_gpu(local) if world > 1 else None      # NUMA-local pinned staging buffers (e2e leg)
    du.init("nccl", dev)
    L.lib()
    B = args.batch or wl.B
    step = HotPathStep(wl, dev, B=B, seed=du.shard_seed(1234, rank), channels_last=(args.layout == "nhwc"))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB of L2

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if not args.no_graph:
        step.capture(warmup=3)
    launches_per_step = getattr(step, "launches_per_step", None)
    loss_host = torch.empty(1, pin_memory=True)

    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    pending = {"ev": None}
    counter, log_every = {"n": 0}, max(1, args.log_every)

    def one_step(e2e):
        """e2e: the step consumes inputs that came from pinned host memory.  The H2D copy of the NEXT step's
        inputs is issued on a copy stream as soon as this step has taken its own out of the staging buffer,
        and the step does not end before that copy has landed, so every timed step contains exactly one full
        H2D transfer (overlapped with compute), one device-to-device hand-over and the D2H read of the loss."""
        if e2e:
            if pending["ev"] is None:
                pending["ev"] = step.prefetch(copy_stream)
            main.wait_event(pending["ev"])
            step.commit_staging()
            copy_stream.wait_stream(main)
            pending["ev"] = step.prefetch(copy_stream)
        loss = step.step()
        # the only collective on the path is the LOGGED loss (utils/reduce.py:10-30 reduces what the progress bar
        # shows): averaged over the ranks every `log_every` steps, off the per-step critical path
        counter["n"] += 1
        if counter["n"] % log_every == 0:
            du.average_loss(loss.detach())
        if e2e:
            loss_host.copy_(loss.detach(), non_blocking=True)
            main.wait_event(pending["ev"])
        return loss

    def timed(e2e, steps):
        evs = []
        for _ in range(steps):
            flush.zero_()                                         # evict L2 between timed iterations
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            one_step(e2e)
            e.record()
            evs.append((s, e))
        return evs

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                                           # sampled from the warm-up through both timed regions
    for _ in range(max(args.warmup, 3)):
        one_step(False)
        warm_loss = one_step(True)
    du.average_loss(warm_loss.detach().clone())          # NCCL sets its channels up on the first collective: not in the timed region
    if launches_per_step is None:
        before = L.lib().drosfm_launch_count()
        one_step(False)
        launches_per_step = int(L.lib().drosfm_launch_count() - before)

    sync_all()
    evs = timed(False, args.steps)
    sync_all()
    ms = sum(s.elapsed_time(e) for s, e in evs)
    evs = timed(True, args.steps)
    sync_all()
    ms_e2e = sum(s.elapsed_time(e) for s, e in evs)
    clk = clocks.stop() if rank == 0 else None
    ms, ms_e2e = du.max_over_ranks([ms, ms_e2e], device=dev)
    value = du.whole_job_rate(B * args.steps, world, ms * 1e-3)
    e2e_value = du.whole_job_rate(B * args.steps, world, ms_e2e * 1e-3)

    # per-kernel durations: the same K steps issued eagerly with CUDA events around every C-ABI launch
    roofline = None
    if rank == 0:
        from dro_sfm_b200 import ops as _ops
        graph, step.graph = step.graph, None
        # the same split calls on ONE stream: concurrent kernels would stretch each other's events
        overlap, _ops.OVERLAP = _ops.OVERLAP, ("serial" if _ops.OVERLAP else False)
        step.step()
        torch.cuda.synchronize()
        n_inst = min(args.steps, 10)
        L.profile_begin()
        for _ in range(n_inst):
            flush.zero_()
            # hold the stream back for ~15 ms so the host enqueues the whole step ahead of the device: the events
            # then bracket device time only, not the Python/ctypes launch latency of an idle GPU
            torch.cuda._sleep(30_000_000)
            step.step()
            torch.cuda.synchronize()
        recs = L.profile_end()
        step.graph = graph
        _ops.OVERLAP = overlap
        if os.environ.get("DROSFM_BENCH_DUMP"):
            n_one = len(recs) // n_inst
            for name, a, t_ms in recs[n_one:2 * n_one]:
                print("  %-28s %8.2f us" % (KERNEL_KEYS.get(name, lambda _a: name)(a), t_ms * 1e3), file=sys.stderr)
        per = {}
        for name, a, t_ms in recs:
            key = KERNEL_KEYS.get(name, lambda _a: name)(a)
            per.setdefault(key, []).append(t_ms)
        alg = step.algorithmic_bytes()
        stage = alg.pop("stage_operands")
        totals = {k: sum(v) for k, v in per.items()}
        counts = {k: len(v) // n_inst for k, v in per.items()}
        # launch units: single C-ABI calls, except the photometric loss (one unit per direction, see UNITS)
        unit_ms, unit_launches, unit_kernels = {}, {}, {}
        grouped = set()
        for unit, members in UNITS.items():
            present = [m for m in members if m in totals]
            if present:
                unit_ms[unit] = sum(totals[m] for m in present) / n_inst
                unit_launches[unit] = sum(counts[m] for m in present)
                unit_kernels[unit] = present
                grouped.update(present)
        for k in totals:
            if k not in grouped and k in alg:
                unit_ms[k] = totals[k] / len(per[k])            # per launch
                unit_launches[k] = counts[k]
                unit_kernels[k] = [k]
        per_step_ms = {k: (unit_ms[k] if k in UNITS else unit_ms[k] * unit_launches[k]) for k in unit_ms}
        dom = max(per_step_ms, key=lambda k: per_step_ms[k])
        peak, peak_src = measured_peaks()
        achieved = alg[dom] / (unit_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(dom)

        def entry(k):
            gbps = alg[k] / (unit_ms[k] * 1e-3) / 1e9
            return {"launches_per_step": unit_launches[k], "ms": round(unit_ms[k], 5), "bytes": alg[k], "GBps": round(gbps, 1),
                    "frac": round(gbps / peak, 4), "share_of_step_kernel_time": round(per_step_ms[k] / sum(per_step_ms.values()), 4),
                    "c_abi_calls": unit_kernels[k]}
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "note": "algorithmic bytes = SURVEY.md 8(d) figures (every distinct operand of the OPERATOR once); unit = one "
                            "C-ABI launch, except the photometric loss, which SURVEY defines as one fused operator per "
                            "direction: its unit is the direction's launches together (the staged path's warped copy and "
                            "g_warped are this design's traffic and are NOT counted).  Durations: CUDA events around every "
                            "C-ABI call of the step, eager, one stream, device held back so that host launch latency is "
                            "excluded, L2 flushed before each step.  The loss kernels are instruction-issue bound (ncu: "
                            "issue slots 50-75 % busy, DRAM < 15 %)",
                    "bytes_per_launch": alg[dom], "avg_launch_ms": unit_ms[dom], "launches": unit_launches[dom],
                    "c_abi_calls": unit_kernels[dom], "cuda_kernels": NCU_NAMES.get(dom), "timed_steps": n_inst,
                    "share_of_kernel_time": per_step_ms[dom] / sum(per_step_ms.values()),
                    "step_algorithmic_GBps": alg["step_total"] / (ms / args.steps * 1e-3) / 1e9,
                    "step_frac": alg["step_total"] / (ms / args.steps * 1e-3) / 1e9 / peak,
                    "kernel_ms_per_step": {k: round(v / n_inst, 4) for k, v in sorted(totals.items())},
                    "per_unit": {k: entry(k) for k in sorted(unit_ms)},
                    # individual stages of the staged photometric path against their own operand bytes (orientation only)
                    "per_stage_operands": {k: {"ms": round(totals[k] / len(per[k]), 5), "operand_bytes": stage[k],
                                               "GBps": round(stage[k] / (totals[k] / len(per[k]) * 1e-3) / 1e9, 1)}
                                           for k in sorted(stage) if k in totals}}

    if rank == 0:
        cpu = parity = aten = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, best, cpu_loss, kind = time_cpu(wl, 1, 3, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": "B=1 frame of %s (all %d cost calls + loss, fwd+bwd), best of 3 after 1 warm-up, %.2f s each; %s"
                             % (wl.name, 2 * wl.V * wl.T, best, "the unmodified reference's own functions on PyTorch-CPU"
                                if kind == "reference" else "PyTorch-CPU oracle port of the reference")}
            # the timed GPU step and the CPU arm compute the same thing: loss of the SAME B=1 inputs on both
            chk = HotPathStep(wl, dev, B=1, seed=1234, channels_last=(args.layout == "nhwc"))
            gpu_loss = float(chk.step().detach())
            parity = {"gpu_loss": gpu_loss, "cpu_loss": cpu_loss, "rel_err": abs(gpu_loss - cpu_loss) / max(abs(cpu_loss), 1e-30),
                      "inputs": "B=1, seed 1234 of %s on both arms; full-step gradients are compared in "
                                "tests/test_hotpath_parity_gpu.py" % wl.name}
            del chk
            aten = time_gpu_aten_reference(wl, B, dev)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "H": wl.H, "W": wl.W, "views": wl.V, "gru_steps": wl.T, "predictions": wl.n,
                       "batch_per_gpu": B, "global_batch": B * world, "feature_layout": args.layout,
                       "cuda_graph": not args.no_graph, "l2": "flushed between timed iterations (256 MiB write)",
                       "parallelism": "dp%d" % world, "host_cpus_rank0": (len(host_cpus) if host_cpus else None)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": step.h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "h2d_GBps_per_rank": step.h2d_bytes / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "host_batch": "what the data loader delivers per step, in pinned memory: target + source pictures as uint8 "
                                  "(converted on the device by drosfm_images_u8_to_f32 inside the timed step), float64 "
                                  "intrinsics, GT depth / poses for supervised workloads; feature maps, inverse depths and pose "
                                  "vectors are produced on the device by the networks and stay resident"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity, "gpu_aten_reference": aten,
        }
    else:
        out = None
    if world > 1 and not getattr(args, "keep_group", False):
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: isolated warp / feature-cost / SSIM-photometric sweep, GB/s against the HBM roofline
# ------------------------------------------------------------------------------------------------
def run_microbench(args):
    """Per-operator sweep over resolution (192x640 -> 384x1280) and source views (2, 4, 8) with batches sized so that the
    algorithmic bytes of ONE launch are >= 4x the 126 MB L2 -- the regime in which the HBM roofline applies.  Algorithmic
    bytes: SURVEY.md section 8(d).  One JSON line; `sweep` holds every measurement."""
    from dro_sfm_b200 import ops, synthetic as syn
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, peak_src = measured_peaks()
    L2 = 126e6
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    def add(op, H, W, V, B, ms, nbytes, launches):
        gbps = nbytes / (ms * 1e-3) / 1e9
        rows.append({"op": op, "H": H, "W": W, "views": V, "batch": B, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                     "GBps": round(gbps, 1), "frac": round(gbps / peak, 4), "launches": launches})

    shapes = [(192, 640), (256, 832), (320, 1024), (384, 1280)]
    views = [2, 4, 8]
    if args.quick:
        shapes, views = [(192, 640), (384, 1280)], [2, 8]
    g = syn.gen(5)
    for H, W in shapes:
        K = syn.intrinsics("kitti", 1, H, W)
        P, h, w = H * W, H // 8, W // 8
        # ---- warp: view_synthesis of one source view (kernels 1 + 2 fused), 28 B/px forward, 32 B/px backward
        B = min(512, max(1, int(4 * L2 / (28 * P)) + 1))
        img = syn.images(g, 2, H, W).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev)
        inv = syn.inv_depth(g, 2, H, W, 0.5, 80.0).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev).requires_grad_(True)
        pose = syn.pose_vec(g, B, "kitti").to(dev).requires_grad_(True)
        Kb = K.repeat(B, 1, 1).to(dev)
        out = [None]

        def f_warp():
            out[0] = ops.view_synthesis(img, inv, pose, Kb, None, 1.0, "zeros", inverse_depth=True)
        gout = torch.randn(B, 3, H, W, device=dev)
        add("warp_fwd", H, W, 1, B, timed(f_warp), 28 * P * B, 1)
        add("warp_bwd", H, W, 1, B, timed(lambda: torch.autograd.grad(out[0], (inv, pose), gout, retain_graph=True)), 32 * P * B, 1)
        del img, inv, gout, out
        for V in views:
            # ---- feature-metric cost (depth_cost_calc: V views, mean), C = 128 at 1/8 resolution
            C, p = 128, h * w
            fwd_b, bwd_b = (V + 2) * 4 * C + 4, (2 * V + 3) * 4 * C + 8
            B = min(1024, max(1, int(4 * L2 / (fwd_b * p)) + 1))
            cl = torch.channels_last
            fmap = syn.features(g, 2, C, h, w).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev).contiguous(memory_format=cl).requires_grad_(True)
            frefs = [syn.features(g, 2, C, h, w).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev).contiguous(memory_format=cl).requires_grad_(True)
                     for _ in range(V)]
            invl = syn.inv_depth(g, 2, h, w, 0.5, 80.0).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev).requires_grad_(True)
            poses = [syn.pose_vec(g, B, "kitti", 1.0 if v % 2 == 0 else -1.0).to(dev).requires_grad_(True) for v in range(V)]
            Kb = K.repeat(B, 1, 1).to(dev)
            gc = torch.randn(B, C, h, w, device=dev).contiguous(memory_format=cl)
            res = [None]

            def f_cost():
                res[0] = ops.feat_cost(invl, fmap, frefs, poses, Kb, None, 0.125, inverse_depth=True)
            add("feat_cost_fwd", H, W, V, B, timed(f_cost), fwd_b * p * B, 1)
            add("feat_cost_bwd", H, W, V, B, timed(lambda: torch.autograd.grad(res[0], [invl, fmap] + frefs + poses, gc, retain_graph=True)),
                bwd_b * p * B, 1)
            del fmap, frefs, gc, res
            # ---- photometric loss of one prediction (SSIM + L1, auto-mask, per-pixel min), V source views
            fwd_b, bwd_b = 16 + 12 * V, 20 + 12 * V
            B = min(256, max(1, int(4 * L2 / (fwd_b * P)) + 1))
            image = syn.images(g, 2, H, W).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev)
            ctx = [syn.images(g, 2, H, W).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev) for _ in range(V)]
            invs = [syn.inv_depth(g, 2, H, W, 0.5, 80.0).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev).requires_grad_(True)]
            pv = [[(syn.pose_vec(g, B, "kitti", 1.0 if v % 2 == 0 else -1.0) * 0.3).to(dev).requires_grad_(True)] for v in range(V)]
            Kb = K.repeat(B, 1, 1).to(dev)
            tot = [None]
            before = ops.L.lib().drosfm_launch_count()

            def f_photo():
                tot[0] = ops.photometric_loss(image, ctx, invs, Kb, Kb, pv, smooth_w=0.0)[0]
            f_photo()
            n_f = int(ops.L.lib().drosfm_launch_count() - before)
            leaves = invs + [x for row in pv for x in row]
            before = ops.L.lib().drosfm_launch_count()
            torch.autograd.grad(tot[0].sum(), leaves, retain_graph=True)
            n_b = int(ops.L.lib().drosfm_launch_count() - before)
            add("photometric_fwd", H, W, V, B, timed(f_photo), fwd_b * P * B, n_f)
            add("photometric_bwd", H, W, V, B, timed(lambda: torch.autograd.grad(tot[0].sum(), leaves, retain_graph=True)), bwd_b * P * B, n_b)
            del image, ctx, invs, pv, tot
            torch.cuda.empty_cache()
    best = {}
    for r in rows:
        if (r["H"], r["W"]) == shapes[-1]:
            best[r["op"]] = max(best.get(r["op"], 0.0), r["frac"])
    out = {"metric": "isolated warp / feature-cost / SSIM-photometric GB/s vs HBM roofline", "unit": "GB/s", "workload": "microbench",
           "config": {"workload": "microbench (BASELINE.json configs[4])", "shapes": shapes, "views": views,
                      "batch": "per row: algorithmic bytes of one launch >= 4 x 126 MB L2", "l2": "flushed between timed iterations"},
           "peak": peak, "peak_source": peak_src, "dtype": "f32", "data": "synthetic", "n_gpus": 1,
           "best_frac_at_%dx%d" % shapes[-1]: best, "sweep": rows}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train_kitti_mf_selfsup")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the YAML's batch_size)")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--log-every", type=int, default=10, help="all-reduce the logged loss every k steps (multi-GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="microbench: corner points of the sweep only")
    args = ap.parse_args()
    from dro_sfm_b200 import synthetic as syn
    if args.workload == "microbench":
        return run_microbench(args)
    if args.workload == "all":
        # every BASELINE.json training configuration through the same step; the headline line stays configs[1]
        args.keep_group = True
        results = {}
        for name in syn.WORKLOADS:
            args.no_cpu_baseline = name != "train_kitti_mf_selfsup"
            out = run_gpu(args, syn.WORKLOADS[name]) if args.impl != "reference" else None
            if out is not None:
                results[name] = out
        if results:
            head = dict(results["train_kitti_mf_selfsup"])
            head["other_workloads"] = {k: {"value": v["value"], "e2e": v["e2e"]["value"], "ms_per_step": v["ms_per_step"],
                                           "gpu_launches_per_step": v["gpu_launches_per_step"], "config": v["config"],
                                           "step_frac": v["roofline"]["step_frac"], "kernel_ms_per_step": v["roofline"]["kernel_ms_per_step"],
                                           "dominant": v["roofline"]["kernel"],
                                           "dominant_frac": v["roofline"]["frac"]} for k, v in results.items() if k != "train_kitti_mf_selfsup"}
            print(json.dumps(head))
        return
    wl = syn.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        out = run_gpu(args, wl)
        if out is not None:
            print(json.dumps(out))


if __name__ == "__main__":
    main()
