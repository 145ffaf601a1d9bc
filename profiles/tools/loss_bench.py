"""Loss phase only (photometric + smoothness fwd+bwd) of the benchmark workload: graph replay time + per-call event times."""
import os, sys, json
import torch
sys.path.insert(0, '.')
from dro_sfm_b200 import synthetic as syn, _lib as L, ops
from dro_sfm_b200.hotpath import HotPathStep
from dro_sfm_b200.geometry import Pose

wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train_kitti_mf_selfsup"]
step = HotPathStep(wl, "cuda:0")

def loss_only():
    poses = [[Pose.from_vec(p, 'euler') for p in row] for row in step.poses]
    out = step.loss_mod(step.image, step.context, step.inv_depths, step.K, step.K, poses)
    out['loss'].backward()

s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step.zero_grads(); loss_only()
torch.cuda.current_stream().wait_stream(s)
step.zero_grads()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    loss_only()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(30):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
# per call, eager, serial
ops.OVERLAP = "serial"
step.zero_grads(); loss_only(); torch.cuda.synchronize()
L.profile_begin()
for _ in range(5):
    flush.zero_(); torch.cuda._sleep(20_000_000)
    step.zero_grads(); loss_only(); torch.cuda.synchronize()
recs = L.profile_end()
per = {}
for name, a, ms in recs:
    per.setdefault(name.replace("drosfm_", ""), []).append(ms * 1e3)
print(json.dumps({"so": os.path.basename(L.SO_PATH), "workload": wl.name, "loss_graph_us": round(ts[len(ts) // 2] * 1e3, 1),
                  "calls_us": {k: round(sum(v) / len(v), 1) for k, v in per.items()}}), flush=True)
