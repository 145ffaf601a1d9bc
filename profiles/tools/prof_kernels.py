"""Runs one eager hot-path step of the bench workload (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from dro_sfm_b200 import synthetic as syn
from dro_sfm_b200.hotpath import HotPathStep
wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "train_kitti_mf_selfsup"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl.B
layout = sys.argv[3] if len(sys.argv) > 3 else "nchw"
step = HotPathStep(wl, "cuda:0", B=B, channels_last=(layout == "nhwc"))
for _ in range(2):
    step.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
