import os, sys, json
sys.path.insert(0, ".")
import torch
import importlib.util
spec = importlib.util.spec_from_file_location("cb", "profiles/tools/cost_bench_lib.py"); cb = importlib.util.module_from_spec(spec); spec.loader.exec_module(cb)
from dro_sfm_b200 import _lib as L
for wl in ("train_kitti_mf_selfsup", "train_scannet_mf_selfsup_view5"):
    print(json.dumps({"so": os.path.basename(L.SO_PATH), "workload": wl, "cost_phase_us": round(cb.bench(wl, True, 0), 1)}), flush=True)
