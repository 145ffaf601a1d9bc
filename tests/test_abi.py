"""CPU checks of the C-ABI boundary: the library builds/loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "drosfm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drosfm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for need in ("drosfm_warp_coords_fwd", "drosfm_warp_coords_bwd", "drosfm_view_synthesis_fwd", "drosfm_view_synthesis_bwd",
                 "drosfm_feat_cost_fwd", "drosfm_feat_cost_bwd", "drosfm_photometric_fwd", "drosfm_photometric_bwd"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from dro_sfm_b200 import _lib
    from dro_sfm_b200.build import build
    build()
    handle = ctypes.CDLL(_lib.SO_PATH)
    missing = [n for n in declared_symbols() if not hasattr(handle, n)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().drosfm_version() == _lib.ABI_VERSION
    assert _lib.lib().drosfm_ws_bytes(4) % (4 * _lib.SLOT_BYTES) == 0


def test_structs_match_header_layout():
    from dro_sfm_b200 import _lib
    # drosfm_cams_t: 2 pointers, int32, 2 floats, (pad), 2 pointers, int32 (+pad) on LP64
    assert ctypes.sizeof(_lib.Cams) == 56
    assert _lib.Cams.Twc.offset == 32 and _lib.Cams.pose_kind.offset == 48
    assert ctypes.sizeof(_lib.PhotoOpts) == 40 and _lib.PhotoOpts.clip_scratch.offset == 32
    assert ctypes.sizeof(_lib.CostJob) == 56 and _lib.CostJob.cost.offset == 40 and _lib.CostJob.disp_range.offset == 52
    assert ctypes.sizeof(_lib.CostJobGrads) == 48 and _lib.CostJobGrads.flags.offset == 40


def test_no_cpu_fallback():
    import torch
    from dro_sfm_b200 import ops
    K = torch.eye(3).repeat(1, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.warp_coords(torch.ones(1, 1, 4, 4), torch.eye(4).repeat(1, 1, 1), K)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dro_sfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_relayout_of_host_tensors_is_torchs_copy():
    """ops.relayout only calls the library for dense fp32 CUDA tensors; everything else takes torch's strided copy."""
    import torch
    from dro_sfm_b200 import ops, _lib
    x = torch.randn(2, 8, 3, 5)
    cl = ops.relayout(x, _lib.NHWC)
    assert cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(cl, x)
    assert ops.relayout(cl, _lib.NHWC) is cl
    back = ops.relayout(cl, _lib.NCHW)
    assert back.is_contiguous() and torch.equal(back, x)


def test_graphed_step_tree_helpers_and_plain_split():
    """Host logic that needs no GPU: the batch-pytree helpers of graphs.GraphedStep; split_feature_maps is a plain
    torch.split (no layout conversion, no kernel) for tensors that are not CUDA float32 NCHW maps."""
    import torch
    from dro_sfm_b200 import graphs
    from dro_sfm_b200.networks import cost as cost_mod
    batch = {"rgb": torch.zeros(2, 3), "ctx": [torch.zeros(2), torch.ones(2)], "idx": 7, "t": (torch.zeros(1),)}
    clone = graphs._tree_map(lambda t: t.clone(), batch)
    assert clone["idx"] == 7 and clone["ctx"][1] is not batch["ctx"][1] and isinstance(clone["t"], tuple)
    new = {"rgb": torch.full((2, 3), 5.0), "ctx": [torch.full((2,), 2.0), torch.full((2,), 3.0)], "idx": 8, "t": (torch.ones(1),)}
    graphs._tree_copy_(clone, new)
    assert float(clone["rgb"].sum()) == 30.0 and float(clone["ctx"][1][0]) == 3.0 and float(clone["t"][0]) == 1.0
    stacked = torch.randn(6, 8, 4, 5)
    parts = cost_mod.split_feature_maps(stacked, [2, 2, 2])
    assert len(parts) == 3 and all(torch.equal(p, q) for p, q in zip(parts, torch.split(stacked, 2)))
