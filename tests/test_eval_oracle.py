"""Evaluation-path oracle (oracle/eval_oracle.py) against the reference's own outputs (tests/golden/eval.npz)."""
import numpy as np
import pytest

from oracle import eval_oracle as eo

CROPS = {0: "", 1: "garg", 2: "eigen_nyu"}


@pytest.mark.parametrize("tag", ["pp_a", "pp_b"])
@pytest.mark.parametrize("method", ["mean", "max", "min"])
def test_post_process_bit_exact(golden, tag, method):
    G = golden("eval")
    out = eo.post_process_inv_depth(G[tag + "_inv"], G[tag + "_inv_flipped"], method)
    assert np.array_equal(out.view(np.uint32), G["%s_%s" % (tag, method)].view(np.uint32))


def test_post_process_unknown_method():
    with pytest.raises(ValueError):
        eo.post_process_inv_depth(np.ones((1, 1, 2, 2), np.float32), np.ones((1, 1, 2, 2), np.float32), "median")


@pytest.mark.parametrize("tag", ["m_garg", "m_none", "m_nyu"])
@pytest.mark.parametrize("scale", [1, 0])
def test_depth_metrics(golden, tag, scale):
    G = golden("eval")
    lo, hi, crop = G[tag + "_cfg"]
    out = eo.compute_depth_metrics(CROPS[int(crop)], lo, hi, G[tag + "_gt"], G[tag + "_pred"], bool(scale))
    np.testing.assert_allclose(out, G["%s_scale%d" % (tag, scale)], rtol=1e-5, atol=1e-7)   # fp32 means; 1e-5 = north_star


def test_linspace_matches_torch():
    """The scalar two-sided rule; torch's vectorised CPU kernel builds each SIMD group from its first element, so the
    host bits may differ by one ulp (the CUDA kernel follows the scalar rule: tests/test_eval_gpu.py)."""
    import torch
    for W in (1, 2, 7, 64, 641, 1241):
        ref = torch.linspace(0., 1., W).numpy()
        np.testing.assert_allclose(eo.linspace01(W), ref, rtol=2.4e-7, atol=0)


@pytest.mark.parametrize("tag", ["pp_a", "pp_b"])
def test_torch_restatement_post_process(golden, tag):
    import torch
    G = golden("eval")
    for method in ("mean", "max", "min"):
        out = eo.post_process_inv_depth_torch(torch.from_numpy(G[tag + "_inv"]), torch.from_numpy(G[tag + "_inv_flipped"]), method)
        assert np.array_equal(out.numpy().view(np.uint32), G["%s_%s" % (tag, method)].view(np.uint32))


@pytest.mark.parametrize("tag", ["m_garg", "m_none", "m_nyu"])
def test_torch_restatement_metrics(golden, tag):
    import torch
    G = golden("eval")
    lo, hi, crop = G[tag + "_cfg"]
    for scale in (1, 0):
        out = eo.compute_depth_metrics_torch(CROPS[int(crop)], float(lo), float(hi), torch.from_numpy(G[tag + "_gt"]),
                                             torch.from_numpy(G[tag + "_pred"]), bool(scale))
        assert np.array_equal(out.numpy(), G["%s_scale%d" % (tag, scale)])
