"""The oracle port (oracle/torch_oracle.py, what bench.py's CPU arm and every parity test check against) against the
UNMODIFIED reference on a whole hot-path step -- every cost call fwd+bwd + the loss fwd+bwd -- on fresh seeded inputs
(not the committed fixtures): loss and all leaf gradients must be BIT-EQUAL on the CPU.  Runs wherever the reference is
importable (/root/reference in the build container, baseline/_ref on the GPU box); skipped otherwise."""
import numpy as np
import pytest
import torch

import bench
from oracle import reference
from dro_sfm_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not reference.available(), reason="reference tree not present")

SMALL = {
    "selfsup": syn.Workload("small_selfsup", 64, 96, 2, 2, 2, 1, 3, 0.5, 80.0, False, "kitti"),
    "sup": syn.Workload("small_sup", 48, 64, 2, 2, 2, 2, 2, 0.2, 10.0, True, "scannet"),
}


@pytest.mark.parametrize("kind", list(SMALL))
def test_oracle_port_is_bit_equal_to_the_reference(kind):
    wl = SMALL[kind]
    batch = syn.hot_path_batch(wl, seed=4321, C=32, B=2)
    threads = torch.get_num_threads()
    loss_o, g_o = bench.cpu_step(wl, batch, threads, torch.float32, return_grads=True)
    loss_r, g_r = reference.hot_path_step(wl, batch, "cpu", torch.float32, return_grads=True)
    assert np.float32(loss_o).tobytes() == np.float32(loss_r).tobytes(), (loss_o, loss_r)
    assert len(g_o) == len(g_r)
    for k, (a, b) in enumerate(zip(g_o, g_r)):
        assert a is not None and b is not None, k
        assert torch.equal(a, b), "leaf %d: oracle port and reference differ (max |d| = %.3e)" % (k, float((a - b).abs().max()))
