"""CPU oracle for the dro-sfm dense depth-pose warping hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dro_sfm_b200/`` may import this
package: it is the checker for the CUDA path, never the thing measured or
shipped.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: the reference (xyang9527/dro-sfm) ships no tests, golden vectors
or fixtures ("parity unpinned" by the reference itself).
The oracle is therefore pinned against outputs of the reference's own Python
code, imported unmodified from ``/root/reference`` in the build container by
``tests/golden/make_golden.py`` and committed as fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures.
"""
from .torch_oracle import *  # noqa: F401,F403
