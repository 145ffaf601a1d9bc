"""Import the UNMODIFIED reference (build container: /root/reference; GPU box: the staged copy baseline/_ref).

Thin alias of ``oracle/reference.py`` (shims, loader), kept for ``make_golden.py`` and the tests that import it by
this name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reference as _ref  # noqa: E402

REF_ROOT = _ref.root() or "/root/reference"
available = _ref.available
load = _ref.load
