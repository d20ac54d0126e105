"""Import the *real* reference (`/root/reference`, read-only) with `sys.modules` stubs for the
packages it imports but does not exercise on the hot path.

Only used by `tests/golden/make_golden.py` in the build container.  Nothing under `tests/` that
runs with `-m gpu`, `bench.py` or `smoke()` may import this module: `/root/reference` does not
exist on the GPU box.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_stage  # noqa: E402

REFERENCE_ROOT = ref_stage.REFERENCE_ROOT


def install() -> bool:
    """Make `import kod...` work from /root/reference.  Returns False when the reference tree is absent."""
    return ref_stage.install(REFERENCE_ROOT)
