"""CPU oracle for the img2latex batched-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline -- never as a fallback for the CUDA path.

Parity status: the reference ships no golden vectors for this path
(SURVEY.md F12), so the oracle is pinned against the *live reference modules*
executed in the build container (``oracle/ref_shim.py`` +
``tests/golden/make_golden.py``); the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.
"""
from .port import *  # noqa: F401,F403
from . import metrics  # noqa: F401  (evaluate metrics restated; pinned against the live reference)
from . import resize  # noqa: F401  (Pillow resampling restated; pinned against live Pillow)
