"""CPU oracle for the S1->S2 conditional-UNet DDIM/DDPM sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The shipped path (``s1s2_b200`` + ``libs1s2_b200.so``)
never imports this package and has no CPU fallback.

Parity status: the reference repository holds no tests, golden vectors or
known-answer fixtures for this path, so the
oracle is pinned against *outputs of the reference itself run in the build
container*: ``oracle/gen_golden.py`` imports the reference's own files from
``/root/reference`` by path, runs them on seeded inputs and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors.  The overlap-blend stitch
does not exist in the reference at all and is therefore "parity unpinned".

Each function cites the reference file:line it restates (paths relative to the
reference repository root).
"""
from . import schedule, unet, samplers, patch, metrics  # noqa: F401
