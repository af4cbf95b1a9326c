"""CPU oracle for the streaming-simulator hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / the timed CPU baseline.  The product package
(``mansy_immersivevideostreaming_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Parity status: PINNED.  ``oracle/sim_oracle.py`` was checked in this container against the
unmodified reference (``/root/reference`` imported through ``oracle/ref_loader.py``) on the
shipped Jin2022/4G data and on synthetic datasets written in the reference's own on-disk
formats, and against the reference's shipped ground-truth tile masks; the vectors that came out
of those runs are committed under ``tests/golden/`` together with the generating script
(``oracle/make_golden.py``).  The reference itself holds no tests or golden vectors
(SURVEY.md section 4).
"""
