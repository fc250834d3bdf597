"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DeepAVFusion pre-training hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it, and only as the checker / the CPU arm that is timed beside the GPU path.
The product path (``deepavfusion_b200``) never imports this package and fails loudly when the
CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference itself*: ``oracle/make_golden.py`` imports
``/root/reference/models/*.py`` in the build container (through ``oracle/timm_shim``, a
restatement of the five ``timm==0.9.2`` classes the reference needs), checks this restatement
against it and writes the fixtures in ``tests/golden/``.
"""
