"""CPU oracle for the TurbDiff denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline - never as the thing shipped.  The product package
(``generative-turbulence_b200/turbdiff_b200``) never imports this package.

The oracle is a restatement, in plain functional PyTorch on CPU tensors (fp32 or
fp64), of the algorithms in the reference's ``turbdiff/models/ddpm.py``,
``models/utils.py``, ``models/attention.py``, ``models/cell_type_embeddings.py``,
``models/normalization.py`` and ``data/ofles.py::grid_embedding``.  Each function
cites the reference lines it follows.

Parity pin: the reference ships no tests or golden vectors,
so the pin is made by running the *unmodified reference* in the build container
(``tests/golden/make_golden.py``, which imports ``/root/reference`` through a
stub-module shim) and committing its outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures.
"""
