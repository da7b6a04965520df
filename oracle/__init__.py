"""oracle/ — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this directory is imported by the product package ``modelcompression_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of ``bench.py`` may use it, and only
as the checker or as the timed CPU baseline.

Pinning: every function here was checked against the UNMODIFIED reference (imported from /root/reference with the
shims in ``ref_shim.py``) by ``make_golden.py``, which also wrote the fixtures in ``tests/golden/``; the CPU test
suite re-checks the oracle against those fixtures.  Third-party arithmetic the reference relies on (NumPy
``np.percentile`` / ``.sum``; PyTorch conv/BN/sigmoid/softmax/sort) is not pinned by the reference itself (no
requirements file): the pin is "this container's NumPy 2.3.5 / torch 2.11.0" (SURVEY.md §8c).
"""
