"""CPU oracle for the multi-modal-qg hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``multi-modal-qg_b200/``) imports this package.
The only callers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and they use it as
the checker / the CPU comparator, never as the thing measured or shipped.

Parity status: **pinned against live reference output**.  The reference ships no
tests, golden vectors or known-answer files, so the oracle is
pinned the other way the task allows: ``tests/golden/make_golden.py`` imports the
reference's own ``model/encoder.py`` / ``model/decoder.py`` from /root/reference,
drives them with the per-sample loop of ``train.py:143-177`` and ``train.py:74-110``
and stores the inputs, weights, loss, gradients and greedy tokens as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds the oracle to those.
"""
