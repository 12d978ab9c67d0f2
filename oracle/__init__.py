"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Vec2Wav generator forward pass.

Nothing in the product path (wavthruvec_pytorch_b200/) imports this package.  It is used by
tests/, by __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
--impl reference leg as the thing timed on the host cores.

Parity status: PINNED against the reference itself.  The reference ships no tests, golden
vectors or checkpoints, so `make_golden.py` imports the reference's own
`vec2wav/models.py` in the build container, runs it on seeded fixtures and commits inputs and
outputs under tests/golden/; the restatements here are checked against those files
(tests/test_oracle.py).
"""
