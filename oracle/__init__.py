"""CPU oracle for the KuroSiwo training hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this package, and only as the checker or the reported CPU baseline: nothing under
`kurosiwo_b200/` (the product) imports it, and the product has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the reference
modules from /root/reference (in the build container), runs them on seeded inputs and commits the
results under `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against them.
"""
