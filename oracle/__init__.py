"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithms on the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / the CPU baseline.  The product
(cet_pick_b200/) never imports it and has no CPU fallback.

Parity pin: the reference (nextpyp/cet_pick) ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
executed in the build container by tests/golden/make_golden.py (committed) and stored as
tests/golden/*.npz; tests/test_oracle_golden.py checks the oracle against every one of them.
"""
