"""CPU oracle for the Instant-angelo hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under instant_angelo_b200/ imports this package.  Allowed importers: tests/,
__graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs.
Parity status: the third-party boundary (tiny-cuda-nn, nerfacc 0.3.3) is PARITY UNPINNED upstream
(no tests or golden vectors exist in the reference); the reference's own Python (VolumeSDF,
get_alpha, forward_, VanillaMLP, colour heads) is pinned by tests/golden/make_golden.py.
"""
