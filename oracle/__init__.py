"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

`oracle/` holds the CPU restatement of the reference's algorithm for the hot
path (collocation constraint vector + dense forward-difference Jacobian).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker / the timed CPU
baseline.  `opengoddard_b200` must never import this package.
"""
