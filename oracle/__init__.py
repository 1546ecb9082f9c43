"""CPU oracle for the ESPER diffusion-map hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`manifoldem_esper_b200/`) may import this; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and only as the checker / CPU baseline.
"""
