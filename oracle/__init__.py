"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the coarse-registration forward path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this
package; nothing under `gaussreg_b200/` does.
"""
