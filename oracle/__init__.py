"""TEST INFRASTRUCTURE - CPU oracle for the WDPM redistribution path.

Nothing under wdpm_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs do.
"""
