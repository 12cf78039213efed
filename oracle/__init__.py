"""ORACLE — test infrastructure only.

CPU restatements of the reference algorithms used to check the CUDA product.  Nothing in
`climt_b200/` may import this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.
"""
