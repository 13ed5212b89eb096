"""ORACLE — test infrastructure only.

CPU restatements (plain fp32 PyTorch / numpy) of the reference's algorithms on the hot path, each
function citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import this package; the product
(medsegpretrainimagenet_b200/) never does and fails loudly without its CUDA library.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
the reference itself — imported from /root/reference in the build container by
oracle/reference_harness.py — (a) live in tests/test_oracle_vs_reference.py and (b) through the
committed fixtures tests/golden/*.pt minted by tools/make_golden.py.
"""
