"""oracle/ -- CPU restatement of the reference's algorithm for the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package,
and only as the checker or the reported CPU baseline: the product (dedalus_b200/) never imports it and has no CPU
fallback.

Parity status: PINNED.  Every module here is checked by tests/test_oracle.py against fixtures produced by the
UNMODIFIED reference (Dedalus v3.0.5 at /root/reference, run single-process under tests/golden/ref_shim.py by
tests/golden/make_golden.py): transform inputs/outputs of the reference's scipy and matrix plugins, the
reference's per-pencil M/L matrices in natural ordering, and K-step states of KdV-Burgers, 2-D / 3-D
Rayleigh-Benard and the S2 shallow-water problem (sphere_oracle.py: complex per-m formulation with dense solves).  The reference is Python and cannot travel to the GPU box, so the vectors are committed under
tests/golden/ together with the generating script.
"""
