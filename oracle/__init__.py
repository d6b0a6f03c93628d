"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's SpectralMixingLayer / Wirtinger filter hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the timed CPU baseline. The product package
(tensor-cuda-fft-_b200/) never imports it and has no CPU fallback.

Parity pinning: the reference's own tests hold no golden vectors for this path, so the
restatements here are pinned against outputs of the UNMODIFIED reference module executed in the build
container (oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
