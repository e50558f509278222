"""CPU oracle for the EDMP guided-sampling hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (edmp_b200/) never does; it fails loudly
when its CUDA library is missing instead of falling back to anything here.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so every function here is pinned against the *reference itself*
executed under import shims in the build container (oracle/ref_shim.py,
oracle/make_golden.py); the resulting fixtures are committed under tests/golden/.
"""
