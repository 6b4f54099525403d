"""CPU: how tightly the reference algorithm pins an AIC entry at all.  The oracle is run twice on the same case, once
with glibc's log/atan2 (what a gfortran build of the reference links) and once with those two functions evaluated in
binary128 and rounded once.  Their difference is the noise any conforming libm (CUDA's included) puts on an entry; it
calibrates the metric of tests/test_gpu_parity.py."""
import numpy as np

import fixtures
import oracle_binding as ob


def test_libm_noise_floor_of_the_reference_algorithm():
    case, _, _ = fixtures.make_case("test_13")  # supersonic half wing, 446 unknowns
    A, _, S = ob.assemble(case, with_scale=True)
    ob.lib().orc_set_exact_libm(1)
    try:
        A2, _ = ob.assemble(case)
    finally:
        ob.lib().orc_set_exact_libm(0)
    d = np.abs(A - A2)
    nz = A != 0
    plain = d[nz] / np.abs(A[nz])
    # plain relative agreement to 1e-12 is NOT a property of the reference algorithm (cancelling atan2 sums) ...
    assert plain.max() > 1e-12
    # ... while relative to the summed terms the two agree to rounding
    assert (d / np.where(S > 0, S, 1.0)).max() < 1e-15
    assert ((S >= np.abs(A) * (1 - 1e-12)) | (S == 0)).all()
    case.close()
