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


def test_a_faithfully_rounded_libm_moves_a_quarter_of_the_entries_beyond_1e12():
    """Mode 2: log/atan2 correctly rounded, then moved by one ulp in 3/8 of the calls -- a libm with < 1 ulp error, like
    CUDA's.  On the mirrored half wing this alone puts ~30 % of the entries beyond plain-relative 1e-12 (max ~4e-5):
    the statistics the GPU shows against the oracle are those of its libm, not of its arithmetic."""
    case, _, _ = fixtures.make_case("test_05")
    A, _, S = ob.assemble(case, with_scale=True)
    ob.lib().orc_set_exact_libm(2)
    try:
        A2, _ = ob.assemble(case)
    finally:
        ob.lib().orc_set_exact_libm(0)
    d = np.abs(A - A2)
    nz = A != 0
    plain = d[nz] / np.abs(A[nz])
    assert (plain > 1e-12).mean() > 0.1 and plain.max() > 1e-6
    assert (d / np.where(S > 0, S, 1.0)).max() < 2e-15
    case.close()
