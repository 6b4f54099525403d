"""CPU: host setup + oracle (assemble + solve) + host post-processing reproduce the golden tuples
of the reference's own regression suite (test/test_machline.py).  This is what pins the oracle."""
import pytest

import fixtures
import oracle_binding as ob


@pytest.mark.parametrize("name", fixtures.golden_case_names())
def test_oracle_reproduces_reference_golden(name):
    case, expect, tol = fixtures.make_case(name)
    res, info, _ = ob.run_case(case)
    fixtures.check_tuple(res, expect, tol, slack=fixtures.oracle_slack(name))
    case.close()


@pytest.mark.parametrize("name", fixtures.comparison_names())
def test_oracle_reproduces_reference_comparison(name):
    """Reference tests that compare two runs with each other (test 10: the full wing against the mirrored half wing)."""
    a, b, tol = fixtures.comparison_cases(name)
    ra, _, _ = ob.run_case(a)
    rb, _, _ = ob.run_case(b)
    ta = [ra.C_p_max, ra.C_p_min, *[float(v) for v in ra.C_F]]
    tb = [rb.C_p_max, rb.C_p_min, *[float(v) for v in rb.C_F]]
    for x, y, t in zip(ta, tb, tol):
        assert abs(x - y) < t, (ta, tb)
    a.close()
    b.close()
