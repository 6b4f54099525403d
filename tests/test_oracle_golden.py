"""CPU: host setup + oracle (assemble + solve) + host post-processing reproduce the golden tuples
of the reference's own regression suite (test/test_machline.py).  This is what pins the oracle."""
import pytest

import fixtures
import oracle_binding as ob


@pytest.mark.parametrize("name", fixtures.golden_case_names())
def test_oracle_reproduces_reference_golden(name):
    case, expect, tol = fixtures.make_case(name)
    res, info, _ = ob.run_case(case)
    fixtures.check_tuple(res, expect, tol)
    case.close()
