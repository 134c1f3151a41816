import pytest

from edge_suite import CASES, FATAL, run_case, run_fatal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_edge_case_cuda(cuda_lib, case):
    run_case(cuda_lib, case)


@pytest.mark.parametrize("case", FATAL, ids=[c["id"] for c in FATAL])
def test_fatal_input_cuda(cuda_lib, case):
    run_fatal(cuda_lib, case)
