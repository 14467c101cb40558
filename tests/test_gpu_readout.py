"""-m gpu: SURVEY 8f rows 1-2 on the device - sampling (chi-square against the oracle's dense rho, incl. MeasureX /
MeasureY rotations, a reduced register, randomSample and the sequential sampler) and the readout family
(trace_rho_rho, trace_composited_rho*, dense multi-qubit expect, pauli_expect, cal_dm(reduced), cal_fidelity,
density2prob, bitstring probabilities) through libmpdo_b200.so."""
import pytest
import torch

import readout_cases as rc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype', [torch.complex128, torch.complex64])
def test_sampling_chi_square(cuda_prims, dtype):
    report = rc.check_sampling(dtype, 'cuda:0', shots=4096)
    print({k: (round(v[0], 1), v[1]) for k, v in report.items()})


@pytest.mark.parametrize('dtype,tol', [(torch.complex128, 1e-10), (torch.complex64, 2e-5)])
def test_readout_family(cuda_prims, dtype, tol):
    out = rc.check_readout_family(dtype, 'cuda:0', tol)
    print({k: f'{v:.1e}' for k, v in out.items()})
