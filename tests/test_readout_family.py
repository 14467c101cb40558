"""CPU tier: sampling (incl. MeasureX / MeasureY and a reduced register) and the readout family through the host API
on the torch model of the device primitives; the same bodies run on cuda:0 in tests/test_gpu_readout.py."""
import pytest
import torch

import readout_cases as rc
from cpu_prims import CpuPrims
from MPDOSimulator import _engine


@pytest.fixture(autouse=True)
def cpu_model_prims():
    _engine._PRIMS = CpuPrims()   # CPU model of the device primitives, injected by the test
    yield
    _engine._PRIMS = None


def test_sampling_chi_square_cpu_model():
    rc.check_sampling(torch.complex128, 'cpu', shots=2048)


def test_readout_family_cpu_model():
    rc.check_readout_family(torch.complex128, 'cpu', 1e-10)


def test_gate_parameters_that_require_grad_are_refused():
    import MPDOSimulator as Simulator
    c = Simulator.TensorCircuit(qn=2, ideal=True, dtype=torch.complex128, device='cpu')
    theta = torch.tensor(0.3, requires_grad=True)
    with pytest.raises(NotImplementedError):
        c.rz(theta, [0])
    c.rz(theta.detach(), [0])     # the documented way


def test_adaptive_ranks_cpu_model():
    import extension_cases as ec
    ec.check_adaptive_ranks(torch.complex128, 'cpu', 1e-9)
    ec.check_adaptive_ranks(torch.complex128, 'cpu', 1e-9, err=5e-3, chi=12, kappa=5)


def test_long_range_gates_cpu_model():
    import extension_cases as ec
    ec.check_long_range_gates(torch.complex128, 'cpu', 1e-9)


def test_chi_formats_and_cp_tomography_cpu_model(tmp_path):
    import extension_cases as ec
    ec.check_chi_formats_and_cp(torch.complex128, 'cpu', 1e-9, str(tmp_path))
