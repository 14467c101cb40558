"""-m gpu: SURVEY 8f row 3 through the native engine - adaptive ranks from `max_truncation_err` (mpdo_bond_svd_step /
mpdo_kappa_truncate with the relative rule evaluated on the device) and long-range two-qubit gates."""
import pytest
import torch

import extension_cases as ec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('dtype,tol', [(torch.complex128, 1e-9), (torch.complex64, 3e-4)])
def test_adaptive_ranks_native(cuda_prims, dtype, tol):
    from MPDOSimulator import _engine
    from MPDOSimulator._engine.native import NativeEngine
    assert isinstance(_engine.engine_for(dtype), NativeEngine)
    ec.check_adaptive_ranks(dtype, 'cuda:0', tol)
    ec.check_adaptive_ranks(dtype, 'cuda:0', tol, err=5e-3, chi=12, kappa=5)


@pytest.mark.parametrize('dtype,tol', [(torch.complex128, 1e-9), (torch.complex64, 1e-4)])
def test_long_range_gates(cuda_prims, dtype, tol):
    ec.check_long_range_gates(dtype, 'cuda:0', tol)


@pytest.mark.parametrize('dtype,tol', [(torch.complex128, 1e-9), (torch.complex64, 1e-4)])
def test_chi_formats_and_cp_tomography(cuda_prims, tmp_path, dtype, tol):
    """.npz ('chi') and .mat ('exp') process matrices, CPEXP gates from chiFileDict['CP']."""
    ec.check_chi_formats_and_cp(dtype, 'cuda:0', tol, str(tmp_path))
