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


def test_barrier_timeout_fails_loudly(cuda_prims):
    """The device-wide barrier of the persistent Cholesky / Jacobi kernels is bounded: a CTA that waits too long
    traps, and the library call reports a CUDA error instead of returning a half-updated factorisation. Run in a
    throw-away process (a trap leaves the CUDA context unusable)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import torch; torch.cuda.init()\n"
        "from MPDOSimulator._engine import lib\n"
        "h = lib.load()\n"
        "rc = h.mpdo_debug_barrier_timeout(None)\n"
        "print('rc', rc, (h.mpdo_last_error() or b'').decode())\n"
        "sys.exit(0 if rc != 0 else 7)\n" % (root, os.path.join(root, 'tomography-assisted-mpdo-qcircuit_b200')))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    print(out.stdout.strip(), out.stderr.strip()[-200:])
    assert out.returncode == 0, 'a barrier time-out must surface as an error'
