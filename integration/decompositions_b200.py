"""Seam S1 of the reference, bound to libmpdo_b200.so: drop-in `svd` / `qr` with the signatures of the
TensorNetwork-pytorch backend file the reference asks users to overwrite (README.md:17,
decompositions.py:51-57,149-154), for CUDA tensors. A maintainer who keeps the reference tree copies this file
next to `decompositions.py` and routes CUDA tensors to it; tests/test_gpu_integration_shim.py executes it against
the oracle's restatement of decompositions.py.

It keeps the reference's FORMULATION (whole matrices are decomposed); the speed of this repository comes from the
reformulated path (MPDOSimulator/_engine, csrc/engine.cu), so the supported route is the drop-in package.
Arithmetic is complex128 whatever the input dtype (results are cast back), which is at least as accurate as the
reference's LAPACK call in the input precision."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.normpath(os.path.join(_HERE, '..', 'tomography-assisted-mpdo-qcircuit_b200', 'lib', 'libmpdo_b200.so'))
lib = C.CDLL(os.environ.get('MPDO_B200_LIB', _LIB))
lib.mpdo_decompose_rows.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int, C.c_double, C.c_double, C.c_int,
                                                                     C.c_void_p]
lib.mpdo_rank_rule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p]
lib.mpdo_chol_psd_scratch_bytes.restype = C.c_int64
lib.mpdo_chol_psd_scratch_bytes.argtypes = [C.c_int, C.c_int]
lib.mpdo_chol_psd.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_void_p]
lib.mpdo_last_error.restype = C.c_char_p


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f'{what} failed (rc={rc}): {(lib.mpdo_last_error() or b"").decode()}')


def _svd_rows(M):
    """M [n, m] complex128, n <= m  ->  (Uh [n,n], s [n] descending, Wh [n,m]) with M = Uh^h diag(s) Wh."""
    n, m = M.shape
    Y = torch.empty(n, m + n, dtype=torch.complex128, device=M.device)
    work = torch.zeros(48, dtype=torch.int32, device=M.device)
    s = torch.empty(n, dtype=torch.float64, device=M.device)
    Wh = torch.empty(n, m, dtype=torch.complex128, device=M.device)
    Uh = torch.empty(n, n, dtype=torch.complex128, device=M.device)
    _check(lib.mpdo_decompose_rows(1, n, m, M.data_ptr(), Y.data_ptr(), work.data_ptr(), s.data_ptr(), Wh.data_ptr(),
                                   Uh.data_ptr(), 1, 1e-300, 1e-15, 30, _stream(M)), 'mpdo_decompose_rows')
    return Uh, s, Wh


def svd(torch_mod, tensor, pivot_axis, max_singular_values=None, max_truncation_error=None, relative=False):
    """decompositions.svd (reference decompositions.py:51-146), always on its full-SVD branch: one-sided Jacobi on
    the device (mpdo_decompose_rows) + the reference's rank rule evaluated on the device (mpdo_rank_rule)."""
    left, right = tensor.shape[:pivot_axis], tensor.shape[pivot_axis:]
    M = tensor.reshape(left.numel(), right.numel()).to(torch.complex128)
    rows, cols = M.shape
    if rows <= cols:
        Uh, s, Wh = _svd_rows(M.contiguous())
        u, vh = Uh.mH, Wh                                       # M = u diag(s) vh
    else:                                                       # decompose M^h and swap the roles
        Uh, s, Wh = _svd_rows(M.mH.contiguous())                # M^h = Uh^h diag(s) Wh
        u, vh = Wh.mH, Uh
    n = s.numel()
    keep = torch.empty(1, dtype=torch.int32, device=M.device)
    cap = n if max_singular_values is None else min(int(max_singular_values), n)
    sv = s.clone()
    _check(lib.mpdo_rank_rule(1, n, sv.data_ptr(), n, 0, cap,
                              -1.0 if max_truncation_error is None else float(max_truncation_error), int(relative),
                              int(tensor.dtype == torch.complex64), keep.data_ptr(), 0, _stream(M)), 'mpdo_rank_rule')
    k = int(keep.item())
    dt = tensor.dtype
    return (u[:, :k].reshape(*left, k).to(dt), s[:k].to(dt), vh[:k].reshape(k, *right).to(dt), s[k:].to(dt))


def qr(torch_mod, tensor, pivot_axis, non_negative_diagonal=False):
    """decompositions.qr (reference decompositions.py:149-195) as Cholesky-QR: q is an isometry on the numerical
    range of the matrix and tensor = q . r; r is NOT triangular (its rows come in pivot order), which the
    left-to-right sweep does not need (TNNOptimizer.py:98-106 only contracts r into the next site)."""
    left, right = tensor.shape[:pivot_axis], tensor.shape[pivot_axis:]
    A = tensor.reshape(left.numel(), right.numel()).to(torch.complex128)
    n = A.shape[1]
    G = (A.mH @ A).contiguous()          # the package computes this with mpdo_contract (hermitian = 1) on the view
    Lh, Linv = torch.empty_like(G), torch.zeros_like(G)
    scratch = torch.empty(max(int(lib.mpdo_chol_psd_scratch_bytes(1, n)), 256), dtype=torch.uint8, device=G.device)
    _check(lib.mpdo_chol_psd(1, n, G.data_ptr(), scratch.data_ptr(), Lh.data_ptr(), Linv.data_ptr(), None, 1e-14,
                             _stream(G)), 'mpdo_chol_psd')
    q = (A @ Linv.mH).reshape(*left, n).to(tensor.dtype)
    return q, Lh.reshape(n, *right).to(tensor.dtype)
