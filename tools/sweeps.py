"""Dev helper (GPU): sweeps and time of the Hermitian PSD eigen-decomposition, classic vs preconditioned route."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tomography-assisted-mpdo-qcircuit_b200'))
import torch
from MPDOSimulator._engine import lib as L
lib = L.load()
dev = 'cuda:0'
torch.manual_seed(0)
def run(n, decay, tol, pre, batch=1):
    A = torch.randn(n, n, dtype=torch.complex128, device=dev)
    Q, _ = torch.linalg.qr(A)
    lam = torch.tensor([max(decay ** i, 1e-30) for i in range(n)], dtype=torch.float64, device=dev)
    G = ((Q * lam.to(torch.complex128)) @ Q.mH).contiguous().unsqueeze(0).repeat(batch, 1, 1).contiguous()
    nbytes = lib.mpdo_eigh_psd_scratch_bytes(batch, n)
    scratch = torch.zeros((nbytes,), dtype=torch.uint8, device=dev)
    s = torch.empty((batch, n), dtype=torch.float64, device=dev)
    Z = torch.empty((batch, n, n), dtype=torch.complex128, device=dev)
    st = C.c_void_p(torch._C._cuda_getCurrentRawStream(0))
    p = lambda t: C.c_void_p(t.data_ptr())
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = lib.mpdo_eigh_psd(batch, n, p(G), p(scratch), p(s), p(Z), pre, 1e-15, tol, 30, st)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        assert rc == 0, lib.mpdo_last_error()
    lib.mpdo_timing_enable(1)
    lib.mpdo_eigh_psd(batch, n, p(G), p(scratch), p(s), p(Z), pre, 1e-15, tol, 30, st)
    torch.cuda.synchronize()
    tcls = []
    for cls in (1, 2):
        sec, fl, by, sol = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        cnt_, x = C.c_int64(), C.c_double()
        lib.mpdo_timing_summary(cls, 0.0, C.byref(sec), C.byref(fl), C.byref(by), C.byref(cnt_), C.byref(sol), C.byref(x))
        tcls.append(1e3 * sec.value)
    lib.mpdo_timing_enable(0)
    up = lambda v: (v + 255) & ~255
    ycols = n if pre else 2 * n
    Cc = (n + 7) // 8 if (pre and n > 112) else 1
    off = up(batch * n * ycols * 16) + up(batch * 2 * Cc * (n + 1) * 16 if (pre and Cc > 1) else 0) + up(batch * 16)
    work = scratch[off:off + 48 * 4].view(torch.int32)
    cnt = work[:32].cpu().tolist()
    sweeps = next((i + 1 for i, c in enumerate(cnt) if c == 0), 32)
    err = ((s[0] - lam).abs().max() / lam[0]).item()
    rec = (Z[0].mH @ (s[0].to(torch.complex128)[:, None] * Z[0]) - G[0]).abs().max().item()
    rank = int((s[0] > 0).sum())
    print('n=%4d B=%d decay=%.3f pre=%d: %.2f ms (jacobi %.2f, chol %.2f), sweeps=%d, rank=%d, eig err %.1e, rec err %.1e' % (
        n, batch, decay, pre, 1e3 * dt, tcls[0], tcls[1], sweeps, rank, err, rec), flush=True)
for n in (64, 128, 256, 512):
    for decay in (0.99, 0.9):
        for pre in (0, 1):
            run(n, decay, 1e-10, pre)
run(24, 0.9, 1e-10, 0, 64); run(24, 0.9, 1e-10, 1, 64)
run(100, 0.9, 1e-10, 0, 16); run(100, 0.9, 1e-10, 1, 16)
run(256, 0.9, 1e-10, 0, 4); run(256, 0.9, 1e-10, 1, 4)
