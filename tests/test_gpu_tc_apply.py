"""-m gpu: the tcgen05 / TMA tile of the complex64 apply contractions (csrc/tc_apply.cu: tcgen05.mma kind::tf32 with a
3xTF32 split, TMA operand tiles, TMEM accumulators) against a complex128 torch model of the same product, and against
the FFMA tiles it replaces (mpdo_tc_enable(0)). Shapes: the QR-sweep / chi-sweep / gate-split applies of the chi = 128
and chi = 256 configurations, plus ragged edges (M not a multiple of 128, 2N not a multiple of the column tile, K not
a multiple of the k-block), conjugated operands, a scale factor, a batch, and a strided small operand."""
import pytest
import torch

pytestmark = pytest.mark.gpu
C64, C128 = torch.complex64, torch.complex128


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)).to('cuda')


def run(p, A, B, conjA=False, conjB=False, alpha=1.0):
    """C[b,i,j] = alpha * op(A)[b,i,k] op(B)[b,k,j] through mpdo_contract."""
    Bn, M, K = A.shape
    N = B.shape[2]
    out = torch.full((Bn, M, N), float('nan'), dtype=C64, device=A.device)
    p.contract(A, (1, 1, 1), B, (1, 1, 1), out, (1, 1, 1), conjA=conjA, conjB=conjB, alpha=alpha)
    return out


def model(A, B, conjA=False, conjB=False, alpha=1.0):
    a, b = A.to(C128), B.to(C128)
    return alpha * torch.matmul(a.conj() if conjA else a, b.conj() if conjB else b)


CASES = [
    # (batch, M, K, N, conjA, conjB, alpha)        the site-tensor side is M
    (1, 65536, 256, 256, False, True, 1.0),        # chi = 256: Q = T . Linv^h of the QR sweep
    (1, 32768, 128, 128, False, True, 1.0),        # chi = 128
    (1, 8192, 64, 64, False, True, 1.0),           # chi = 64
    (1, 16384, 512, 64, False, True, 1.0),         # chi sweep: T_l . U sqrt(S) from a 512-wide bond down to 64
    (1, 8192, 190, 64, False, False, 1.0),         # ragged k (380 real, not a multiple of 32)
    (1, 5000, 96, 100, True, False, -0.5),         # ragged rows and columns, conj(A), scale
    (3, 8192, 64, 36, False, False, 2.0),          # batch, narrow result
    (2, 4000, 32, 256, True, True, 1.0),           # two column tiles of 256 real columns
]


@pytest.mark.parametrize('case', CASES)
def test_tensor_core_apply_matches_fp64_model(cuda_prims, case):
    Bn, M, K, N, conjA, conjB, alpha = case
    p = cuda_prims
    A, B = rnd(Bn, M, K, seed=1), rnd(Bn, K, N, seed=2)
    want = model(A, B, conjA, conjB, alpha)
    prev = p.lib.mpdo_tc_enable(1)
    launches0 = p.launch_count()
    got_tc = run(p, A, B, conjA, conjB, alpha)
    launches_tc = p.launch_count() - launches0
    p.lib.mpdo_tc_enable(0)
    got_simt = run(p, A, B, conjA, conjB, alpha)
    launches_simt = p.launch_count() - launches0 - launches_tc
    p.lib.mpdo_tc_enable(prev)
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    err_tc = (got_tc.to(C128) - want).abs().max().item() / scale
    err_simt = (got_simt.to(C128) - want).abs().max().item() / scale
    fro_tc = (torch.linalg.norm(got_tc.to(C128) - want) / torch.linalg.norm(want)).item()
    fro_simt = (torch.linalg.norm(got_simt.to(C128) - want) / torch.linalg.norm(want)).item()
    print(f'{case}: tcgen05 3xTF32 rel Frobenius {fro_tc:.2e} (max abs / max |C| {err_tc:.2e}); '
          f'FFMA {fro_simt:.2e} ({err_simt:.2e})')
    assert not torch.isnan(got_tc.real).any()
    # 1e-6 relative (Frobenius) against the fp64 model; the worst single entry may sit a few times higher (fp32
    # accumulation of up to 2048 real products: the FFMA tiles reach 1.1e-6 on the same metric)
    assert fro_tc <= 1e-6 and err_tc <= 4e-6
    assert fro_simt <= 1e-6 and err_simt <= 4e-6
    assert (launches_tc, launches_simt) == (2, 1), (launches_tc, launches_simt)   # prep + tensor-core kernel | FFMA kernel


def test_wide_tile_mode_trades_accuracy_for_speed(cuda_prims):
    """mpdo_tc_enable(2): widest column tile whatever K - longer accumulator chains, still within 3e-6."""
    p = cuda_prims
    A, B = rnd(1, 65536, 256, seed=5), rnd(1, 256, 256, seed=6)
    want = model(A, B)
    prev = p.lib.mpdo_tc_enable(2)
    got = run(p, A, B)
    p.lib.mpdo_tc_enable(prev)
    fro = (torch.linalg.norm(got.to(C128) - want) / torch.linalg.norm(want)).item()
    print(f'wide tiles, K = 256: rel Frobenius {fro:.2e}')
    assert fro <= 3e-6


def test_tensor_core_apply_takes_views_of_the_small_operand(cuda_prims):
    """op(B) may be any strided view (here a transposed slice); the big operand must be a plain matrix."""
    p = cuda_prims
    A = rnd(1, 4096, 64, seed=3)
    Bfull = rnd(1, 80, 96, seed=4)
    Bv = Bfull[:, :48, 10:74].permute(0, 2, 1)            # [1, 64, 48], non-contiguous
    out = torch.empty((1, 4096, 48), dtype=C64, device='cuda')
    p.lib.mpdo_tc_enable(1)
    p.contract(A, (1, 1, 1), Bv, (1, 1, 1), out, (1, 1, 1))
    want = model(A, Bv)
    assert (torch.linalg.norm(out.to(C128) - want) / torch.linalg.norm(want)).item() <= 1e-6
