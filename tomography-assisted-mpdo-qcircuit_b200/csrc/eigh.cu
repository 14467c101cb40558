// Hermitian positive-semidefinite eigen-decomposition for the Gram matrices of the truncation path (sm_100a).
//
// Plain one-sided Jacobi on the rows of [G | I] needs 15-30+ sweeps on the graded spectra that MPDO bonds
// produce (measured: n = 512, eigenvalues 0.9^i: not converged after 30 sweeps). The preconditioned route of
// Drmac & Veselic (SIAM J. Matrix Anal. Appl. 29, 2008) is used instead:
//
//   1. rank-revealing pivoted Cholesky  G = L L^h  (columns of L in pivot order, stops at numerical rank r);
//   2. one-sided Jacobi on the r rows of Y = L^h (no accumulator: G = Y^h Y is invariant under row mixing);
//   3. eigenvalues = squared row norms, eigenvectors = normalised rows.
//
// The factor has the square root of G's condition number and is strongly column graded, so the Jacobi phase
// converges in a handful of sweeps on rows of length n instead of 2n, and skips the null space altogether.
//
// Cholesky kernel: the rows of L are dealt to C CTAs (R rows each, resident in shared memory); a step picks the
// largest remaining diagonal (each CTA publishes its best candidate together with that candidate's row of L, one
// device-wide barrier, everybody reads the winner), then every warp finishes one entry of the new column with a
// dot product against the pivot row (left-looking: no trailing-matrix traffic). Small matrices use C = 1 and no
// barrier; batches of them run one CTA per matrix.
#include <stdlib.h>

#include "common.cuh"

namespace mpdo {

struct CholArgs {
  int n, R;
  double rel;            // stop when the largest remaining diagonal <= rel * largest initial diagonal
  const double2* G;      // [batch][n][n]
  double2* Y;            // [batch][n][n]: row k = conj(column k of L); rows >= rank are zero
  double2* X;            // optional [batch][n][n], zeroed by the caller: X[k, piv_c] = (L11^-1)[k, c]
  double2* slots;        // [batch][2][C][n + 1] candidate exchange (C > 1 only)
  int* info;             // [batch][4]: rank, barrier counter, error flag, unused
};

constexpr int CHOL_THREADS = 256;

// shared-memory layout: Ls[R][n] | prow[n] | dl[R] (double) | chosen[R] (int) | piv[n] (int) | pad | Ws[R][n]
__host__ __device__ inline size_t chol_ws_offset(int n, int R) {
  size_t off = ((size_t)R * n + n) * sizeof(double2) + (size_t)R * 12 + (size_t)n * 4;
  return (off + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(CholArgs p) {
  extern __shared__ double2 csm[];
  __shared__ double sVal;
  __shared__ int sIdx, sCta;
  const int n = p.n, R = p.R, C = (int)gridDim.x;
  const int c = blockIdx.x, bidx = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  double2* Ls = csm;                          // [R][n]: own rows of L, column index = step
  double2* prow = Ls + (size_t)R * n;         // [n]: pivot row of the current step
  double* dl = reinterpret_cast<double*>(prow + n);   // [R]: remaining diagonal of own rows
  int* chosen = reinterpret_cast<int*>(dl + R);       // [R]
  int* pivRow = chosen + R;                             // [n]: pivot row of every step (inverse only)
  double2* Ws = reinterpret_cast<double2*>(reinterpret_cast<char*>(csm) + chol_ws_offset(n, R));   // [R][n]: own
                                                                                                  // columns of L11^-1
  double2* X = p.X ? p.X + (long long)bidx * n * n : nullptr;
  const int row0 = c * R;
  const int rows = max(0, min(R, n - row0));
  const double2* G = p.G + (long long)bidx * n * n;
  double2* Y = p.Y + (long long)bidx * n * n;
  double2* slots = p.slots + (long long)bidx * 2 * C * (n + 1);
  int* info = p.info + (long long)bidx * 4;
  unsigned* bar = reinterpret_cast<unsigned*>(info + 1);
  unsigned phase = 0;

  for (int r = tid; r < rows; r += blockDim.x) {
    dl[r] = G[(long long)(row0 + r) * n + row0 + r].x;
    chosen[r] = 0;
  }
  __syncthreads();

  double thresh = 0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    // ---- own best candidate (ties: lowest row) ----
    if (warp == 0) {
      double best = -1.0;
      int bi = 0x7fffffff;
      for (int r = lane; r < rows; r += 32)
        if (!chosen[r] && dl[r] > best) {
          best = dl[r];
          bi = r;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (lane == 0) {
        sVal = best;
        sIdx = bi;
      }
    }
    __syncthreads();
    double pval = sVal;
    int pg = sIdx == 0x7fffffff ? -1 : row0 + sIdx;   // global row of the pivot
    if (C > 1) {
      const int li = sIdx;
      double2* mine = slots + ((long long)(k & 1) * C + c) * (n + 1);
      if (tid == 0) mine[0] = make_double2(pval, (double)pg);
      if (pg >= 0)
        for (int j = tid; j < k; j += blockDim.x) mine[1 + j] = Ls[(size_t)li * n + j];
      matrix_barrier(bar, (unsigned)C, phase, info + 2);
      if (warp == 0) {
        double best = -1.0;
        int bc = 0x7fffffff, bg = -1;
        for (int cc = lane; cc < C; cc += 32) {
          const double2 h = __ldcg(slots + ((long long)(k & 1) * C + cc) * (n + 1));
          if (h.y >= 0 && h.x > best) {
            best = h.x;
            bc = cc;
            bg = (int)h.y;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
          const int og = __shfl_xor_sync(0xffffffffu, bg, o);
          if (ob > best || (ob == best && oc < bc)) {
            best = ob;
            bc = oc;
            bg = og;
          }
        }
        if (lane == 0) {
          sVal = best;
          sCta = bc;
          sIdx = bg;
        }
      }
      __syncthreads();
      pval = sVal;
      pg = sIdx;
      if (pg >= 0) {
        const double2* wrow = slots + ((long long)(k & 1) * C + sCta) * (n + 1) + 1;
        for (int j = tid; j < k; j += blockDim.x) prow[j] = __ldcg(wrow + j);
      }
    } else if (pg >= 0) {
      for (int j = tid; j < k; j += blockDim.x) prow[j] = Ls[(size_t)sIdx * n + j];
    }
    if (k == 0) thresh = p.rel * pval;
    if (pg < 0 || !(pval > thresh) || !(pval > 0.0) || *((volatile int*)(info + 2))) {   // same data in every CTA
      rank = k;
      break;
    }
    __syncthreads();
    const double piv = sqrt(pval), inv = 1.0 / piv;
    // ---- column k of L: one warp per own row ----
    for (int r = warp; r < rows; r += nwarps) {
      const int i = row0 + r;
      double2 out = make_double2(0.0, 0.0);
      if (i == pg) {
        out.x = piv;
      } else if (!chosen[r]) {
        double ar = 0, ai = 0;
        const double2* lrow = Ls + (size_t)r * n;
        for (int j = lane; j < k; j += 32) {   // L[i,j] * conj(L[p,j])
          const double2 a = lrow[j], b = prow[j];
          ar = fma(a.x, b.x, fma(a.y, b.y, ar));
          ai = fma(a.y, b.x, fma(-a.x, b.y, ai));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ar += __shfl_xor_sync(0xffffffffu, ar, o);
          ai += __shfl_xor_sync(0xffffffffu, ai, o);
        }
        const double2 g = G[(long long)pg * n + i];   // G[i,p] = conj(G[p,i])
        out.x = (g.x - ar) * inv;
        out.y = (-g.y - ai) * inv;
      }
      if (lane == 0) {
        if (i == pg)
          chosen[r] = 1;
        else if (!chosen[r])
          dl[r] = fmax(dl[r] - (out.x * out.x + out.y * out.y), 0.0);
        Ls[(size_t)r * n + k] = out;
        Y[(long long)k * n + i] = make_double2(out.x, -out.y);
      }
    }
    if (X) {
      // Row k of W = L11^-1 (both indices in pivot order), column c owned by CTA c % C:
      //   W[k,k] = 1/piv,  W[k,c] = -(sum_{c<=j<k} L[p_k,j] W[j,c]) / piv
      if (tid == 0) pivRow[k] = pg;
      for (int slot = warp; slot * C + c <= k; slot += nwarps) {
        const int col = slot * C + c;
        double2* wcol = Ws + (size_t)slot * n;
        double2 w = make_double2(inv, 0.0);
        if (col < k) {
          double ar = 0, ai = 0;
          for (int j = col + lane; j < k; j += 32) {
            const double2 a = prow[j], b = wcol[j];
            ar = fma(a.x, b.x, fma(-a.y, b.y, ar));
            ai = fma(a.x, b.y, fma(a.y, b.x, ai));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o);
            ai += __shfl_xor_sync(0xffffffffu, ai, o);
          }
          w.x = -ar * inv;
          w.y = -ai * inv;
        }
        if (lane == 0) {
          wcol[k] = w;
          X[(long long)k * n + (col == k ? pg : pivRow[col])] = w;
        }
      }
    }
    __syncthreads();
  }
  // rows of Y beyond the rank are zero
  const long long tail = (long long)(n - rank) * rows;
  for (long long idx = tid; idx < tail; idx += blockDim.x) {
    const int k = rank + (int)(idx / rows), r = (int)(idx % rows);
    Y[(long long)k * n + row0 + r] = make_double2(0.0, 0.0);
  }
  if (c == 0 && tid == 0) info[0] = rank;
}

static size_t chol_smem(int n, int R, bool inverse) {
  return chol_ws_offset(n, R) + (inverse ? (size_t)R * n * sizeof(double2) : 0);
}

// rows per CTA / CTAs per matrix and the number of matrices one launch can hold co-resident (a batch beyond that is
// factorised in consecutive launches of `chunk` matrices), or false when not even one matrix can be scheduled
static bool chol_plan(int batch, int n, bool inverse, int* Rout, int* Cout, int* chunkOut) {
  static int smemMax = 0, sms = 0, coop = 0;
  if (!smemMax) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax - 1024) != cudaSuccess) {
      cudaGetLastError();
      smemMax = 0;
      return false;
    }
  }
  const size_t cap = (size_t)smemMax - 1024;
  *chunkOut = batch;
  if (chol_smem(n, n, inverse) <= cap) {   // whole factor in one CTA
    *Rout = n;
    *Cout = 1;
    return true;
  }
  if (!coop) return false;
  int bestR = 0, bestC = 0, bestChunk = 0;
  for (int R = 8; R < n; R *= 2) {
    const size_t sm = chol_smem(n, R, inverse);
    if (sm > cap) break;
    const int C = (n + R - 1) / R;
    int perSm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, chol_kernel, CHOL_THREADS, sm) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    const long long capacity = (long long)perSm * sms;
    if ((long long)C * batch <= capacity) {   // the whole batch at the widest spread that fits
      *Rout = R;
      *Cout = C;
      return true;
    }
    if (capacity >= C) {   // remember the plan with the most matrices per launch (fewest CTAs per matrix)
      bestR = R;
      bestC = C;
      bestChunk = (int)(capacity / C);
    }
  }
  if (!bestR) return false;
  *Rout = bestR;
  *Cout = bestC;
  *chunkOut = bestChunk;
  return true;
}

struct EighLayout {
  size_t y, slots, info, work, total;
};

static EighLayout eigh_layout(int batch, int n, bool pre, int C) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  EighLayout l;
  const size_t ycols = pre ? (size_t)n : 2 * (size_t)n;
  l.y = 0;
  size_t off = up((size_t)batch * n * ycols * sizeof(double2));
  l.slots = off;
  off += up(pre && C > 1 ? (size_t)batch * 2 * C * (n + 1) * sizeof(double2) : 0);
  l.info = off;
  off += up((size_t)batch * 4 * sizeof(int));
  l.work = off;
  off += up((size_t)batch * 48 * sizeof(int));
  l.total = off;
  return l;
}

}  // namespace mpdo

extern "C" int64_t mpdo_eigh_psd_scratch_bytes(int batch, int n) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 256;
  // large enough for either route (the classic one needs [n][2n] rows; the exchange slots of the preconditioned
  // one never exceed 2 * ceil(n/8) rows of n + 1)
  EighLayout a = eigh_layout(batch, n, false, 1);
  EighLayout b = eigh_layout(batch, n, true, (n + 7) / 8);
  return (int64_t)(a.total > b.total ? a.total : b.total);
}

namespace mpdo {
// Launch the factorisation. slots / info: scratch ([batch][2][C][n+1] complex128, [batch][4] int32).
// Returns 1 when the shape cannot be scheduled (caller falls back), 0 on success, other values on CUDA errors.
static int chol_launch(int batch, int n, const void* G, void* Y, void* X, void* slotsBase, size_t slotsBytes,
                       int* info, double rel, cudaStream_t st, int* rcOut) {
  int R = 0, C = 0, chunk = 0;
  *rcOut = 0;
  if (!chol_plan(batch, n, X != nullptr, &R, &C, &chunk)) return 1;
  if (C > 1 && (size_t)batch * 2 * C * (n + 1) * sizeof(double2) > slotsBytes) return 1;
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * 4 * (size_t)batch, st);
  if (e == cudaSuccess && X) e = cudaMemsetAsync(X, 0, sizeof(double2) * (size_t)batch * n * n, st);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "chol memset: %s", cudaGetErrorString(e));
    *rcOut = (int)e;
    return 0;
  }
  const size_t smem = chol_smem(n, R, X != nullptr);
  for (int b0 = 0; b0 < batch; b0 += chunk) {   // one launch unless the batch exceeds what can be co-resident
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    const long long mat = (long long)b0 * n * n;
    CholArgs a;
    a.n = n;
    a.R = R;
    a.rel = rel;
    a.G = (const double2*)G + mat;
    a.Y = (double2*)Y + mat;
    a.X = X ? (double2*)X + mat : nullptr;
    a.slots = (double2*)slotsBase + (long long)b0 * 2 * C * (n + 1);
    a.info = info + 4LL * b0;
    TimedLaunch timed(2, 0.0, 0.0, st);
    if (C == 1) {
      chol_kernel<<<dim3(1, nb), CHOL_THREADS, smem, st>>>(a);
    } else {
      void* args[] = {(void*)&a};
      e = cudaLaunchCooperativeKernel((const void*)chol_kernel, dim3(C, nb), dim3(CHOL_THREADS), args, smem, st);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "chol_kernel (cooperative): %s", cudaGetErrorString(e));
        *rcOut = (int)e;
        return 0;
      }
    }
    *rcOut = check_launch("chol_kernel");
    if (*rcOut) return 0;
  }
  return 0;
}

__global__ void gather_rank_kernel(int batch, const int* __restrict__ info, int32_t* __restrict__ rank) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) rank[b] = info[4 * b];
}
}  // namespace mpdo

extern "C" int mpdo_eigh_psd(int batch, int n, const void* G, void* scratch, double* lam, void* Vh, int precondition,
                             double rel, double tol, int maxSweeps, void* stream) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 0;
  if (!G || !scratch || !lam || !Vh) return fail(MPDO_EINVAL, "mpdo_eigh_psd: null argument");
  if (batch > 65535) return fail(MPDO_EINVAL, "mpdo_eigh_psd: batch > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)scratch;
  static const bool off = getenv("MPDO_EIGH_CLASSIC") != nullptr;   // debugging knob
  if (precondition && !off && n > 1) {
    const EighLayout l = eigh_layout(batch, n, true, (n + 7) / 8);
    int rc = 0;
    if (chol_launch(batch, n, G, base + l.y, nullptr, base + l.slots, l.info - l.slots, (int*)(base + l.info), rel, st,
                    &rc) == 0) {
      if (rc) return rc;
      int32_t* work = (int32_t*)(base + l.work);
      rc = jacobi_rows_ranked(batch, n, n, n, n, (long long)n * n, base + l.y, tol, maxSweeps, work,
                              (const int*)(base + l.info), 4, stream);
      if (rc) return rc;
      return mpdo_rows_finalize(batch, n, n, 0, n, (long long)n * n, base + l.y, lam, Vh, nullptr, 3, 1e-300, stream);
    }
  }
  // classic route: Jacobi on [G | I]; rows of J.G are lam_j v_j^h and the accumulator holds Vh (complete basis)
  const EighLayout l = eigh_layout(batch, n, false, 1);
  return mpdo_decompose_rows(batch, n, n, G, base + l.y, (int32_t*)(base + l.work), lam, nullptr, Vh, 0, 0.0, tol,
                             maxSweeps, stream);
}

extern "C" int64_t mpdo_chol_psd_scratch_bytes(int batch, int n) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 256;
  const EighLayout l = eigh_layout(batch, n, true, (n + 7) / 8);
  return (int64_t)(l.total - l.slots);
}

extern "C" int mpdo_chol_psd(int batch, int n, const void* G, void* scratch, void* Lh, void* Linv, int32_t* rank,
                             double rel, void* stream) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 0;
  if (!G || !scratch || !Lh) return fail(MPDO_EINVAL, "mpdo_chol_psd: null argument");
  if (batch > 65535) return fail(MPDO_EINVAL, "mpdo_chol_psd: batch > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const EighLayout l = eigh_layout(batch, n, true, (n + 7) / 8);
  char* base = (char*)scratch - l.slots;   // the layout's Y slab is the caller's Lh
  int rc = 0;
  if (chol_launch(batch, n, G, Lh, Linv, base + l.slots, l.info - l.slots, (int*)(base + l.info), rel, st, &rc))
    return fail(MPDO_ENOSMEM, "mpdo_chol_psd: batch of this order cannot be made co-resident");
  if (rc) return rc;
  if (rank) {
    gather_rank_kernel<<<(batch + 127) / 128, 128, 0, st>>>(batch, (const int*)(base + l.info), rank);
    return check_launch("gather_rank_kernel");
  }
  return 0;
}
