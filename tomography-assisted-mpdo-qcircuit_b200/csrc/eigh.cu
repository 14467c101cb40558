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
//
// Groups of up to 16 CTAs at 8 rows each (n <= 128) are launched as one thread-block cluster: candidates are
// published in the CTA's own shared memory, the step barrier is the hardware cluster barrier, and the winner's row
// of L is read straight out of the winning CTA's shared memory (DSMEM) - no global-memory slots, no L2 round trips,
// and no cooperative launch, so any batch size is one launch. Larger groups (n = 512 with the inverse: 64 CTAs) keep
// the device-wide barrier through L2 (see chol_cluster_plan for the measured crossover).
#include <cooperative_groups.h>
#include <stdlib.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mpdo {

struct CholArgs {
  int n, R;
  double rel;            // stop when the largest remaining diagonal <= rel * largest initial diagonal
  const double2* G;      // [batch][n][n]
  double2* Y;            // [batch][n][n]: row k = conj(column k of L); rows >= rank are zero
  double2* X;            // optional [batch][n][n], zeroed by the caller: X[k, piv_c] = (L11^-1)[k, c]
  double2* slots;        // [batch][2][C][n + 1] candidate exchange (C > 1 only)
  int* info;             // [batch][4]: rank, barrier counter, error flag, unused
  int cluster;           // 1: the C CTAs of a matrix form one thread-block cluster (exchange through DSMEM)
};

constexpr int CHOL_THREADS = 256;

// shared-memory layout: Ls[R][n] | prow[n] | grow[R] | dl[R] (double) | chosen[R] (int) | piv[n] (int) | pad | Ws[R][n]
__host__ __device__ inline size_t chol_ws_offset(int n, int R) {
  size_t off = ((size_t)R * n + n + R) * sizeof(double2) + (size_t)R * 12 + (size_t)n * 4;
  return (off + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(CHOL_THREADS) chol_kernel(CholArgs p) {
  extern __shared__ double2 csm[];
  __shared__ double sVal;
  __shared__ int sIdx, sCta;
  __shared__ double wbestV[32];   // per-warp best remaining diagonal / its local row
  __shared__ int wbestI[32];
  __shared__ double2 hdr[2];   // cluster mode: own candidate (value, global row) of the even / odd step
  const int n = p.n, R = p.R, C = (int)gridDim.x;
  const int c = blockIdx.x, bidx = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  double2* Ls = csm;                          // [R][n]: own rows of L, column index = step
  double2* prow = Ls + (size_t)R * n;         // [n]: pivot row of the current step
  double2* grow = prow + n;                   // [R]: G[pivot, own rows] of the current step (one coalesced load per
                                              // step; read per row from global memory it cost ~700 cycles per row)
  double* dl = reinterpret_cast<double*>(grow + R);   // [R]: remaining diagonal of own rows
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
  {   // per-warp best diagonal among the warp's rows (r = warp, warp + nwarps, ...)
    double wb = -1.0;
    int wi = 0x7fffffff;
    for (int r = warp; r < rows; r += nwarps)
      if (dl[r] > wb) {
        wb = dl[r];
        wi = r;
      }
    if (lane == 0) {
      wbestV[warp] = wb;
      wbestI[warp] = wi;
    }
  }
  __syncthreads();

  double thresh = 0;
  int rank = n;
#ifdef MPDO_CHOL_PROFILE   // nvcc -DMPDO_CHOL_PROFILE: CTA 0 of matrix 0 prints its cycles per phase of a step
  long long tp[6] = {0, 0, 0, 0, 0, 0};
  long long tc = clock64();
#define CHOL_MARK(i)                 \
  do {                               \
    const long long now_ = clock64(); \
    tp[i] += now_ - tc;              \
    tc = now_;                       \
  } while (0)
#else
#define CHOL_MARK(i)
#endif
  for (int k = 0; k < n; ++k) {
    // ---- own best candidate (ties: lowest row): the per-warp bests were left in shared memory by the column
    // phase of the previous step (or by the prologue), so this is nwarps broadcast reads, no shuffle tree ----
    double cbest = -1.0;
    int cidx = 0x7fffffff;
    for (int w = 0; w < nwarps; ++w) {
      const double v = wbestV[w];
      const int i = wbestI[w];
      if (v > cbest || (v == cbest && i < cidx)) {
        cbest = v;
        cidx = i;
      }
    }
    CHOL_MARK(0);
    double pval = cbest;
    int pg = cidx == 0x7fffffff ? -1 : row0 + cidx;   // global row of the pivot
    if (C > 1 && p.cluster) {
      cg::cluster_group cl = cg::this_cluster();
      if (tid == 0) hdr[k & 1] = make_double2(pval, (double)pg);
      cl.sync();   // every CTA's candidate of this step is visible cluster-wide (and step k-1 is complete everywhere)
      CHOL_MARK(1);
      if (warp == 0) {
        double best = -1.0;
        int bc = 0x7fffffff, bg = -1;
        for (int cc = lane; cc < C; cc += 32) {
          const double2 h = *cl.map_shared_rank(&hdr[k & 1], cc);
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
      CHOL_MARK(2);
      pval = sVal;
      pg = sIdx;
      if (pg >= 0) {
        // columns 0..k-1 of the winner's row are final (written in earlier steps), so they are read in place
        const double2* wrow = cl.map_shared_rank(Ls + (size_t)(pg - sCta * R) * n, sCta);
        for (int j = tid; j < k; j += blockDim.x) prow[j] = wrow[j];
      }
    } else if (C > 1) {
      const int li = cidx;
      double2* mine = slots + ((long long)(k & 1) * C + c) * (n + 1);
      if (tid == 0) mine[0] = make_double2(pval, (double)pg);
      if (pg >= 0)
        for (int j = tid; j < k; j += blockDim.x) mine[1 + j] = Ls[(size_t)li * n + j];
      matrix_barrier(bar, (unsigned)C, phase, info + 2);
      CHOL_MARK(1);
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
      CHOL_MARK(2);
      pval = sVal;
      pg = sIdx;
      if (pg >= 0) {
        const double2* wrow = slots + ((long long)(k & 1) * C + sCta) * (n + 1) + 1;
        for (int j = tid; j < k; j += blockDim.x) prow[j] = __ldcg(wrow + j);
      }
    } else if (pg >= 0) {
      for (int j = tid; j < k; j += blockDim.x) prow[j] = Ls[(size_t)cidx * n + j];
    }
    if (k == 0) thresh = p.rel * pval;
    if (pg < 0 || !(pval > thresh) || !(pval > 0.0) ||
        (C > 1 && !p.cluster && *((volatile int*)(info + 2)))) {   // same data in every CTA
      rank = k;
      break;
    }
    for (int r = tid; r < rows; r += blockDim.x) grow[r] = G[(long long)pg * n + row0 + r];
    __syncthreads();
    CHOL_MARK(3);
    const double piv = sqrt(pval), inv = 1.0 / piv;
    // ---- column k of L: one warp per own row; every lane carries the sums, so the warp also tracks the best
    // remaining diagonal among its rows for the next step's pivot search ----
    double wb = -1.0;
    int wi = 0x7fffffff;
    for (int r = warp; r < rows; r += nwarps) {
      const int i = row0 + r;
      const int wasChosen = chosen[r];
      const double dold = dl[r];
      __syncwarp();   // all lanes have read dl[r] / chosen[r] before lane 0 updates them
      double2 out = make_double2(0.0, 0.0);
      double dnew = dold;
      if (i == pg) {
        out.x = piv;
      } else if (!wasChosen) {
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
        const double2 g = grow[r];   // G[p,i]; G[i,p] is its conjugate
        out.x = (g.x - ar) * inv;
        out.y = (-g.y - ai) * inv;
        dnew = fmax(dold - (out.x * out.x + out.y * out.y), 0.0);
        if (dnew > wb) {   // rows ascend within a warp: ties keep the lowest row
          wb = dnew;
          wi = r;
        }
      }
      if (lane == 0) {
        if (i == pg)
          chosen[r] = 1;
        else if (!wasChosen)
          dl[r] = dnew;
        Ls[(size_t)r * n + k] = out;
        Y[(long long)k * n + i] = make_double2(out.x, -out.y);
      }
    }
    if (lane == 0) {
      wbestV[warp] = wb;
      wbestI[warp] = wi;
    }
    CHOL_MARK(4);
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
    CHOL_MARK(5);
  }
#ifdef MPDO_CHOL_PROFILE
  if (c == 0 && bidx == 0 && tid == 0)
    printf("[chol n=%d C=%d R=%d cluster=%d rank=%d] cycles/step: argmax %lld  barrier %lld  headers %lld  pivot-row %lld  column %lld  inverse+sync %lld\n",
           n, C, R, p.cluster, rank, tp[0] / max(rank, 1), tp[1] / max(rank, 1), tp[2] / max(rank, 1), tp[3] / max(rank, 1),
           tp[4] / max(rank, 1), tp[5] / max(rank, 1));
#endif
  if (C > 1 && p.cluster) cg::this_cluster().sync();   // nobody leaves while a neighbour may still read its rows
  // rows of Y beyond the rank are zero
  const long long tail = (long long)(n - rank) * rows;
  for (long long idx = tid; idx < tail; idx += blockDim.x) {
    const int k = rank + (int)(idx / rows), r = (int)(idx % rows);
    Y[(long long)k * n + row0 + r] = make_double2(0.0, 0.0);
  }
  if (c == 0 && tid == 0) info[0] = rank;
}

// ---- small matrices (n <= CS_MAXN): one CTA, thread per row ---------------------------------------------------
// The general kernel above gives one WARP to a row and reduces every dot product with 64-bit shuffles; with the 8 rows
// per warp of a 64 x 64 matrix that is ~9000 cycles per pivot step (0.28 ms per factorisation, and an MPDO layer at
// chi = 64 needs ~30 of them one after the other). Here a THREAD owns a row: L is kept transposed in shared memory
// (LT[j][i] = L[i,j], odd leading dimension), so the threads of a warp read consecutive entries of column j while
// every thread walks its own dot product sequentially - no shuffle reductions, no warp-per-row serialisation. The
// rows of the left inverse W = L11^-1 are computed at the same time by a second group of threads (thread per column).
// Same pivot order (largest remaining diagonal, ties to the lowest row), stop rule and outputs as chol_kernel.
constexpr int CS_MAXN = 80;       // LT and W in shared memory: 2 n (n + 1) 16 B = 207 KB at n = 80
constexpr int CS_ROWT = 96;       // threads 0..95: rows;  96..191: columns of the inverse
constexpr int CS_THREADS = 192;

__global__ void __launch_bounds__(CS_THREADS) chol_small_kernel(CholArgs p) {
  extern __shared__ double2 csm[];
  __shared__ double sBestV[CS_ROWT / 32];
  __shared__ int sBestI[CS_ROWT / 32];
  __shared__ int sPiv[CS_MAXN];
  const int n = p.n, ld = n | 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bidx = blockIdx.y;
  const double2* G = p.G + (long long)bidx * n * n;
  double2* Y = p.Y + (long long)bidx * n * n;
  double2* X = p.X ? p.X + (long long)bidx * n * n : nullptr;
  double2* LT = csm;                                   // [n][ld]: LT[j * ld + i] = L[i, j]
  double2* Wm = LT + (size_t)n * ld;                   // [n][ld]: Wm[j * ld + c] = W[j, c] (pivot order), if X
  double2* prow = (X ? Wm : LT) + (size_t)n * ld;      // [n]: L[pivot, 0..k)
  double2* grow = prow + n;                            // [n]: G[pivot, :]
  const bool rowThread = tid < CS_ROWT;
  const int i = tid;                                   // row owned by a row thread
  const int c = tid - CS_ROWT;                         // inverse column owned by the other threads
  double d = (rowThread && i < n) ? G[(long long)i * n + i].x : -1.0;
  bool chosen = !(rowThread && i < n);
  double thresh = 0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    if (rowThread) {   // largest remaining diagonal, ties to the lowest row
      double v = chosen ? -1.0 : d;
      int vi = chosen ? 0x7fffffff : i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (lane == 0) {
        sBestV[warp] = v;
        sBestI[warp] = vi;
      }
    }
    __syncthreads();
    double pval = -1.0;
    int pg = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < CS_ROWT / 32; ++w) {
      const double v = sBestV[w];
      const int vi = sBestI[w];
      if (v > pval || (v == pval && vi < pg)) {
        pval = v;
        pg = vi;
      }
    }
    if (k == 0) thresh = p.rel * pval;
    if (pg == 0x7fffffff || !(pval > thresh) || !(pval > 0.0)) {   // the same data in every thread
      rank = k;
      break;
    }
    for (int j = tid; j < k; j += CS_THREADS) prow[j] = LT[(size_t)j * ld + pg];
    if (!rowThread)
      for (int r = c; r < n; r += CS_THREADS - CS_ROWT) grow[r] = G[(long long)pg * n + r];
    if (tid == 0) sPiv[k] = pg;
    __syncthreads();
    const double piv = sqrt(pval), inv = 1.0 / piv;
    if (rowThread) {
      if (i < n) {
        double2 out = make_double2(0.0, 0.0);
        if (i == pg) {
          out.x = piv;
          chosen = true;
        } else if (!chosen) {
          double ar0 = 0, ai0 = 0, ar1 = 0, ai1 = 0;   // sum_j L[i,j] conj(L[p,j]), two chains for ILP
          int j = 0;
          for (; j + 1 < k; j += 2) {
            const double2 a0 = LT[(size_t)j * ld + i], b0 = prow[j];
            const double2 a1 = LT[(size_t)(j + 1) * ld + i], b1 = prow[j + 1];
            ar0 = fma(a0.x, b0.x, fma(a0.y, b0.y, ar0));
            ai0 = fma(a0.y, b0.x, fma(-a0.x, b0.y, ai0));
            ar1 = fma(a1.x, b1.x, fma(a1.y, b1.y, ar1));
            ai1 = fma(a1.y, b1.x, fma(-a1.x, b1.y, ai1));
          }
          if (j < k) {
            const double2 a0 = LT[(size_t)j * ld + i], b0 = prow[j];
            ar0 = fma(a0.x, b0.x, fma(a0.y, b0.y, ar0));
            ai0 = fma(a0.y, b0.x, fma(-a0.x, b0.y, ai0));
          }
          const double2 g = grow[i];                   // G[p,i]; G[i,p] is its conjugate
          out.x = (g.x - (ar0 + ar1)) * inv;
          out.y = (-g.y - (ai0 + ai1)) * inv;
          d = fmax(d - (out.x * out.x + out.y * out.y), 0.0);
        }
        LT[(size_t)k * ld + i] = out;
        Y[(long long)k * n + i] = make_double2(out.x, -out.y);
      }
    } else if (X && c <= k) {
      // row k of W = L11^-1 (pivot order):  W[k,k] = 1/piv,  W[k,c] = -(sum_{c<=j<k} L[p_k,j] W[j,c]) / piv
      double2 w = make_double2(inv, 0.0);
      if (c < k) {
        double ar0 = 0, ai0 = 0, ar1 = 0, ai1 = 0;
        int j = c;
        for (; j + 1 < k; j += 2) {
          const double2 a0 = prow[j], b0 = Wm[(size_t)j * ld + c];
          const double2 a1 = prow[j + 1], b1 = Wm[(size_t)(j + 1) * ld + c];
          ar0 = fma(a0.x, b0.x, fma(-a0.y, b0.y, ar0));
          ai0 = fma(a0.x, b0.y, fma(a0.y, b0.x, ai0));
          ar1 = fma(a1.x, b1.x, fma(-a1.y, b1.y, ar1));
          ai1 = fma(a1.x, b1.y, fma(a1.y, b1.x, ai1));
        }
        if (j < k) {
          const double2 a0 = prow[j], b0 = Wm[(size_t)j * ld + c];
          ar0 = fma(a0.x, b0.x, fma(-a0.y, b0.y, ar0));
          ai0 = fma(a0.x, b0.y, fma(a0.y, b0.x, ai0));
        }
        w.x = -(ar0 + ar1) * inv;
        w.y = -(ai0 + ai1) * inv;
      }
      Wm[(size_t)k * ld + c] = w;
      X[(long long)k * n + (c == k ? pg : sPiv[c])] = w;
    }
    // no barrier here: the next step's pivot search only touches registers, and its first barrier orders this
    // step's shared-memory writes before anybody reads them
  }
  for (long long idx = (long long)rank * n + tid; idx < (long long)n * n; idx += CS_THREADS)
    Y[idx] = make_double2(0.0, 0.0);                    // rows of Y beyond the rank are zero
  if (tid == 0) p.info[4LL * bidx] = rank;
}

static size_t chol_small_smem(int n, bool inverse) {
  const size_t ld = (size_t)(n | 1);
  return ((inverse ? 2 : 1) * (size_t)n * ld + 2 * (size_t)n) * sizeof(double2);
}

// ---- medium matrices (CS_MAXN < n <= 256): one 16-CTA thread-block cluster, 16 threads per row ----------------------
// Same idea as chol_small_kernel with the rows of L dealt to the 16 CTAs of a cluster (<= 16 rows each, transposed
// slices in shared memory): a dot product is split over 16 lanes (four 64-bit shuffle stages instead of five, two rows
// per warp in flight), the candidates of a step are exchanged through distributed shared memory behind ONE cluster
// barrier per pivot step, and the winner's row is gathered straight out of the owning CTA's slice. Column c of the
// left inverse lives in CTA c % 16 and is extended by a second group of 256 threads while the first computes the new
// column of L. Outputs, pivot order and stop rule as chol_kernel.
constexpr int CC_CTAS = 16;        // cluster size (non-portable: > 8)
constexpr int CC_R = 16;           // rows per CTA, so n <= 256
constexpr int CC_LD = CC_R + 1;    // odd leading dimension of the transposed slice
constexpr int CC_THREADS = 512;    // 0..255: 16 rows x 16 lanes;  256..511: 16 inverse columns x 16 lanes

__global__ void __launch_bounds__(CC_THREADS) chol_cluster_kernel(CholArgs p) {
  extern __shared__ double2 csm[];
  __shared__ double2 hdr[2];        // own candidate (value, global row) of the even / odd step
  __shared__ double sVal;
  __shared__ int sIdx;
  __shared__ double dl[CC_R];
  __shared__ int chosen[CC_R];
  __shared__ int sPiv[CC_CTAS * CC_R];
  cg::cluster_group cl = cg::this_cluster();
  const int n = p.n, bidx = blockIdx.y;
  const int c = (int)cl.block_rank();
  const int tid = threadIdx.x, lane = tid & 31;
  const int R = (n + CC_CTAS - 1) / CC_CTAS;
  const int row0 = c * R;
  const int rows = max(0, min(R, n - row0));
  const double2* G = p.G + (long long)bidx * n * n;
  double2* Y = p.Y + (long long)bidx * n * n;
  double2* X = p.X ? p.X + (long long)bidx * n * n : nullptr;
  double2* LT = csm;                               // [n][CC_LD]: LT[j * CC_LD + r] = L[row0 + r, j]
  double2* prow = LT + (size_t)n * CC_LD;          // [n]: L[pivot, 0..k)
  double2* grow = prow + n;                        // [CC_R]: G[pivot, own rows]
  double2* Wc = grow + CC_R;                       // [CC_R][n]: Wc[s * n + j] = W[j, s * 16 + c] (if X)
  for (int r = tid; r < CC_R; r += CC_THREADS) {
    dl[r] = r < rows ? G[(long long)(row0 + r) * n + row0 + r].x : -1.0;
    chosen[r] = r < rows ? 0 : 1;
  }
  __syncthreads();
  double thresh = 0;
  int rank = n;
  for (int k = 0; k < n; ++k) {
    if (tid < 32) {   // own best candidate (ties: lowest row)
      double v = (lane < CC_R && !chosen[lane]) ? dl[lane] : -1.0;
      int vi = (lane < CC_R && !chosen[lane]) ? row0 + lane : 0x7fffffff;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (lane == 0) hdr[k & 1] = make_double2(v, vi == 0x7fffffff ? -1.0 : (double)vi);
    }
    cl.sync();        // every CTA's candidate of this step (and its column k-1 of L) is visible cluster-wide
    if (tid < 32) {
      double v = -1.0;
      int vi = 0x7fffffff;
      if (lane < CC_CTAS) {
        const double2 h = *cl.map_shared_rank(&hdr[k & 1], lane);
        if (h.y >= 0) {
          v = h.x;
          vi = (int)h.y;
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) {
          v = ov;
          vi = oi;
        }
      }
      if (lane == 0) {
        sVal = v;
        sIdx = vi == 0x7fffffff ? -1 : vi;
      }
    }
    __syncthreads();
    const double pval = sVal;
    const int pg = sIdx;
    if (k == 0) thresh = p.rel * pval;
    if (pg < 0 || !(pval > thresh) || !(pval > 0.0)) {   // the same data in every CTA of the cluster
      rank = k;
      break;
    }
    {
      const int owner = pg / R, pr = pg - owner * R;
      const double2* remote = cl.map_shared_rank(LT, owner);        // columns < k of the winner's row are final
      for (int j = tid; j < k; j += CC_THREADS) prow[j] = remote[(size_t)j * CC_LD + pr];
      if (tid >= CC_THREADS - 32 && lane < rows) grow[lane] = G[(long long)pg * n + row0 + lane];
      if (tid == 0) sPiv[k] = pg;
    }
    __syncthreads();
    const double piv = sqrt(pval), inv = 1.0 / piv;
    const int t = tid & 15;
    const unsigned half = (lane < 16) ? 0x0000ffffu : 0xffff0000u;   // the 16 lanes that share a row / a column
    if (tid < 256) {
      const int r = tid >> 4;
      if (r < rows) {                          // uniform over the 16 lanes of a row
        const int i = row0 + r;
        const int wasChosen = chosen[r];
        const double dold = dl[r];
        double2 out = make_double2(0.0, 0.0);
        double dnew = dold;
        double ar = 0, ai = 0;
        if (i != pg && !wasChosen)
          for (int j = t; j < k; j += 16) {   // sum_j L[i,j] conj(L[p,j])
            const double2 a = LT[(size_t)j * CC_LD + r], b = prow[j];
            ar = fma(a.x, b.x, fma(a.y, b.y, ar));
            ai = fma(a.y, b.x, fma(-a.x, b.y, ai));
          }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          ar += __shfl_xor_sync(half, ar, o);
          ai += __shfl_xor_sync(half, ai, o);
        }
        if (i == pg) {
          out.x = piv;
        } else if (!wasChosen) {
          const double2 g = grow[r];           // G[p,i]; G[i,p] is its conjugate
          out.x = (g.x - ar) * inv;
          out.y = (-g.y - ai) * inv;
          dnew = fmax(dold - (out.x * out.x + out.y * out.y), 0.0);
        }
        __syncwarp(half);                      // dl / chosen were read by all 16 lanes before the leader updates them
        if (t == 0) {
          if (i == pg)
            chosen[r] = 1;
          else if (!wasChosen)
            dl[r] = dnew;
          LT[(size_t)k * CC_LD + r] = out;
          Y[(long long)k * n + i] = make_double2(out.x, -out.y);
        }
      }
    } else {
      // row k of W = L11^-1 (pivot order); column col = s * 16 + c is extended by slot s of this CTA:
      //   W[k,k] = 1/piv,  W[k,col] = -(sum_{col<=j<k} L[p_k,j] W[j,col]) / piv
      const int s_ = (tid - 256) >> 4;
      const int col = s_ * CC_CTAS + c;
      const bool active = X != nullptr && col <= k && col < n;   // uniform over the 16 lanes of a slot
      if (active) {
        double ar = 0, ai = 0;
        if (col < k) {
          const double2* wcol = Wc + (size_t)s_ * n;
          for (int j = col + t; j < k; j += 16) {
            const double2 a = prow[j], b = wcol[j];
            ar = fma(a.x, b.x, fma(-a.y, b.y, ar));
            ai = fma(a.x, b.y, fma(a.y, b.x, ai));
          }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          ar += __shfl_xor_sync(half, ar, o);
          ai += __shfl_xor_sync(half, ai, o);
        }
        if (t == 0) {
          const double2 w = col == k ? make_double2(inv, 0.0) : make_double2(-ar * inv, -ai * inv);
          Wc[(size_t)s_ * n + k] = w;
          X[(long long)k * n + (col == k ? pg : sPiv[col])] = w;
        }
      }
    }
    __syncthreads();
  }
  cl.sync();   // nobody leaves while a neighbour may still read its slice
  for (long long idx = tid; idx < (long long)(n - rank) * rows; idx += CC_THREADS) {
    const int k = rank + (int)(idx / rows), r = (int)(idx % rows);
    Y[(long long)k * n + row0 + r] = make_double2(0.0, 0.0);       // rows of Y beyond the rank are zero
  }
  if (c == 0 && tid == 0) p.info[4LL * bidx] = rank;
}

// ---- medium matrices, no pivoting: blocked right-looking Cholesky on a 16-CTA cluster --------------------------------
// The pivoted kernels above pay one cluster barrier (and two dependent DSMEM round trips) for every pivot: 3.3 us per
// step, 0.64 / 0.96 ms at n = 192 / 256 - a third of the device time of an MPDO layer at chi = 64, all of it on the
// sequential chain of the truncation sweep. Where the factor only preconditions the Jacobi eigen-solver the pivot order
// buys nothing on this workload: on the Gram matrices of the sweeps (tools/data/grams_sel.npz, n = 148 / 194 / 256)
// one-sided Jacobi takes the same number of sweeps (8 / 7 / 9), the same number of rotations to within 1 % and reaches
// the same relative eigenvalue accuracy (1e-13) on the unpivoted factor as on the pivoted one (model: tools/jacobi_model
// .py). Without pivoting the factorisation blocks: the rows are dealt to the ceil(n / 16) CTAs of a cluster in panels
// of R = 16 (CTA c keeps rows 16 c .. 16 c + 15 of the lower triangle in shared memory), and a panel costs two cluster
// barriers whatever its width:
//   1. the owner factors its R x R diagonal block and inverts the factor;                           cluster barrier
//   2. every CTA below multiplies its block of the panel by the inverse (L_qp = A_qp L_pp^-h);      cluster barrier
//   3. every CTA below pulls the panel blocks L_rp of the CTAs between the owner and itself through distributed shared
//      memory and updates its slab, A_qr -= L_qp L_rp^h.
// Rank safety without pivoting: a diagonal that has dropped to rel * max diag(G) is a null direction; for a positive
// semidefinite matrix its whole column of the Schur complement is then zero to the same level (|S_ik|^2 <= S_ii S_kk),
// so the column is skipped (zero column of L, the rows of Y stay compact) instead of being divided by.
// Outputs as chol_kernel (Y row k = conj(column k of L) over the accepted columns, info[0] = their number) except that
// the columns follow the initial diagonal order; no left inverse (p.X must be null; Y zeroed by the caller).
constexpr int CU_CTAS = 16;
constexpr int CU_RMAX = 16;
constexpr int CU_THREADS = 256;
constexpr int CU_LB = CU_RMAX + 1;   // leading dimension of the staged R x R blocks

__global__ void __launch_bounds__(CU_THREADS) chol_blocked_kernel(CholArgs p) {
  extern __shared__ double2 usm[];
  __shared__ double sDiag[CU_CTAS * CU_RMAX];   // diagonal of G
  __shared__ int sPerm[CU_CTAS * CU_RMAX];      // position -> original index, by decreasing diagonal
  __shared__ int sAcc[CU_RMAX + 1];  // owner of the current panel: compact index of each column or -1; [R] = count
  __shared__ int sAccL[CU_RMAX + 1]; // local copy of the current owner's header
  cg::cluster_group cl = cg::this_cluster();
  const int n = p.n, R = p.R, bidx = blockIdx.y;
  const int c = (int)cl.block_rank();
  const int C = (n + R - 1) / R;     // panels = CTAs that own rows
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;
  const int row0 = c * R;
  const int rows = max(0, min(R, n - row0));
  const int ld = n + 1;
  const double2* G = p.G + (long long)bidx * n * n;
  double2* Y = p.Y + (long long)bidx * n * n;
  double2* A = usm;                                   // [R][ld]: A[r * ld + j] = entry (row0 + r, j), j <= row0 + r
  double2* Lall = A + (size_t)R * ld;                 // [C - 1][R][CU_LB]: panel blocks of the CTAs above
  double2* Dinv = Lall + (size_t)(C > 1 ? C - 1 : 1) * R * CU_LB;   // [R][CU_LB]: inverse of the diagonal factor

  // Symmetric permutation by decreasing diagonal first ("diagonal pivoting at the start", every CTA computes the same
  // order): a direction whose diagonal is tiny from the outset is then met last, after everything it is coupled to has
  // been eliminated. Skipping a tiny pivot EARLY would drop couplings of up to sqrt(rel) relative size
  // (|S_ik|^2 <= S_ii S_kk is all positive semidefiniteness gives; measured 3e-8 on a complex128 circuit).
  for (int i = tid; i < n; i += CU_THREADS) sDiag[i] = G[(long long)i * n + i].x;
  __syncthreads();
  for (int i = tid; i < n; i += CU_THREADS) {
    const double di = sDiag[i];
    int before = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = sDiag[j];
      before += (dj > di || (dj == di && j < i)) ? 1 : 0;
    }
    sPerm[before] = i;
  }
  __syncthreads();
  for (int idx = tid; idx < rows * n; idx += CU_THREADS) {
    const int r = idx / n, j = idx - r * n;
    if (j <= row0 + r) A[(size_t)r * ld + j] = G[(long long)sPerm[row0 + r] * n + sPerm[j]];
  }
  const double thresh = p.rel * sDiag[sPerm[0]];

  int base = 0;   // accepted columns before the current panel (the same number in every CTA)
  for (int pp = 0; pp < C; ++pp) {
    const int c0 = pp * R;
    const int pr = min(R, n - c0);
    if (c == pp) {
      // ---- 1. diagonal block: unpivoted Cholesky with null-pivot skipping, thread (ti, tj) owns entry (ti, tj)
      int cnt = 0;
      for (int k = 0; k < pr; ++k) {
        __syncthreads();
        const double d = A[(size_t)k * ld + c0 + k].x;
        const bool ok = d > thresh && d > 0.0;
        const double s = ok ? sqrt(d) : 0.0;
        if (tj == k && ti > k && ti < pr) {          // column k below the diagonal (the diagonal entry itself is
          double2 v = A[(size_t)ti * ld + c0 + k];   // still being read as `d` by slower warps: written below)
          if (!ok) {
            v = make_double2(0.0, 0.0);
          } else {
            const double inv = 1.0 / s;
            v.x *= inv;
            v.y *= inv;
          }
          A[(size_t)ti * ld + c0 + k] = v;
        }
        if (tid == 0) sAcc[k] = ok ? cnt : -1;
        cnt += ok ? 1 : 0;
        __syncthreads();
        if (ti == k && tj == k) A[(size_t)k * ld + c0 + k] = make_double2(s, 0.0);
        if (ok && tj > k && tj <= ti && ti < pr) {   // trailing part of the block: A[i][j] -= L[i][k] conj(L[j][k])
          const double2 a = A[(size_t)ti * ld + c0 + k], b = A[(size_t)tj * ld + c0 + k];
          double2 v = A[(size_t)ti * ld + c0 + tj];
          v.x -= a.x * b.x + a.y * b.y;
          v.y -= a.y * b.x - a.x * b.y;
          A[(size_t)ti * ld + c0 + tj] = v;
        }
      }
      if (tid == 0) sAcc[R] = cnt;
      __syncthreads();
      // inverse of the factor, X = L_pp^-1 (lower triangular; zero rows / columns for skipped pivots): thread per
      // column, forward substitution down the column
      if (tid < pr) {
        const int j = tid;
        const double ljj = A[(size_t)j * ld + c0 + j].x;
        for (int i = 0; i < j; ++i) Dinv[i * CU_LB + j] = make_double2(0.0, 0.0);
        Dinv[j * CU_LB + j] = make_double2(ljj > 0.0 ? 1.0 / ljj : 0.0, 0.0);
        for (int i = j + 1; i < pr; ++i) {
          const double lii = A[(size_t)i * ld + c0 + i].x;
          double ar = 0.0, ai = 0.0;
          for (int m = j; m < i; ++m) {
            const double2 a = A[(size_t)i * ld + c0 + m], x = Dinv[m * CU_LB + j];
            ar += a.x * x.x - a.y * x.y;
            ai += a.x * x.y + a.y * x.x;
          }
          const double inv = (lii > 0.0 && ljj > 0.0) ? -1.0 / lii : 0.0;
          Dinv[i * CU_LB + j] = make_double2(ar * inv, ai * inv);
        }
      }
      // rows of Y for the accepted columns of the diagonal block
      if (ti < pr && tj <= ti) {
        const int a = sAcc[tj];
        if (a >= 0) {
          const double2 v = A[(size_t)ti * ld + c0 + tj];
          Y[(long long)(base + a) * n + sPerm[row0 + ti]] = make_double2(v.x, -v.y);
        }
      }
    }
    cl.sync();   // the owner's header and inverse are visible cluster-wide
    if (tid <= R) sAccL[tid] = cl.map_shared_rank(sAcc, pp)[tid];
    if (c > pp && rows > 0) {
      // ---- 2. own block of the panel: L_qp[i][k] = sum_{m <= k} A_qp[i][m] conj(X[k][m])
      double2* Dl = Lall;   // staged copy of the owner's inverse (Lall is free until step 3)
      if (ti < pr && tj < pr) Dl[ti * CU_LB + tj] = cl.map_shared_rank(Dinv, pp)[ti * CU_LB + tj];
      __syncthreads();
      double2 val = make_double2(0.0, 0.0);
      if (ti < rows && tj < pr) {
        for (int m = 0; m <= tj; ++m) {
          const double2 a = A[(size_t)ti * ld + c0 + m], x = Dl[tj * CU_LB + m];
          val.x += a.x * x.x + a.y * x.y;
          val.y += a.y * x.x - a.x * x.y;
        }
      }
      __syncthreads();
      if (ti < rows && tj < pr) {
        A[(size_t)ti * ld + c0 + tj] = val;
        const int a = sAccL[tj];
        if (a >= 0) Y[(long long)(base + a) * n + sPerm[row0 + ti]] = make_double2(val.x, -val.y);
      }
    } else {
      __syncthreads();   // sAccL visible to the whole CTA
    }
    cl.sync();   // every block of the panel is final
    if (c > pp && rows > 0) {
      // ---- 3. trailing update of the slab: pull L_rp of the CTAs pp < r < c, then A_qr -= L_qp L_rp^h for pp < r <= c
      const int nb = c - pp - 1;
      for (int idx = tid; idx < nb * R * pr; idx += CU_THREADS) {
        const int b = idx / (R * pr), rem = idx - b * (R * pr);
        const int j = rem / pr, m = rem - j * pr;
        Lall[((size_t)b * R + j) * CU_LB + m] = cl.map_shared_rank(A, pp + 1 + b)[(size_t)j * ld + c0 + m];
      }
      __syncthreads();
      if (ti < rows) {
        double2 lq[CU_RMAX];
#pragma unroll
        for (int m = 0; m < CU_RMAX; ++m) lq[m] = m < pr ? A[(size_t)ti * ld + c0 + m] : make_double2(0.0, 0.0);
        for (int r = pp + 1; r <= c; ++r) {
          const int rr = min(R, n - r * R);            // rows of CTA r = columns of this block
          if (tj >= rr || (r == c && tj > ti)) continue;
          const double2* lr = r == c ? A + (size_t)tj * ld + c0 : Lall + ((size_t)(r - pp - 1) * R + tj) * CU_LB;
          double sr = 0.0, si = 0.0;
#pragma unroll
          for (int m = 0; m < CU_RMAX; ++m) {
            if (m < pr) {
              const double2 b = lr[m];
              sr += lq[m].x * b.x + lq[m].y * b.y;
              si += lq[m].y * b.x - lq[m].x * b.y;
            }
          }
          double2 v = A[(size_t)ti * ld + r * R + tj];
          v.x -= sr;
          v.y -= si;
          A[(size_t)ti * ld + r * R + tj] = v;
        }
      }
    }
    base += sAccL[R];
    __syncthreads();   // sAccL is rewritten after the next cluster barrier; the slab update is complete
  }
  cl.sync();   // nobody leaves while a neighbour may still read its slab
  if (c == 0 && tid == 0) p.info[4LL * bidx] = base;
}

// panels of CU_RMAX rows, one CTA each: the cluster has ceil(n / 16) CTAs
static size_t chol_blocked_smem(int n) {
  const int R = CU_RMAX, C = (n + R - 1) / R;
  return ((size_t)R * (n + 1) + (size_t)(C > 1 ? C - 1 : 1) * R * CU_LB + (size_t)R * CU_LB) * sizeof(double2);
}

static size_t chol_cluster_smem(int n, bool inverse) {
  return ((size_t)n * CC_LD + (size_t)n + CC_R + (inverse ? (size_t)CC_R * n : 0)) * sizeof(double2);
}

static size_t chol_smem(int n, int R, bool inverse) {
  return chol_ws_offset(n, R) + (inverse ? (size_t)R * n * sizeof(double2) : 0);
}

// rows per CTA / CTAs per matrix and the number of matrices one launch can hold co-resident (a batch beyond that is
// factorised in consecutive launches of `chunk` matrices), or false when not even one matrix can be scheduled
static bool chol_plan(int batch, int n, bool inverse, int* Rout, int* Cout, int* chunkOut) {
  // one-time set-up behind a mutex: strands call this from a dozen host threads at once, and a thread must not see
  // the cached attribute before the kernel's shared-memory limit has actually been raised (observed as a sporadic
  // "chol_kernel: invalid argument" on the first layer of a process)
  static std::mutex initMu;
  static int smemMax = 0, sms = 0, coop = 0;
  {
    std::lock_guard<std::mutex> lk(initMu);
    if (!smemMax) {
      int dev = 0, sm = 0;
      if (cudaGetDevice(&dev) != cudaSuccess) return false;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      cudaDeviceGetAttribute(&sm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      if (cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm - 1024) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      smemMax = sm;
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

// Cluster plan: the widest spread of at most 16 CTAs per matrix whose rows fit shared memory, checked once per shape
// against the device (cudaOccupancyMaxActiveClusters). Cluster sizes above 8 are non-portable and opted into.
static bool chol_cluster_plan(int n, bool inverse, int* Rout, int* Cout) {
  static const bool off = getenv("MPDO_CHOL_NOCLUSTER") != nullptr;   // A/B knob: device-wide barrier through L2
  if (off) return false;
  static std::mutex initMu;   // see chol_plan: the cached value is published only after the attributes are set
  static int smemMax = 0;
  {
    std::lock_guard<std::mutex> lk(initMu);
    if (!smemMax) {
      int dev = 0, sm = 0;
      if (cudaGetDevice(&dev) != cudaSuccess) return false;
      cudaDeviceGetAttribute(&sm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      if (cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm - 1024) != cudaSuccess ||
          cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        smemMax = -1;
      } else {
        smemMax = sm;
      }
    }
  }
  if (smemMax <= 0) return false;
  const size_t cap = (size_t)smemMax - 1024;
  {
    // Only the widest spread (R = 8 rows per CTA, one row per warp) is worth a cluster: measured on B200, n = 128
    // (16 CTAs) 0.450 ms against 0.522 ms through L2, but n = 256 as 16 CTAs x 16 rows 1.242 ms against 1.079 ms as
    // 32 CTAs x 8 rows through L2 - the per-step column and inverse phases grow with the rows a CTA owns and outweigh
    // the cheaper barrier. So: clusters for n <= 128, the L2 barrier beyond.
    const int R = 8;
    const int C = (n + R - 1) / R;
    if (C > 16 || C < 2) return false;
    const size_t sm = chol_smem(n, R, inverse);
    if (sm > cap) return false;
    // one query per (C, shared memory) pair; the answer does not change while the process lives
    static std::mutex mu;
    static std::vector<std::pair<std::pair<int, size_t>, int>> seen;
    std::lock_guard<std::mutex> lk(mu);
    int ok = -1;
    for (auto& e : seen)
      if (e.first.first == C && e.first.second == sm) ok = e.second;
    if (ok < 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(C, 1);
      cfg.blockDim = dim3(CHOL_THREADS);
      cfg.dynamicSmemBytes = sm;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, chol_kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        nclusters = 0;
      }
      ok = nclusters >= 1 ? 1 : 0;
      seen.push_back({{C, sm}, ok});
    }
    if (!ok) return false;
    *Rout = R;
    *Cout = C;
    return true;
  }
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
  static const bool noSmall = getenv("MPDO_CHOL_NOSMALL") != nullptr;   // A/B knob: always the general kernel
  if (n <= CS_MAXN && batch <= 65535 && !noSmall) {
    static std::mutex smallMu;
    static bool configured = false;
    std::unique_lock<std::mutex> smallLock(smallMu);
    if (!configured) {
      if (cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)chol_small_smem(CS_MAXN, true)) != cudaSuccess) {
        cudaGetLastError();
        return 1;
      }
      configured = true;
    }
    smallLock.unlock();
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * 4 * (size_t)batch, st);
    if (e == cudaSuccess && X) e = cudaMemsetAsync(X, 0, sizeof(double2) * (size_t)batch * n * n, st);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "chol memset: %s", cudaGetErrorString(e));
      *rcOut = (int)e;
      return 0;
    }
    CholArgs a;
    a.n = n;
    a.R = n;
    a.rel = rel;
    a.G = (const double2*)G;
    a.Y = (double2*)Y;
    a.X = (double2*)X;
    a.slots = nullptr;
    a.info = info;
    a.cluster = 0;
    const double cflops = 8.0 * batch * ((double)n * n * n / 3.0) * (X ? 2.0 : 1.0);
    TimedLaunch timed(2, cflops, 16.0 * batch * (double)n * n * (X ? 3.0 : 2.0), st);
    chol_small_kernel<<<dim3(1, batch), CS_THREADS, chol_small_smem(n, X != nullptr), st>>>(a);
    *rcOut = check_launch("chol_small_kernel");
    return 0;
  }
  static const bool noCluster16 = getenv("MPDO_CHOL_NOCLUSTER16") != nullptr;   // A/B knob
  if (n > CS_MAXN && n <= CC_CTAS * CC_R && batch <= 65535 && !noCluster16) {
    static std::mutex ccMu;
    static int ccOk = -1;    // 1: usable (attributes set, a cluster of 16 fits), 0: not
    {
      std::lock_guard<std::mutex> lk(ccMu);
      if (ccOk < 0) {
        ccOk = 0;
        if (cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)chol_cluster_smem(CC_CTAS * CC_R, true)) == cudaSuccess &&
            cudaFuncSetAttribute(chol_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
          cudaLaunchConfig_t q = {};
          q.gridDim = dim3(CC_CTAS, 1);
          q.blockDim = dim3(CC_THREADS);
          q.dynamicSmemBytes = chol_cluster_smem(CC_CTAS * CC_R, true);
          cudaLaunchAttribute qa[1];
          qa[0].id = cudaLaunchAttributeClusterDimension;
          qa[0].val.clusterDim.x = CC_CTAS;
          qa[0].val.clusterDim.y = 1;
          qa[0].val.clusterDim.z = 1;
          q.attrs = qa;
          q.numAttrs = 1;
          int nClusters = 0;
          if (cudaOccupancyMaxActiveClusters(&nClusters, chol_cluster_kernel, &q) == cudaSuccess && nClusters >= 1)
            ccOk = 1;
        }
        if (!ccOk) cudaGetLastError();
      }
    }
    if (ccOk == 1) {
      cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * 4 * (size_t)batch, st);
      if (e == cudaSuccess && X) e = cudaMemsetAsync(X, 0, sizeof(double2) * (size_t)batch * n * n, st);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "chol memset: %s", cudaGetErrorString(e));
        *rcOut = (int)e;
        return 0;
      }
      CholArgs a;
      a.n = n;
      a.R = (n + CC_CTAS - 1) / CC_CTAS;
      a.rel = rel;
      a.G = (const double2*)G;
      a.Y = (double2*)Y;
      a.X = (double2*)X;
      a.slots = nullptr;
      a.info = info;
      a.cluster = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CC_CTAS, batch);
      cfg.blockDim = dim3(CC_THREADS);
      cfg.dynamicSmemBytes = chol_cluster_smem(n, X != nullptr);
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CC_CTAS;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      const double cflops = 8.0 * batch * ((double)n * n * n / 3.0) * (X ? 2.0 : 1.0);
      TimedLaunch timed(2, cflops, 16.0 * batch * (double)n * n * (X ? 3.0 : 2.0), st);
      e = cudaLaunchKernelEx(&cfg, chol_cluster_kernel, a);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "chol_cluster_kernel: %s", cudaGetErrorString(e));
        *rcOut = (int)e;
        return 0;
      }
      *rcOut = check_launch("chol_cluster_kernel");
      return 0;
    }
  }
  bool cluster = false;
  int optin = 0, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const bool oneCta = optin > 1024 && chol_smem(n, n, X != nullptr) <= (size_t)optin - 1024;
  if (!oneCta && chol_cluster_plan(n, X != nullptr, &R, &C)) {
    cluster = true;   // more than one CTA per matrix, at most 16: one thread-block cluster each, any batch size
    chunk = batch;
  } else {
    if (!chol_plan(batch, n, X != nullptr, &R, &C, &chunk)) return 1;
    if (C > 1 && (size_t)batch * 2 * C * (n + 1) * sizeof(double2) > slotsBytes) return 1;
  }
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
    a.cluster = cluster ? 1 : 0;
    // algorithmic work (SURVEY 8d convention, 8 real flops per complex multiply-add): n^3 / 3 for the factor and the
    // same again for the left inverse
    const double cflops = 8.0 * nb * ((double)n * n * n / 3.0) * (X ? 2.0 : 1.0);
    TimedLaunch timed(2, cflops, 16.0 * nb * (double)n * n * (X ? 3.0 : 2.0), st);
    if (cluster) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(C, nb);
      cfg.blockDim = dim3(CHOL_THREADS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, chol_kernel, a);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "chol_kernel (cluster of %d): %s", C, cudaGetErrorString(e));
        *rcOut = (int)e;
        return 0;
      }
    } else if (C == 1) {
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

namespace mpdo {
// Launch of the blocked unpivoted factorisation (factor only). Returns 1 when the shape / device cannot run it (the
// caller takes the pivoted kernels), 0 on success with *rcOut = 0, or 0 with a CUDA error code in *rcOut.
static int chol_blocked_launch(int batch, int n, const void* G, void* Y, int* info, double rel, cudaStream_t st,
                               int* rcOut) {
  *rcOut = 0;
  static const bool pivoted = getenv("MPDO_CHOL_PIVOTED") != nullptr;   // A/B knob: always the pivoted kernels
  static const int minN = getenv("MPDO_CHOL_BLOCKED_MIN") ? atoi(getenv("MPDO_CHOL_BLOCKED_MIN")) : 17;   // A/B knob
  if (pivoted || n < minN || n > CU_CTAS * CU_RMAX || batch > 65535) return 1;
  const int C = (n + CU_RMAX - 1) / CU_RMAX;   // CTAs per matrix = cluster size
  static std::mutex mu;
  static int usable[CU_CTAS + 1];   // per cluster size: 0 unknown, 1 usable, -1 not
  {
    std::lock_guard<std::mutex> lk(mu);
    static bool attrs = false;
    if (!attrs) {
      attrs = true;
      if (cudaFuncSetAttribute(chol_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)chol_blocked_smem(CU_CTAS * CU_RMAX)) != cudaSuccess ||
          cudaFuncSetAttribute(chol_blocked_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        for (int i = 0; i <= CU_CTAS; ++i) usable[i] = -1;
      }
    }
    if (usable[C] == 0) {
      usable[C] = -1;
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(C, 1);
      q.blockDim = dim3(CU_THREADS);
      q.dynamicSmemBytes = chol_blocked_smem(C * CU_RMAX);
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = C;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int nClusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nClusters, chol_blocked_kernel, &q) == cudaSuccess && nClusters >= 1)
        usable[C] = 1;
      else
        cudaGetLastError();
    }
    if (usable[C] != 1) return 1;
  }
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int) * 4 * (size_t)batch, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(Y, 0, sizeof(double2) * (size_t)batch * n * n, st);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "chol memset: %s", cudaGetErrorString(e));
    *rcOut = (int)e;
    return 0;
  }
  CholArgs a;
  a.n = n;
  a.R = CU_RMAX;
  a.rel = rel;
  a.G = (const double2*)G;
  a.Y = (double2*)Y;
  a.X = nullptr;
  a.slots = nullptr;
  a.info = info;
  a.cluster = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, batch);
  cfg.blockDim = dim3(CU_THREADS);
  cfg.dynamicSmemBytes = chol_blocked_smem(n);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  TimedLaunch timed(2, 8.0 * batch * ((double)n * n * n / 3.0), 16.0 * batch * (double)n * n * 2.0, st);
  e = cudaLaunchKernelEx(&cfg, chol_blocked_kernel, a);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "chol_blocked_kernel: %s", cudaGetErrorString(e));
    *rcOut = (int)e;
    return 0;
  }
  *rcOut = check_launch("chol_blocked_kernel");
  return 0;
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
    // precondition = 2: the factor only preconditions the Jacobi phase and the caller's data are fp32 - the blocked
    // factorisation without pivoting (see chol_blocked_kernel); shapes it cannot take fall through to the pivoted kernels
    if ((precondition == 2 && chol_blocked_launch(batch, n, G, base + l.y, (int*)(base + l.info), rel, st, &rc) == 0) ||
        chol_launch(batch, n, G, base + l.y, nullptr, base + l.slots, l.info - l.slots, (int*)(base + l.info), rel, st,
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
