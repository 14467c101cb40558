// Batched one-sided (Hestenes) Jacobi orthogonalisation of the rows of small complex matrices, the
// core of every SVD / Hermitian eigen-decomposition on the MPDO truncation path (sm_100a).
//
// A matrix Y (n rows) is partitioned into blocks of b rows. One CTA owns one pair of blocks: it
// stages the 2b rows in shared memory, one warp per row pair rotates them (dot products with
// warp-shuffle reductions, then the 2x2 unitary applied in place), and writes the rows back. Block
// pairs of one round are disjoint (round-robin tournament), so a round is one launch over
// (pairs, batch); a sweep is nbp-1 rounds. When the whole matrix fits one block pair the sweeps are
// looped inside the kernel and a decomposition is a single launch.
//
// Everything is fp64: B200 issues DFMA at half the FFMA rate, and fp64 cores make the complex64
// path insensitive to the conditioning of the Gram matrices it feeds in here.
#include <stdlib.h>

#include <mutex>

#include <cooperative_groups.h>

#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mpdo {

struct JacobiArgs {
  int n, m, mt, ld;
  long long batchStride;
  int b;       // rows per block (= warps per CTA)
  int nbp;     // number of blocks rounded up to even
  int round;   // tournament round (multi-launch mode)
  int sweep;   // sweep index (multi-launch mode)
  int maxSweeps;
  int loop;    // 1: whole matrix in this CTA, loop sweeps in-kernel
  double tol;
  int* cnt;    // [batch][WORK_INTS]: 32 per-sweep rotation counters, then one double = max row norm^2
  const int* rank;   // optional [batch * rankStride]: only the first rank[b] rows of matrix b are non-zero
  int rankStride;
};

constexpr int WORK_INTS = 48;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One-sided Jacobi converges quadratically: a sweep in which the largest relative inner product met is mu leaves
// nothing above ~mu^2 behind (measured on the Gram matrices of the sweeps, tools/jacobi_model.py: 9.2e-5 -> 2.5e-8 ->
// 1e-10). A rotation reports bit 1 when |<x,y>|^2 > JACOBI_QUAD tol |x|^2 |y|^2, i.e. mu > 0.3 sqrt(tol); a sweep
// without such a rotation has converged to tol and the sweep that would only confirm it is not run.
constexpr double JACOBI_QUAD = 0.09;

// Rotate rows x, y (shared memory) so that their first m entries become orthogonal; G lanes cooperate on one
// pair (32/G pairs per warp). The kernel is instruction-issue bound (ncu: ~2 IPC, half of the issue slots busy, 6
// warps per scheduler each spending ~440 instructions per rotation, most of them reduction / scalar / index
// overhead rather than row arithmetic), so short rows share that overhead between several pairs of one warp.
// Every lane of the warp runs the shuffles; `active` only predicates the memory traffic. Returns 0 if the pair was left
// alone, 1 after a rotation by a small angle, 3 after a large one (JACOBI_QUAD).
template <int G>
__device__ __forceinline__ int rotate_pair(double2* __restrict__ x, double2* __restrict__ y, bool active, int m,
                                           int mt, double tol,
                                           double floor2, int sub) {
  double a = 0, bq = 0, gr = 0, gi = 0;
  if (active) {
#pragma unroll 4
    for (int k = sub; k < m; k += G) {
      double2 u = x[k], v = y[k];
      a = fma(u.x, u.x, fma(u.y, u.y, a));
      bq = fma(v.x, v.x, fma(v.y, v.y, bq));
      // g = sum x * conj(y)
      gr = fma(u.x, v.x, fma(u.y, v.y, gr));
      gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    bq += __shfl_xor_sync(0xffffffffu, bq, o);
    gr += __shfl_xor_sync(0xffffffffu, gr, o);
    gi += __shfl_xor_sync(0xffffffffu, gi, o);
  }
  const double g2 = gr * gr + gi * gi;
  const double ab = a * bq;
  // rows that are both at rounding level of the matrix scale carry no information: leave them alone
  if (!active || !(ab > floor2) || !(g2 > tol * tol * ab)) return 0;
  // Rotation by the smaller angle theta with tan(2 theta) = |g| / |d|, d = (beta - alpha)/2. With h = sqrt(d^2+|g|^2):
  // cos(theta) = (h+|d|) / sqrt(2h(h+|d|)),  sin(theta) e^{i phi} = sign(d) g / sqrt(2h(h+|d|)):
  // two dependent slow fp64 operations (sqrt, rsqrt) instead of the textbook chain of seven.
  const double d = 0.5 * (bq - a);
  const double ad = fabs(d);
  const double h = sqrt(fma(d, d, g2));
  const double ru = rsqrt(2.0 * h * (h + ad));
  const double c = (h + ad) * ru;
  const double sg = d >= 0 ? ru : -ru;
  // x' = c x - s e^{i phi} y ; y' = s e^{-i phi} x + c y
  const double sr = sg * gr, si = sg * gi;
#pragma unroll 4
  for (int k = sub; k < mt; k += G) {
    double2 u = x[k], v = y[k];
    double2 xn, yn;
    xn.x = c * u.x - (sr * v.x - si * v.y);
    xn.y = c * u.y - (sr * v.y + si * v.x);
    yn.x = (sr * u.x + si * u.y) + c * v.x;
    yn.y = (sr * u.y - si * u.x) + c * v.y;
    x[k] = xn;
    y[k] = yn;
  }
  return (g2 > JACOBI_QUAD * tol * ab) ? 3 : 1;
}

// Round-robin partner tables: `np` players (even), round r in [0, np-1), pair q in [0, np/2). Division free.
__device__ __forceinline__ void rr_pair(int np, int r, int q, int& p0, int& p1) {
  const int w = np - 1;
  if (q == 0) {
    p0 = w;
    p1 = r >= w ? r - w : r;
  } else {
    p0 = r + q;
    if (p0 >= w) p0 -= w;
    p1 = r - q + w;
    if (p1 >= w) p1 -= w;
  }
}

template <int G>
__global__ void __launch_bounds__(1024) jacobi_kernel(JacobiArgs p, double2* __restrict__ Yall) {
  extern __shared__ double2 smem[];  // 2b rows of mt entries
  constexpr int PPW = 32 / G;        // pairs per warp
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int slot = warp * PPW + lane / G, nslots = nwarps * PPW, sub = lane % G;
  const int b = p.b, mt = p.mt;
  int* cnt = p.cnt + (long long)bidx * WORK_INTS;
  const double amax = *reinterpret_cast<const double*>(cnt + 32);
  const double floor2 = fmax(1e-290, 1e-48 * amax * amax);

  if (!p.loop && p.sweep > 0 && cnt[p.sweep - 1] == 0) return;  // converged in the previous sweep

  int I, J;
  if (p.loop) {
    I = 0;
    J = 1;
  } else {
    rr_pair(p.nbp, p.round, blockIdx.x, I, J);
  }
  double2* Y = Yall + (long long)bidx * p.batchStride;
  const int nrows = p.rank ? min(p.n, p.rank[(long long)bidx * p.rankStride]) : p.n;

  // stage the 2b rows
  for (int v = warp; v < 2 * b; v += nwarps) {
    const int row = (v < b) ? I * b + v : J * b + (v - b);
    double2* dst = smem + v * mt;
    if (row < nrows) {
      const double2* src = Y + (long long)row * p.ld;
      for (int k = lane; k < mt; k += 32) dst[k] = src[k];
    }
  }
  __syncthreads();

  const int be = (b + 1) & ~1;  // b rounded up to even for the intra-block tournament
  const int half = be / 2;
  const int nsweeps = p.loop ? p.maxSweeps : 1;
  for (int sw = 0; sw < nsweeps; ++sw) {
    int rot = 0;
    if (p.loop || p.round == 0) {
      // intra-block pairs of both blocks: be/2 pairs per block per step, be-1 steps
      for (int step = 0; step < be - 1; ++step) {
        for (int base = 0; base < be; base += nslots) {
          const int q = base + slot;
          const int blk = q >= half ? 1 : 0;
          int a0, a1;
          rr_pair(be, step, q - blk * half, a0, a1);
          const int first = (blk ? J : I) * b;
          const bool active = q < be && a0 < b && a1 < b && first + a0 < nrows && first + a1 < nrows;
          rot |= rotate_pair<G>(smem + (blk * b + a0) * mt, smem + (blk * b + a1) * mt, active, p.m, mt, p.tol,
                                floor2, sub);
        }
        __syncthreads();
      }
    }
    // cross pairs: row q of block I with row (q+step) mod b of block J
    for (int step = 0; step < b; ++step) {
      for (int base = 0; base < b; base += nslots) {
        const int q = base + slot;
        int a1 = q + step;
        if (a1 >= b) a1 -= b;
        const bool active = q < b && I * b + q < nrows && J * b + a1 < nrows;
        rot |= rotate_pair<G>(smem + q * mt, smem + (b + a1) * mt, active, p.m, mt, p.tol, floor2, sub);
      }
      __syncthreads();
    }
    if (p.loop) {
      const int any = __syncthreads_or(rot);
      if (threadIdx.x == 0) cnt[sw] = any;
      if (!any) break;
    } else {
      if (rot && sub == 0) atomicAdd(&cnt[p.sweep], 1);
    }
  }

  // write back
  for (int v = warp; v < 2 * b; v += nwarps) {
    const int row = (v < b) ? I * b + v : J * b + (v - b);
    if (row < nrows) {
      double2* dst = Y + (long long)row * p.ld;
      const double2* src = smem + v * mt;
      for (int k = lane; k < mt; k += 32) dst[k] = src[k];
    }
  }
}

// Persistent variant for matrices that span several block pairs: one cooperative launch runs every round of every
// sweep, the CTAs of a matrix meeting at a device-wide barrier after each round and reading the sweep's rotation
// count to stop together. Replaces ~(nbp-1) launches plus one host synchronisation per sweep of the multi-launch
// driver (34k launches per cfg2 layer in steady state), which is what bounded large cores.
template <int G>
__global__ void __launch_bounds__(1024) jacobi_persistent_kernel(JacobiArgs p, double2* __restrict__ Yall) {
  extern __shared__ double2 smem[];
  constexpr int PPW = 32 / G;
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int slot = warp * PPW + lane / G, nslots = nwarps * PPW, sub = lane % G;
  const int b = p.b, mt = p.mt;
  int* cnt = p.cnt + (long long)bidx * WORK_INTS;
  const double amax = *reinterpret_cast<const double*>(cnt + 32);
  const double floor2 = fmax(1e-290, 1e-48 * amax * amax);
  unsigned* bar = reinterpret_cast<unsigned*>(cnt + 34);
  int* errflag = cnt + 35;
  unsigned phase = 0;
  double2* Y = Yall + (long long)bidx * p.batchStride;
  const int nrows = p.rank ? min(p.n, p.rank[(long long)bidx * p.rankStride]) : p.n;
  if (nrows < 2) return;   // uniform over the CTAs of this matrix: nobody reaches a barrier
  // only the blocks that hold non-zero rows take part in the tournament
  const int nbp = (((nrows + b - 1) / b) + 1) & ~1;
  // gridDim.x CTAs share the nbp/2 block pairs of a round (one each when the matrix has the GPU to itself, all of
  // them in one CTA when the batch fills the GPU: then there is no device-wide barrier at all)
  const unsigned ncta = min((unsigned)(nbp / 2), gridDim.x);
  if (blockIdx.x >= ncta) return;
  const int be = (b + 1) & ~1;
  const int half = be / 2;

  for (int sw = 0; sw < p.maxSweeps; ++sw) {
    int rotSweep = 0;
    for (int round = 0; round < nbp - 1; ++round) {
      int rot = 0;
      for (int pairIdx = blockIdx.x; pairIdx < nbp / 2; pairIdx += ncta) {
        int I, J;
        rr_pair(nbp, round, pairIdx, I, J);
        for (int v = warp; v < 2 * b; v += nwarps) {  // stage (L2 reads: other SMs wrote these rows last round)
          const int row = (v < b) ? I * b + v : J * b + (v - b);
          double2* dst = smem + v * mt;
          if (row < nrows) {
            const double2* src = Y + (long long)row * p.ld;
            for (int k = lane; k < mt; k += 32) dst[k] = __ldcg(src + k);
          }
        }
        __syncthreads();
        if (round == 0) {
          for (int step = 0; step < be - 1; ++step) {
            for (int base = 0; base < be; base += nslots) {
              const int q = base + slot;
              const int blk = q >= half ? 1 : 0;
              int a0, a1;
              rr_pair(be, step, q - blk * half, a0, a1);
              const int first = (blk ? J : I) * b;
              const bool active = q < be && a0 < b && a1 < b && first + a0 < nrows && first + a1 < nrows;
              rot |= rotate_pair<G>(smem + (blk * b + a0) * mt, smem + (blk * b + a1) * mt, active, p.m, mt, p.tol,
                                    floor2, sub);
            }
            __syncthreads();
          }
        }
        for (int step = 0; step < b; ++step) {
          for (int base = 0; base < b; base += nslots) {
            const int q = base + slot;
            int a1 = q + step;
            if (a1 >= b) a1 -= b;
            const bool active = q < b && I * b + q < nrows && J * b + a1 < nrows;
            rot |= rotate_pair<G>(smem + q * mt, smem + (b + a1) * mt, active, p.m, mt, p.tol, floor2, sub);
          }
          __syncthreads();
        }
        for (int v = warp; v < 2 * b; v += nwarps) {  // write back
          const int row = (v < b) ? I * b + v : J * b + (v - b);
          if (row < nrows) {
            double2* dst = Y + (long long)row * p.ld;
            const double2* src = smem + v * mt;
            for (int k = lane; k < mt; k += 32) dst[k] = src[k];
          }
        }
        __syncthreads();   // the staging area is reused by the next pair of this CTA
      }
      if (ncta > 1) {
        if (rot && sub == 0) atomicAdd(&cnt[sw], 1);
        matrix_barrier(bar, ncta, phase, errflag);
      } else {
        rotSweep |= rot;
        __threadfence_block();
      }
    }
    if (ncta > 1) {
      if (*((volatile int*)&cnt[sw]) == 0 || *((volatile int*)errflag)) break;
    } else {
      const int any = __syncthreads_or(rotSweep);
      if (threadIdx.x == 0) cnt[sw] = any;
      if (!any) break;
    }
  }
}

// Persistent kernel for LONG rows with the block pair held in REGISTERS, sliced by columns (mt in 257..1024 for blocks of
// 4 rows, 257..512 for blocks of 8). The kernels above move every row pair through shared memory six times per rotation
// (ncu: they are bound by the 128 B/clk of shared memory, ~1.6 us per step of eight 8 KB row pairs), and with one warp
// per pair a CTA has only b warps. Here thread t of 512 owns columns t, t + 512 of ALL 2b rows of the block pair, so a
// rotation step costs no row traffic at all: every thread forms the partial sums of the b disjoint pairs of the step
// over its columns, the 4b sums are reduced across the warp with a halving butterfly (4b shuffles instead of 20b), meet
// the other warps' in 2 KB of shared memory (one CTA barrier per step, two buffers), lanes 0..b-1 of every warp work
// out the b rotations from the same totals (bitwise identical in every warp) and broadcast them, and the threads update
// their columns. Pairings are compile-time (i, i ^ s), which keeps the rows in registers. Rows come from and go back to
// L2 once per round. Same tournament over blocks, device-wide barrier per round and convergence flags as
// jacobi_persistent_kernel; the sweep that would only confirm convergence is not run (JACOBI_QUAD).
constexpr int JR_THREADS = 512, JR_WARPS = JR_THREADS / 32;

// index of the t-th pair (i, i ^ s), i < i ^ s, among 0..BR-1 (compile-time after unrolling)
template <int BR>
__device__ __forceinline__ constexpr int xor_pair_low(int s, int t) {
  for (int c = 0; c < BR; ++c)
    if (c < (c ^ s)) {
      if (t == 0) return c;
      --t;
    }
  return 0;
}

template <int BR, int CPT>
__global__ void __launch_bounds__(JR_THREADS, 1) jacobi_persistent_cols_kernel(JacobiArgs p, double2* __restrict__ Yall) {
  constexpr int V = 4 * BR;                            // sums per step: (|x|^2, |y|^2, Re g, Im g) of BR pairs
  static_assert(V == 16 || V == 32, "blocks of 4 or 8 rows");
  __shared__ double part[2][JR_WARPS][V];
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = p.mt, m = p.m;
  int* cnt = p.cnt + (long long)bidx * WORK_INTS;
  const double amax = *reinterpret_cast<const double*>(cnt + 32);
  const double floor2 = fmax(1e-290, 1e-48 * amax * amax);
  unsigned* bar = reinterpret_cast<unsigned*>(cnt + 34);
  int* errflag = cnt + 35;
  unsigned phase = 0;
  double2* Y = Yall + (long long)bidx * p.batchStride;
  const int nrows = p.rank ? min(p.n, p.rank[(long long)bidx * p.rankStride]) : p.n;
  if (nrows < 2) return;   // uniform over the CTAs of this matrix: nobody reaches a barrier
  const int nbp = (((nrows + BR - 1) / BR) + 1) & ~1;
  const unsigned ncta = (unsigned)(nbp / 2);   // the launch gives every block pair of a round its own CTA
  if (blockIdx.x >= ncta) return;
  const double tol = p.tol, tol2 = tol * tol;
  const int myIdx = V == 32 ? lane : (lane >> 1);      // which of the V sums this lane holds after the butterfly
  int buf = 0;

  double2 r[2 * BR][CPT];                              // rows 0..BR-1: block I, BR..2BR-1: block J

  // one step over BR disjoint pairs (ia[q], ib[q]), register indices known at compile time
#define JR_STEP(IA, IB, ROT)                                                                         \
  {                                                                                                  \
    double v[V];                                                                                     \
    _Pragma("unroll") for (int q = 0; q < BR; ++q) {                                                 \
      const int ia = IA, ib = IB;                                                                    \
      double a = 0, bq = 0, gr = 0, gi = 0;                                                          \
      _Pragma("unroll") for (int j = 0; j < CPT; ++j) {                                              \
        if ((int)threadIdx.x + j * JR_THREADS < m) {                                                 \
          const double2 u = r[ia][j], w = r[ib][j];                                                  \
          a = fma(u.x, u.x, fma(u.y, u.y, a));                                                       \
          bq = fma(w.x, w.x, fma(w.y, w.y, bq));                                                     \
          gr = fma(u.x, w.x, fma(u.y, w.y, gr));                                                     \
          gi = fma(u.y, w.x, fma(-u.x, w.y, gi));                                                    \
        }                                                                                            \
      }                                                                                              \
      v[4 * q + 0] = a;                                                                              \
      v[4 * q + 1] = bq;                                                                             \
      v[4 * q + 2] = gr;                                                                             \
      v[4 * q + 3] = gi;                                                                             \
    }                                                                                                \
    /* halving butterfly: after the exchange over lane bit `off` a lane keeps half of its sums */   \
    _Pragma("unroll") for (int off = 16, c = V / 2; c >= 1; off >>= 1, c >>= 1) {                    \
      const bool upper = (lane & off) != 0;                                                          \
      _Pragma("unroll") for (int i = 0; i < c; ++i) {                                                \
        const double send = upper ? v[i] : v[i + c];                                                 \
        const double keep = upper ? v[i + c] : v[i];                                                 \
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);                                       \
      }                                                                                              \
    }                                                                                                \
    if (V == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);                                      \
    part[buf][warp][myIdx] = v[0];                                                                   \
    __syncthreads();                                                                                 \
    /* totals (lane l: sum number l mod V), the same order of additions in every warp */             \
    double tot;                                                                                      \
    {                                                                                                \
      const int idx = lane & (V - 1);                                                                \
      double tw[JR_WARPS];                                                                           \
      _Pragma("unroll") for (int w = 0; w < JR_WARPS; ++w) tw[w] = part[buf][w][idx];                \
      _Pragma("unroll") for (int h = JR_WARPS / 2; h >= 1; h >>= 1) {  /* fixed tree: 4 levels */    \
        _Pragma("unroll") for (int w = 0; w < h; ++w) tw[w] += tw[w + h];                            \
      }                                                                                              \
      tot = tw[0];                                                                                   \
    }                                                                                                \
    buf ^= 1;                                                                                        \
    /* lane q (mod BR) works out the rotation of pair q */                                           \
    const int ql = lane & (BR - 1);                                                                  \
    const double a_ = __shfl_sync(0xffffffffu, tot, 4 * ql + 0);                                     \
    const double b_ = __shfl_sync(0xffffffffu, tot, 4 * ql + 1);                                     \
    const double gr_ = __shfl_sync(0xffffffffu, tot, 4 * ql + 2);                                    \
    const double gi_ = __shfl_sync(0xffffffffu, tot, 4 * ql + 3);                                    \
    const double g2 = gr_ * gr_ + gi_ * gi_;                                                         \
    const double ab = a_ * b_;                                                                       \
    double c_ = 1.0, sr_ = 0.0, si_ = 0.0;                                                           \
    int flag = 0;                                                                                    \
    if ((ab > floor2) && (g2 > tol2 * ab)) {                                                         \
      const double d = 0.5 * (b_ - a_);                                                              \
      const double ad = fabs(d);                                                                     \
      const double h = sqrt(fma(d, d, g2));                                                          \
      const double ru = rsqrt(2.0 * h * (h + ad));                                                   \
      c_ = (h + ad) * ru;                                                                            \
      const double sg = d >= 0 ? ru : -ru;                                                           \
      sr_ = sg * gr_;                                                                                \
      si_ = sg * gi_;                                                                                \
      flag = (g2 > JACOBI_QUAD * tol * ab) ? 3 : 1;                                                  \
    }                                                                                                \
    _Pragma("unroll") for (int q = 0; q < BR; ++q) {                                                 \
      const int fl = __shfl_sync(0xffffffffu, flag, q);                                              \
      const double cc = __shfl_sync(0xffffffffu, c_, q);                                             \
      const double srr = __shfl_sync(0xffffffffu, sr_, q);                                           \
      const double sii = __shfl_sync(0xffffffffu, si_, q);                                           \
      ROT |= fl;                                                                                     \
      if (fl) { /* warp-uniform */                                                                   \
        const int ia = IA, ib = IB;                                                                  \
        _Pragma("unroll") for (int j = 0; j < CPT; ++j) {                                            \
          const double2 u = r[ia][j], w = r[ib][j];                                                  \
          double2 xn, yn;                                                                            \
          xn.x = cc * u.x - (srr * w.x - sii * w.y);                                                 \
          xn.y = cc * u.y - (srr * w.y + sii * w.x);                                                 \
          yn.x = (srr * u.x + sii * u.y) + cc * w.x;                                                 \
          yn.y = (srr * u.y - sii * u.x) + cc * w.y;                                                 \
          r[ia][j] = xn;                                                                             \
          r[ib][j] = yn;                                                                             \
        }                                                                                            \
      }                                                                                              \
    }                                                                                                \
  }

  for (int sw = 0; sw < p.maxSweeps; ++sw) {
    for (int round = 0; round < nbp - 1; ++round) {
      int rot = 0;
      int I, J;
      rr_pair(nbp, round, (int)blockIdx.x, I, J);
      // rows -> registers (L2 reads: other SMs wrote these rows last round); rows / columns beyond the matrix are zero
#pragma unroll
      for (int v = 0; v < 2 * BR; ++v) {
        const int row = (v < BR) ? I * BR + v : J * BR + (v - BR);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int k = (int)threadIdx.x + j * JR_THREADS;
          r[v][j] = (row < nrows && k < mt) ? __ldcg(Y + (long long)row * p.ld + k) : make_double2(0.0, 0.0);
        }
      }
      if (round == 0) {   // once per sweep: the pairs inside each of the two blocks, (i, i ^ s) with i < i ^ s
#pragma unroll
        for (int s = 1; s < BR; ++s) {
          JR_STEP(((q >= BR / 2 ? BR : 0) + xor_pair_low<BR>(s, q % (BR / 2))),
                  ((q >= BR / 2 ? BR : 0) + (xor_pair_low<BR>(s, q % (BR / 2)) ^ s)), rot)
        }
      }
#pragma unroll
      for (int s = 0; s < BR; ++s) {
        JR_STEP(q, (BR + (q ^ s)), rot)
      }
#pragma unroll
      for (int v = 0; v < 2 * BR; ++v) {   // write back
        const int row = (v < BR) ? I * BR + v : J * BR + (v - BR);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const int k = (int)threadIdx.x + j * JR_THREADS;
          if (row < nrows && k < mt) Y[(long long)row * p.ld + k] = r[v][j];
        }
      }
      if (ncta > 1) {
        if ((rot & 2) && threadIdx.x == 0) atomicAdd(&cnt[sw], 1);
        matrix_barrier(bar, ncta, phase, errflag);
      } else {
        if ((rot & 2) && threadIdx.x == 0) cnt[sw] = 1;
        __threadfence_block();
        __syncthreads();
      }
    }
    if (*((volatile int*)&cnt[sw]) == 0 || *((volatile int*)errflag)) break;
  }
#undef JR_STEP
}

// Rotation of a register-resident row x (E entries per lane, entry k = lane + 32 e) against a row y in shared
// memory: y is streamed twice (inner products, then the update) and x never leaves the register file, which halves
// the shared-memory traffic of rotate_pair (3 row passes instead of 6). One warp per pair; `active` is warp-uniform.
template <int E>
__device__ __forceinline__ int rotate_reg(double2 (&x)[E], double2* __restrict__ y, bool active, int m, int mt,
                                          double tol, double floor2, int lane) {
  if (!active) return 0;
  double a = 0, bq = 0, gr = 0, gi = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = lane + 32 * e;
    if (k < m) {
      const double2 u = x[e], v = y[k];
      a = fma(u.x, u.x, fma(u.y, u.y, a));
      bq = fma(v.x, v.x, fma(v.y, v.y, bq));
      gr = fma(u.x, v.x, fma(u.y, v.y, gr));
      gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    bq += __shfl_xor_sync(0xffffffffu, bq, o);
    gr += __shfl_xor_sync(0xffffffffu, gr, o);
    gi += __shfl_xor_sync(0xffffffffu, gi, o);
  }
  const double g2 = gr * gr + gi * gi;
  const double ab = a * bq;
  if (!(ab > floor2) || !(g2 > tol * tol * ab)) return 0;
  const double d = 0.5 * (bq - a);
  const double ad = fabs(d);
  const double h = sqrt(fma(d, d, g2));
  const double ru = rsqrt(2.0 * h * (h + ad));
  const double c = (h + ad) * ru;
  const double sg = d >= 0 ? ru : -ru;
  const double sr = sg * gr, si = sg * gi;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = lane + 32 * e;
    if (k < mt) {
      const double2 u = x[e], v = y[k];
      double2 xn, yn;
      xn.x = c * u.x - (sr * v.x - si * v.y);
      xn.y = c * u.y - (sr * v.y + si * v.x);
      yn.x = (sr * u.x + si * u.y) + c * v.x;
      yn.y = (sr * u.y - si * u.x) + c * v.y;
      x[e] = xn;
      y[k] = yn;
    }
  }
  return 1;
}

// Persistent kernel with register-resident rows for long rows (mt <= 32 E): during the cross-pair steps of a round
// warp q keeps row q of block I in registers and meets the rows of block J one after the other in shared memory.
template <int E, int MAXT>
__global__ void __launch_bounds__(MAXT) jacobi_persistent_reg_kernel(JacobiArgs p, double2* __restrict__ Yall) {
  extern __shared__ double2 smem[];
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // nwarps == b
  const int b = p.b, mt = p.mt;
  int* cnt = p.cnt + (long long)bidx * WORK_INTS;
  const double amax = *reinterpret_cast<const double*>(cnt + 32);
  const double floor2 = fmax(1e-290, 1e-48 * amax * amax);
  unsigned* bar = reinterpret_cast<unsigned*>(cnt + 34);
  int* errflag = cnt + 35;
  unsigned phase = 0;
  double2* Y = Yall + (long long)bidx * p.batchStride;
  const int nrows = p.rank ? min(p.n, p.rank[(long long)bidx * p.rankStride]) : p.n;
  if (nrows < 2) return;
  const int nbp = (((nrows + b - 1) / b) + 1) & ~1;
  const unsigned ncta = (unsigned)(nbp / 2);
  if (blockIdx.x >= ncta) return;
  const int be = (b + 1) & ~1;
  const int half = be / 2;

  for (int sw = 0; sw < p.maxSweeps; ++sw) {
    for (int round = 0; round < nbp - 1; ++round) {
      int I, J;
      rr_pair(nbp, round, blockIdx.x, I, J);
      for (int v = warp; v < 2 * b; v += nwarps) {
        const int row = (v < b) ? I * b + v : J * b + (v - b);
        double2* dst = smem + v * mt;
        if (row < nrows) {
          const double2* src = Y + (long long)row * p.ld;
          for (int k = lane; k < mt; k += 32) dst[k] = __ldcg(src + k);
        }
      }
      __syncthreads();
      int rot = 0;
      if (round == 0) {   // pairs inside each block: both rows in shared memory
        for (int step = 0; step < be - 1; ++step) {
          for (int base = 0; base < be; base += nwarps) {
            const int q = base + warp;
            const int blk = q >= half ? 1 : 0;
            int a0, a1;
            rr_pair(be, step, q - blk * half, a0, a1);
            const int first = (blk ? J : I) * b;
            const bool active = q < be && a0 < b && a1 < b && first + a0 < nrows && first + a1 < nrows;
            rot |= rotate_pair<32>(smem + (blk * b + a0) * mt, smem + (blk * b + a1) * mt, active, p.m, mt, p.tol,
                                   floor2, lane);
          }
          __syncthreads();
        }
      }
      // cross pairs: warp q holds row q of block I
      const bool mine = warp < b && I * b + warp < nrows;
      double2 x[E];
      {
        const double2* xr = smem + warp * mt;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int k = lane + 32 * e;
          x[e] = (mine && k < mt) ? xr[k] : make_double2(0.0, 0.0);
        }
      }
      for (int step = 0; step < b; ++step) {
        int a1 = warp + step;
        if (a1 >= b) a1 -= b;
        const bool active = mine && J * b + a1 < nrows;
        rot |= rotate_reg<E>(x, smem + (b + a1) * mt, active, p.m, mt, p.tol, floor2, lane);
        __syncthreads();
      }
      if (mine) {   // write row q of block I back from the registers
        double2* dst = Y + (long long)(I * b + warp) * p.ld;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int k = lane + 32 * e;
          if (k < mt) dst[k] = x[e];
        }
      }
      for (int v = warp; v < b; v += nwarps) {   // block J from shared memory
        const int row = J * b + v;
        if (row < nrows) {
          double2* dst = Y + (long long)row * p.ld;
          const double2* src = smem + (b + v) * mt;
          for (int k = lane; k < mt; k += 32) dst[k] = src[k];
        }
      }
      if (rot && lane == 0) atomicAdd(&cnt[sw], 1);
      matrix_barrier(bar, ncta, phase, errflag);
    }
    if (*((volatile int*)&cnt[sw]) == 0 || *((volatile int*)errflag)) break;
  }
}

// ---- cluster-resident tournament (2b < n <= 32 b): the matrix never leaves the SMs ----------------------------------
// One thread-block cluster of nbp/2 <= 16 CTAs per matrix; CTA q keeps the two row blocks of tournament pair q in its
// shared memory for the whole decomposition. Between rounds the blocks travel to their next CTA through distributed
// shared memory (each warp pulls one row of each block into registers, cluster barrier, then stores it locally), so
// there is no L2 round trip and no device-wide barrier, and any batch size can be launched (clusters are scheduled
// independently: nothing has to be co-resident beyond one cluster).
// The cross-pair steps are register-resident: warp q holds row q of block I in registers for the whole round and
// meets the rows of block J one after the other; a row of J is loaded from shared memory ONCE per step, rotated in
// registers and stored once (2 row passes per rotation instead of the 6 of rotate_pair). ncu on the persistent kernel
// showed the steps bound by shared-memory bandwidth (196 KB per step and SM, 1.5 k of ~2.5 k cycles) and 30% of the
// time in L2 staging and the device-wide barrier; both go away here.

// Sum four doubles over the warp with 10 double shuffles instead of 20: the first two levels halve the number of
// values a lane carries while halving the lanes that carry each, the last three are plain butterflies, then the four
// totals are broadcast from their lane classes.
__device__ __forceinline__ void warp_sum4(double& v0, double& v1, double& v2, double& v3, int lane) {
  const bool h4 = (lane & 16) != 0;
  double s0 = h4 ? v0 : v2, s1 = h4 ? v1 : v3;
  double k0 = h4 ? v2 : v0, k1 = h4 ? v3 : v1;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  const bool h3 = (lane & 8) != 0;
  const double s = h3 ? k0 : k1;
  double k = h3 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  v0 = __shfl_sync(0xffffffffu, k, 0);
  v1 = __shfl_sync(0xffffffffu, k, 8);
  v2 = __shfl_sync(0xffffffffu, k, 16);
  v3 = __shfl_sync(0xffffffffu, k, 24);
}

// Rotation of the register-resident row x (entry k = lane + 32 e) against row y of shared memory: y is loaded once,
// rotated in registers and stored once. `active` is warp-uniform. Entries of x beyond mt are zero and stay zero.
template <int E>
__device__ __forceinline__ int rotate_regs(double2 (&x)[E], double2* __restrict__ yrow, bool active, int m, int mt,
                                           double tol, double floor2, int lane) {
  if (!active) return 0;
  double2 y[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = lane + 32 * e;
    y[e] = k < mt ? yrow[k] : make_double2(0.0, 0.0);
  }
  double a = 0, bq = 0, gr = 0, gi = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (lane + 32 * e < m) {
      const double2 u = x[e], v = y[e];
      a = fma(u.x, u.x, fma(u.y, u.y, a));
      bq = fma(v.x, v.x, fma(v.y, v.y, bq));
      gr = fma(u.x, v.x, fma(u.y, v.y, gr));
      gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
    }
  }
  warp_sum4(a, bq, gr, gi, lane);
  const double g2 = gr * gr + gi * gi;
  const double ab = a * bq;
  if (!(ab > floor2) || !(g2 > tol * tol * ab)) return 0;
  const double d = 0.5 * (bq - a);
  const double ad = fabs(d);
  const double qq = fma(d, d, g2);
  const double h = qq * rsqrt(qq);   // g2 > 0 here; a couple of ulps off the rounded root is immaterial for the angle
  const double ru = rsqrt(2.0 * h * (h + ad));
  const double c = (h + ad) * ru;
  const double sg = d >= 0 ? ru : -ru;
  const double sr = sg * gr, si = sg * gi;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = lane + 32 * e;
    const double2 u = x[e], v = y[e];
    double2 xn, yn;
    xn.x = fma(c, u.x, fma(-sr, v.x, si * v.y));
    xn.y = fma(c, u.y, -fma(sr, v.y, si * v.x));
    yn.x = fma(c, v.x, fma(sr, u.x, si * u.y));
    yn.y = fma(c, v.y, fma(sr, u.y, -si * u.x));
    x[e] = xn;
    if (k < mt) yrow[k] = yn;
  }
  return (g2 > JACOBI_QUAD * tol * ab) ? 3 : 1;
}

constexpr int JC_MAXC = 16;   // CTAs per cluster (non-portable above 8)

// Schedule: odd-even transposition over the nbp block positions (CTA q owns positions 2q and 2q+1, buffers A and B).
// Even rounds pair the two local blocks and swap their positions (a pointer swap); odd rounds pair position 2q-1 -
// CTA q pulls the B block of its left neighbour straight into registers - with the local A block, after which the
// pulled block stays (new A) and the old A rows are pushed into the neighbour's B buffer. Every pair of blocks meets
// exactly once in nbp rounds from any starting arrangement, and only ONE block per CTA crosses the cluster per two
// rounds in each direction (the round-robin tournament moves both blocks of every CTA every round; distributed shared
// memory moves ~20 B per clock and SM, which made that exchange a quarter of the run time), behind one cluster
// barrier per round.
template <int E>
__global__ void __launch_bounds__(256, 1) jacobi_cluster_kernel(JacobiArgs p, double2* __restrict__ Yall) {
  extern __shared__ double2 smem[];   // [2 buffers][b rows][mt]
  __shared__ int flags[JC_MAXC];      // per-CTA "rotated in this sweep", written by every CTA of the cluster
  __shared__ int perm[2 * JC_MAXC];   // block index at every position (the same simulation in every CTA)
  cg::cluster_group cl = cg::this_cluster();
  const int q = (int)blockIdx.x;      // cluster = the gridDim.x CTAs of one matrix: rank in cluster == blockIdx.x
  const int C = (int)gridDim.x;
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // b warps
  const int b = p.b, mt = p.mt;
  int* cnt = p.cnt + (long long)bidx * WORK_INTS;
  const double amax = *reinterpret_cast<const double*>(cnt + 32);
  const double floor2 = fmax(1e-290, 1e-48 * amax * amax);
  double2* Y = Yall + (long long)bidx * p.batchStride;
  const int nrows = p.rank ? min(p.n, p.rank[(long long)bidx * p.rankStride]) : p.n;
  if (nrows < 2) return;              // uniform over the cluster: nobody reaches a barrier
  const int nbp = (((nrows + b - 1) / b) + 1) & ~1;   // blocks that hold non-zero rows, rounded up to even
  const int Ca = nbp / 2;             // CTAs that take part (<= C by construction of the launch)
  const bool act = q < Ca;
  const int be = (b + 1) & ~1, half = be / 2;
  double2* bufA = smem;
  double2* bufB = smem + (size_t)b * mt;
  if (threadIdx.x < JC_MAXC) flags[threadIdx.x] = 0;
  if (threadIdx.x < 2 * JC_MAXC) perm[threadIdx.x] = threadIdx.x;
  if (act) {
    for (int v = warp; v < 2 * b; v += b) {
      const int row = 2 * q * b + v;   // block 2q -> A, block 2q+1 -> B
      double2* dst = smem + (size_t)v * mt;
      if (row < nrows) {
        const double2* src = Y + (long long)row * p.ld;
        for (int k = lane; k < mt; k += 32) dst[k] = src[k];
      } else {
        for (int k = lane; k < mt; k += 32) dst[k] = make_double2(0.0, 0.0);
      }
    }
  }
  __syncthreads();

  double2 x[E];
  for (int sw = 0; sw < p.maxSweeps; ++sw) {
    int rot = 0;
    for (int round = 0; round < nbp; ++round) {
      if ((round & 1) == 0) {
        // ---- even round: the two local blocks ------------------------------------------------------------------
        if (act) {
          const int idA = perm[2 * q], idB = perm[2 * q + 1];
          if (round == 0) {   // once per sweep: the pairs inside each of the two blocks, both rows in shared memory
            for (int step = 0; step < be - 1; ++step) {
              for (int base = 0; base < be; base += b) {
                const int qq = base + warp;
                const int blk = qq >= half ? 1 : 0;
                int a0, a1;
                rr_pair(be, step, qq - blk * half, a0, a1);
                const int firstRow = (blk ? idB : idA) * b;
                double2* rows = blk ? bufB : bufA;
                const bool on = qq < be && a0 < b && a1 < b && firstRow + a0 < nrows && firstRow + a1 < nrows;
                rot |= rotate_pair<32>(rows + (size_t)a0 * mt, rows + (size_t)a1 * mt, on, p.m, mt, p.tol, floor2, lane);
              }
              __syncthreads();
            }
          }
          double2* xs = bufA + (size_t)warp * mt;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            x[e] = k < mt ? xs[k] : make_double2(0.0, 0.0);
          }
          const bool mine = idA * b + warp < nrows;
          for (int step = 0; step < b; ++step) {
            int a1 = warp + step;
            if (a1 >= b) a1 -= b;
            const bool on = mine && idB * b + a1 < nrows;
            rot |= rotate_regs<E>(x, bufB + (size_t)a1 * mt, on, p.m, mt, p.tol, floor2, lane);
            __syncthreads();
          }
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            if (k < mt) xs[k] = x[e];
          }
        }
        {   // the two positions of every CTA swap (same pointer parity in every CTA of the cluster)
          double2* t = bufA;
          bufA = bufB;
          bufB = t;
        }
        __syncthreads();   // perm was read above
        if (threadIdx.x < Ca) {
          const int t = perm[2 * threadIdx.x];
          perm[2 * threadIdx.x] = perm[2 * threadIdx.x + 1];
          perm[2 * threadIdx.x + 1] = t;
        }
      } else {
        // ---- odd round: position 2q-1 (B of the left neighbour) against position 2q (local A) --------------------
        if (act && q >= 1) {
          const int idX = perm[2 * q - 1], idY = perm[2 * q];
          double2* remoteB = cl.map_shared_rank(bufB, q - 1) + (size_t)warp * mt;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            x[e] = k < mt ? remoteB[k] : make_double2(0.0, 0.0);
          }
          const bool mine = idX * b + warp < nrows;
          for (int step = 0; step < b; ++step) {
            int a1 = warp + step;
            if (a1 >= b) a1 -= b;
            const bool on = mine && idY * b + a1 < nrows;
            rot |= rotate_regs<E>(x, bufA + (size_t)a1 * mt, on, p.m, mt, p.tol, floor2, lane);
            __syncthreads();
          }
          // the rotated local block moves to position 2q-1 (the neighbour's B), the pulled block stays as the new A
          double2* ys = bufA + (size_t)warp * mt;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int k = lane + 32 * e;
            if (k < mt) {
              remoteB[k] = ys[k];
              ys[k] = x[e];
            }
          }
        }
        __syncthreads();   // perm was read above
        if (threadIdx.x + 1 < Ca) {
          const int t = perm[2 * threadIdx.x + 1];
          perm[2 * threadIdx.x + 1] = perm[2 * threadIdx.x + 2];
          perm[2 * threadIdx.x + 2] = t;
        }
      }
      if (round == nbp - 1) {   // end of the sweep: tell every CTA of the cluster whether this one rotated
        const int anyRot = __syncthreads_or(rot & 1), anyBig = __syncthreads_or(rot & 2);
        if (threadIdx.x < C) cl.map_shared_rank(flags, threadIdx.x)[q] = (anyRot ? 1 : 0) | (anyBig ? 2 : 0);
      }
      cl.sync();   // parked / pushed rows (and the sweep flags) are visible cluster-wide; perm is updated
    }
    int anyAll = 0;
    for (int c = 0; c < C; ++c) anyAll |= flags[c];
    if (q == 0 && threadIdx.x == 0) cnt[sw] = anyAll;
    static_assert(JACOBI_QUAD > 0, "");
    if (!(anyAll & 1) || !(anyAll & 2)) break;   // no rotation, or small ones only (quadratic convergence): the same
                                                 // flags in every CTA
  }

  if (act) {
    for (int v = warp; v < 2 * b; v += b) {
      const int row = perm[2 * q + (v < b ? 0 : 1)] * b + (v < b ? v : v - b);
      if (row < nrows) {
        double2* dst = Y + (long long)row * p.ld;
        const double2* src = (v < b ? bufA : bufB) + (size_t)(v < b ? v : v - b) * mt;
        for (int k = lane; k < mt; k += 32) dst[k] = src[k];
      }
    }
  }
}

// max over rows of |row[0..m)|^2, stored as a double after the 32 sweep counters of each batch entry
__global__ void __launch_bounds__(256) row_norm_max_kernel(int n, int m, int ld, long long batchStride,
                                                           const double2* __restrict__ Yall, int* __restrict__ work) {
  // one warp per row; non-negative doubles order like their bit patterns, so the maximum is one 64-bit atomicMax
  // (the slot was zeroed with the rest of `work`)
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const double2* row = Yall + (long long)bidx * batchStride + (long long)r * ld;
  double a = 0;
  for (int k = lane; k < m; k += 32) {
    double2 u = row[k];
    a = fma(u.x, u.x, fma(u.y, u.y, a));
  }
  a = warp_sum(a);
  if (lane == 0 && a > 0)
    atomicMax(reinterpret_cast<unsigned long long*>(work + (long long)bidx * WORK_INTS + 32),
              (unsigned long long)__double_as_longlong(a));
}

// ---- finalize: norms, descending sort, optional normalised rows / accumulator ------------------
__global__ void __launch_bounds__(256) rows_finalize_kernel(int n, int m, int mz, int ld, long long batchStride,
                                                            const double2* __restrict__ Yall, double* __restrict__ sOut,
                                                            double2* __restrict__ Yn, double2* __restrict__ Z,
                                                            int normalize, double zeroTol) {
  extern __shared__ double fsm[];  // norms[n], then int rank[n]
  double* norms = fsm;
  int* rank = (int*)(fsm + n);
  const int bidx = blockIdx.x;
  const double2* Y = Yall + (long long)bidx * batchStride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < n; r += nw) {
    const double2* row = Y + (long long)r * ld;
    double a = 0;
    for (int k = lane; k < m; k += 32) {
      double2 u = row[k];
      a = fma(u.x, u.x, fma(u.y, u.y, a));
    }
    a = warp_sum(a);
    if (lane == 0) norms[r] = sqrt(a);
  }
  __syncthreads();
  // rank of every row in the descending order (ties: lower row first): one warp per row counts in parallel
  for (int r = warp; r < n; r += nw) {
    const double v = norms[r];
    int rk = 0;
    for (int q = lane; q < n; q += 32) {
      const double w = norms[q];
      rk += (w > v) || (w == v && q < r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    if (lane == 0) {
      rank[r] = rk;
      sOut[(long long)bidx * n + rk] = (normalize & 2) ? v * v : v;
    }
  }
  __syncthreads();
  double vmax = 0;
  for (int q = lane; q < n; q += 32) vmax = fmax(vmax, norms[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  for (int r = warp; r < n; r += nw) {
    const double2* row = Y + (long long)r * ld;
    const int rk = rank[r];
    const double nv = norms[r];
    if (Yn) {
      double sc = 1.0;
      if (normalize & 1) sc = (nv > zeroTol * vmax && nv > 0) ? 1.0 / nv : 0.0;
      double2* dst = Yn + ((long long)bidx * n + rk) * m;
      for (int k = lane; k < m; k += 32) {
        double2 u = row[k];
        u.x *= sc;
        u.y *= sc;
        dst[k] = u;
      }
    }
    if (Z) {
      double2* dst = Z + ((long long)bidx * n + rk) * mz;
      for (int k = lane; k < mz; k += 32) dst[k] = row[m + k];
    }
  }
}

// ---- finalize for large matrices: the same result from three gridded kernels (a single CTA walking a 1024 x 1024
// matrix twice was 0.5 ms on the critical path of every decomposition) ------------------------------------------
__global__ void __launch_bounds__(256) row_norms_kernel(int n, int m, int ld, long long batchStride,
                                                        const double2* __restrict__ Yall, double* __restrict__ sOut) {
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const double2* row = Yall + (long long)bidx * batchStride + (long long)r * ld;
  double a = 0;
  for (int k = lane; k < m; k += 32) {
    const double2 u = row[k];
    a = fma(u.x, u.x, fma(u.y, u.y, a));
  }
  a = warp_sum(a);
  if (lane == 0) sOut[(long long)bidx * n + r] = sqrt(a);   // unsorted for now
}

__global__ void __launch_bounds__(256) rows_scatter_kernel(int n, int m, int mz, int ld, long long batchStride,
                                                           const double2* __restrict__ Yall,
                                                           const double* __restrict__ normsIn, double2* __restrict__ Yn,
                                                           double2* __restrict__ Z, int normalize, double zeroTol) {
  extern __shared__ double fsm[];  // norms[n]
  const int bidx = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = threadIdx.x; q < n; q += blockDim.x) fsm[q] = normsIn[(long long)bidx * n + q];
  __syncthreads();
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  const double v = fsm[r];
  int rk = 0;
  double vmax = 0;
  for (int q = lane; q < n; q += 32) {
    const double w = fsm[q];
    rk += (w > v) || (w == v && q < r);
    vmax = fmax(vmax, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rk += __shfl_xor_sync(0xffffffffu, rk, o);
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  const double2* row = Yall + (long long)bidx * batchStride + (long long)r * ld;
  if (Yn) {
    double sc = 1.0;
    if (normalize & 1) sc = (v > zeroTol * vmax && v > 0) ? 1.0 / v : 0.0;
    double2* dst = Yn + ((long long)bidx * n + rk) * m;
    for (int k = lane; k < m; k += 32) {
      double2 u = row[k];
      u.x *= sc;
      u.y *= sc;
      dst[k] = u;
    }
  }
  if (Z) {
    double2* dst = Z + ((long long)bidx * n + rk) * mz;
    for (int k = lane; k < mz; k += 32) dst[k] = row[m + k];
  }
}

__global__ void __launch_bounds__(256) sort_norms_kernel(int n, double* __restrict__ s, int squared) {
  extern __shared__ double fsm[];
  const int bidx = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int q = threadIdx.x; q < n; q += blockDim.x) fsm[q] = s[(long long)bidx * n + q];
  __syncthreads();
  for (int r = warp; r < n; r += nw) {
    const double v = fsm[r];
    int rk = 0;
    for (int q = lane; q < n; q += 32) {
      const double w = fsm[q];
      rk += (w > v) || (w == v && q < r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
    if (lane == 0) s[(long long)bidx * n + rk] = squared ? v * v : v;
  }
}

// Y[b, i, 0..m) = L[b, i, 0..m),  Y[b, i, m + j] = (i == j): the matrix followed by the identity that will
// accumulate the row mixing.
__global__ void __launch_bounds__(256) rows_setup_kernel(int n, int m, const double2* __restrict__ L,
                                                         double2* __restrict__ Y) {
  const long long b = blockIdx.y;
  const int mt = m + n;
  const long long total = (long long)n * mt;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / mt), c = (int)(idx % mt);
    double2 v;
    if (c < m) {
      v = L[(b * n + i) * m + c];
    } else {
      v.x = (c - m == i) ? 1.0 : 0.0;
      v.y = 0.0;
    }
    Y[b * total + idx] = v;
  }
}


// Launch of the cluster-resident tournament. Returns -1 when the shape / device cannot run it (the caller takes the
// older paths), 0 on success, a CUDA error code otherwise.
static int jacobi_cluster_launch(const JacobiArgs& a0, int batch, double2* Y, cudaStream_t st) {
  constexpr int bc = 8;
  const int n = a0.n, mt = a0.mt;
  const int C = (((n + bc - 1) / bc) + 1) / 2;
  if (C < 2 || C > JC_MAXC || mt > 512) return -1;
  // one cluster per matrix, one CTA per SM (registers): batches beyond a single wave of clusters keep the older
  // kernels, whose throughput per matrix is better once the GPU is full (measured: 128 x n = 128 9.7 vs 12.9 ms)
  static const int maxCtas = getenv("MPDO_JACOBI_CLUSTER_CTAS") ? atoi(getenv("MPDO_JACOBI_CLUSTER_CTAS")) : 148;
  if ((long long)batch * C > maxCtas) return -1;
  const int E = mt <= 64 ? 2 : (mt <= 128 ? 4 : (mt <= 256 ? 8 : 16));
  const void* fn = E == 2 ? (const void*)jacobi_cluster_kernel<2>
                 : E == 4 ? (const void*)jacobi_cluster_kernel<4>
                 : E == 8 ? (const void*)jacobi_cluster_kernel<8> : (const void*)jacobi_cluster_kernel<16>;
  const size_t smem = (size_t)2 * bc * mt * sizeof(double2);
  static std::mutex mu;
  static int ok[5][JC_MAXC + 1];   // per (E class, cluster size): 0 unknown, 1 usable, -1 not
  const int ei = E == 2 ? 0 : (E == 4 ? 1 : (E == 8 ? 2 : 3));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, batch);
  cfg.blockDim = dim3(32 * bc);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (ok[ei][C] == 0) {
      ok[ei][C] = -1;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bc * 32 * E * (int)sizeof(double2)) ==
              cudaSuccess &&
          cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t qc = cfg;
        qc.gridDim = dim3(C, 1);
        qc.dynamicSmemBytes = (size_t)2 * bc * 32 * E * sizeof(double2);
        int nClusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nClusters, fn, &qc) == cudaSuccess && nClusters >= 1) ok[ei][C] = 1;
      }
      if (ok[ei][C] < 0) cudaGetLastError();
    }
    if (ok[ei][C] < 0) return -1;
  }
  JacobiArgs a = a0;
  a.b = bc;
  a.nbp = 2 * C;
  a.loop = 0;
  void* args[] = {(void*)&a, (void*)&Y};
  TimedLaunch timed(1, 4.0 * batch * (6.0 * a.m * (double)n * n + 20.0 * (double)n * n * n), 32.0 * batch * (double)n * a.m, st);
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "jacobi_cluster_kernel (cluster of %d): %s", C, cudaGetErrorString(e));
    return (int)e;
  }
  return check_launch("jacobi_cluster_kernel");
}
}  // namespace mpdo

namespace mpdo {
int jacobi_rows_ranked(int batch, int n, int m, int mt, int ld, long long batchStride, void* Y, double tol,
                       int maxSweeps, int32_t* work, const int* rank, int rankStride, void* stream) {
  if (batch <= 0 || n <= 0) return 0;
  if (!Y || !work || m <= 0 || mt < m || ld < mt) return fail(MPDO_EINVAL, "mpdo_jacobi_rows: bad argument");
  if (maxSweeps < 1) maxSweeps = 1;
  if (maxSweeps > 32) maxSweeps = 32;
  cudaStream_t st = (cudaStream_t)stream;
  MPDO_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t) * WORK_INTS * (size_t)batch, st));
  if (n == 1) return 0;
  static const bool trace = getenv("MPDO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[mpdo] jacobi n=%d m=%d mt=%d batch=%d\n", n, m, mt, batch);
  if (batch > 65535) return fail(MPDO_EINVAL, "mpdo_jacobi_rows: batch > 65535");
  row_norm_max_kernel<<<dim3((n + 7) / 8, batch), 256, 0, st>>>(n, m, ld, batchStride, (const double2*)Y, work);
  {
    int rc = check_launch("row_norm_max_kernel");
    if (rc) return rc;
  }

  // one-time set-up behind a mutex (strands call this from several host threads at once; the cached limit must not
  // be visible before the kernels' shared-memory attributes are set)
  static std::mutex initMu;
  static int smemMax = 0;
  {
    std::lock_guard<std::mutex> lk(initMu);
    if (!smemMax) {
      int dev = 0, sm = 0;
      MPDO_CUDA(cudaGetDevice(&dev));
      MPDO_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm - 1024));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm - 1024));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm - 1024));
      smemMax = sm;
    }
  }
  {
    // cluster-resident tournament for everything between "fits one CTA comfortably" and 256 rows (see
    // jacobi_cluster_kernel); MPDO_JACOBI_NOCLUSTER / MPDO_JACOBI_CLUSTER_MIN are A/B knobs
    static const bool noCluster = getenv("MPDO_JACOBI_NOCLUSTER") != nullptr;
    static const int clusterMin = getenv("MPDO_JACOBI_CLUSTER_MIN") ? atoi(getenv("MPDO_JACOBI_CLUSTER_MIN")) : 33;
    if (!noCluster && n >= clusterMin && n > 16 && n <= 16 * JC_MAXC && mt <= 512) {
      JacobiArgs a;
      a.n = n;
      a.m = m;
      a.mt = mt;
      a.ld = ld;
      a.batchStride = batchStride;
      a.round = 0;
      a.sweep = 0;
      a.maxSweeps = maxSweeps;
      a.tol = tol;
      a.cnt = work;
      a.rank = rank;
      a.rankStride = rankStride;
      const int rc = jacobi_cluster_launch(a, batch, (double2*)Y, st);
      if (rc >= 0) return rc;
    }
  }
  {
    // long rows, few matrices: block pairs in registers, sliced by columns (jacobi_persistent_cols_kernel).
    // MPDO_JACOBI_NOCOLS is the A/B knob.
    static const bool noCols = getenv("MPDO_JACOBI_NOCOLS") != nullptr;
    static std::mutex colsMu;
    static int colsOk = 0, colsSms = 0;   // 0 unknown, 1 usable, -1 not
    if (!noCols && mt > 256 && mt <= 2 * JR_THREADS) {
      // blocks of 4 rows when every block pair of every matrix gets its own SM (measured, one matrix: n = 512 12.8 ms
      // against 16.7 with blocks of 8 and 16.4 with the shared-memory kernel), else blocks of 8 (rows up to 512 entries)
      int br = 4;
      if ((long long)batch * ((((n + 3) / 4) + 1) / 2) > 148 && mt <= JR_THREADS) br = 8;
      if (const char* e = getenv("MPDO_JACOBI_COLS_BR")) br = atoi(e) == 8 && mt <= JR_THREADS ? 8 : 4;   // tuning knob
      const void* fn = br == 8 ? (const void*)jacobi_persistent_cols_kernel<8, 1>
                               : (mt <= JR_THREADS ? (const void*)jacobi_persistent_cols_kernel<4, 1>
                                                   : (const void*)jacobi_persistent_cols_kernel<4, 2>);
      const int nbr = (n + br - 1) / br;
      const int nbpr = (nbr + 1) & ~1;
      {
        std::lock_guard<std::mutex> lk(colsMu);
        if (colsOk == 0) {
          int dev = 0, coopAttr = 0;
          colsOk = -1;
          if (cudaGetDevice(&dev) == cudaSuccess &&
              cudaDeviceGetAttribute(&coopAttr, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coopAttr &&
              cudaDeviceGetAttribute(&colsSms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
            colsOk = 1;
          else
            cudaGetLastError();
        }
      }
      int perSm = 0;
      if (colsOk == 1 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, JR_THREADS, 0) != cudaSuccess) {
        cudaGetLastError();
        perSm = 0;
      }
      const long long capacity = (long long)perSm * colsSms;
      // (large batches keep the older policy: few CTAs per matrix, all matrices at once)
      if (colsOk == 1 && n > 2 * br && capacity >= nbpr / 2 && (long long)batch * (nbpr / 2) <= capacity) {
        int chunk = (int)(capacity / (nbpr / 2));
        if (chunk > batch) chunk = batch;
        JacobiArgs a;
        a.n = n;
        a.m = m;
        a.mt = mt;
        a.ld = ld;
        a.batchStride = batchStride;
        a.b = br;
        a.nbp = nbpr;
        a.round = 0;
        a.sweep = 0;
        a.maxSweeps = maxSweeps;
        a.loop = 0;
        a.tol = tol;
        a.rankStride = rankStride;
        bool ok = true;
        for (int b0 = 0; b0 < batch && ok; b0 += chunk) {
          const int nbat = batch - b0 < chunk ? batch - b0 : chunk;
          a.cnt = work + (long long)b0 * WORK_INTS;
          a.rank = rank ? rank + (long long)b0 * rankStride : nullptr;
          double2* Yc = (double2*)Y + (long long)b0 * batchStride;
          void* args[] = {(void*)&a, (void*)&Yc};
          TimedLaunch timed(1, 4.0 * nbat * (6.0 * m * (double)n * n + 20.0 * (double)n * n * n),
                            32.0 * nbat * (double)n * m, st);
          cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)(nbpr / 2), nbat), dim3(JR_THREADS), args, 0, st);
          if (e == cudaSuccess) {
            ++g_launches;
          } else {
            cudaGetLastError();
            ok = false;
            if (b0 > 0) {
              snprintf(g_err, sizeof(g_err), "jacobi_persistent_cols (chunk %d): %s", b0, cudaGetErrorString(e));
              return (int)e;
            }
          }
        }
        if (ok) return 0;   // first launch refused: the older kernels below
      }
    }
  }
  const long long rowBytes = (long long)mt * sizeof(double2);
  int bmax = (int)((smemMax - 1024) / (2 * rowBytes));
  if (bmax < 1) return fail(MPDO_ENOSMEM, "mpdo_jacobi_rows: a row pair does not fit shared memory");
  int b = bmax < 32 ? bmax : 32;
  const int half = (n + 1) / 2;
  if (b > half) b = half;
  if (b < half) {
    // The matrix does not fit one CTA. A CTA's time per round grows like b^2 (b steps, each moving 2b rows
    // through shared memory) while the number of CTAs per round is n/(2b): with few matrices in flight, small
    // blocks spread one decomposition over many SMs (measured: b = 32 uses 2 SMs at 150 us per round for n = 128).
    // A full batch already fills the GPU and prefers fewer, longer rounds.
    const long long ctasAt8 = (long long)batch * ((n + 15) / 16);
    // Few matrices in flight: 8-row blocks with a full warp per pair (measured on graded full-rank Gram matrices,
    // preconditioned route: n = 256 6.67 -> 5.19 ms, n = 512 19.2 -> 16.0 ms against 16-row blocks).
    int target = ctasAt8 <= 8 * 148 ? 8 : 32;
    if (const char* e = getenv("MPDO_JACOBI_B")) target = atoi(e) > 0 ? atoi(e) : target;   // tuning knob
    if (b > target) b = target;
  }
  const int nb = (n + b - 1) / b;
  const int nbp = (nb + 1) & ~1;

  JacobiArgs a;
  a.n = n;
  a.m = m;
  a.mt = mt;
  a.ld = ld;
  a.batchStride = batchStride;
  a.b = b;
  a.nbp = nbp;
  a.round = 0;
  a.sweep = 0;
  a.maxSweeps = maxSweeps;
  a.tol = tol;
  a.cnt = work;
  a.rank = rank;
  a.rankStride = rankStride;
  const size_t smem = (size_t)(2 * b) * rowBytes;
  // lanes per row pair: short rows share a warp between several pairs (see rotate_pair)
  // (8 / 16 / 32 measured on single-CTA n = 24, 48, 96: 16 is best or equal up to mt = 192; multi-CTA tournaments
  // with 8-row blocks need the full warp per pair to keep 8 warps on the SM)
  int G = (mt >= 384 || (b < half && b <= 8)) ? 32 : 16;
  if (const char* e = getenv("MPDO_JACOBI_G")) G = atoi(e) == 8 ? 8 : (atoi(e) == 16 ? 16 : 32);   // tuning knob
  const int ppw = 32 / G;
  const unsigned threads = 32u * (unsigned)((b + ppw - 1) / ppw);
  // algorithmic work of the decomposition this launch is (part of): the SURVEY 8d SVD count 4 (6 m n^2 + 20 n^3)
  // for n rows of length m per matrix (an eigen-decomposition has m = n); multi-launch sweeps report it once
  auto launch = [&](dim3 grid) {
    TimedLaunch timed(1, 4.0 * batch * (6.0 * m * (double)n * n + 20.0 * (double)n * n * n), 32.0 * batch * (double)n * m, st);
    if (G == 8)
      jacobi_kernel<8><<<grid, threads, smem, st>>>(a, (double2*)Y);
    else if (G == 16)
      jacobi_kernel<16><<<grid, threads, smem, st>>>(a, (double2*)Y);
    else
      jacobi_kernel<32><<<grid, threads, smem, st>>>(a, (double2*)Y);
  };
  if (nbp == 2) {
    a.loop = 1;
    launch(dim3(1, batch));
    return check_launch("jacobi_kernel(loop)");
  }
  a.loop = 0;
  // Persistent path: every CTA of the grid must be resident at once (device-wide barrier inside the kernel).
  {
    static std::mutex coopMu;
    static int coop = -1, sms = 0;
    std::unique_lock<std::mutex> coopLock(coopMu);
    if (coop < 0) {
      int dev = 0, coopAttr = 0;
      MPDO_CUDA(cudaGetDevice(&dev));
      MPDO_CUDA(cudaDeviceGetAttribute(&coopAttr, cudaDevAttrCooperativeLaunch, dev));
      MPDO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_persistent_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smemMax - 1024));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_persistent_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smemMax - 1024));
      MPDO_CUDA(cudaFuncSetAttribute(jacobi_persistent_reg_kernel<16, 512>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax - 1024));
      coop = getenv("MPDO_JACOBI_MULTILAUNCH") ? 0 : coopAttr;   // (debugging knob) published last
    }
    coopLock.unlock();
    // CTAs per matrix: all nbp/2 block pairs of a round in parallel when the GPU has room, fewer (each CTA then
    // walks several pairs) when the batch is large, down to one CTA per matrix, which needs no device-wide barrier
    // and therefore no cooperative launch (any batch size).
    const void* fn = G == 32 ? (const void*)jacobi_persistent_kernel<32> : (const void*)jacobi_persistent_kernel<16>;
    if (G == 8) fn = nullptr;   // (tuning knob only: the 8-lane variant exists for the single-CTA kernel)
    int perSm = 0;
    if (fn && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, (int)threads, smem) != cudaSuccess) {
      cudaGetLastError();
      perSm = 0;
    }
    const long long capacity = (long long)perSm * sms;
    long long P = batch > 0 ? capacity / batch : 0;
    if (P > nbp / 2) P = nbp / 2;
    // Measured on a 128-matrix batch of n = 512: one CTA walking all pairs of its matrix (P = 1, no barrier) takes
    // 371 ms where the multi-launch driver below takes 206 ms, so sharing is opt-in (MPDO_JACOBI_SHARE) and the
    // persistent kernel is used when every block pair of a round gets its own CTA.
    static const bool share = getenv("MPDO_JACOBI_SHARE") != nullptr;
    // A batch too large to be co-resident with one CTA per block pair is walked in consecutive cooperative launches
    // of `chunk` matrices each (every launch still runs all sweeps of its matrices on the device, no host polling).
    static const bool noChunk = getenv("MPDO_JACOBI_NOCHUNK") != nullptr;   // A/B knob: multi-launch driver instead
    int chunk = batch;
    if (P < nbp / 2 && !share && !noChunk && capacity >= nbp / 2) {
      chunk = (int)(capacity / (nbp / 2));
      P = nbp / 2;
    }
    if (fn && perSm > 0 && coop && (P == nbp / 2 || (share && P >= 1))) {
      unsigned nthreads = threads;
      static const bool noReg = getenv("MPDO_JACOBI_NOREG") != nullptr;   // debugging knob
      if (!noReg && P == nbp / 2 && mt > 256 && mt <= 512 && b <= 16) {
        // rows of 257..512 entries: one warp per row, rows of block I register-resident (see rotate_reg). Measured
        // on B200: 18% faster at mt = 512; no gain at mt = 256 (8 entries per lane) and 40% slower at mt = 1024
        // (32 entries per lane leave 7 warps per SM), so those keep the shared-memory kernel.
        fn = (const void*)jacobi_persistent_reg_kernel<16, 512>;
        nthreads = 32u * (unsigned)b;
      }
      bool ok = true;
      for (int b0 = 0; b0 < batch && ok; b0 += chunk) {
        const int nbat = batch - b0 < chunk ? batch - b0 : chunk;
        JacobiArgs ac = a;
        ac.cnt = work + (long long)b0 * WORK_INTS;
        ac.rank = rank ? rank + (long long)b0 * rankStride : nullptr;
        double2* Yc = (double2*)Y + (long long)b0 * batchStride;
        void* args[] = {(void*)&ac, (void*)&Yc};
        TimedLaunch timed(1, 4.0 * nbat * (6.0 * m * (double)n * n + 20.0 * (double)n * n * n), 32.0 * nbat * (double)n * m, st);
        cudaError_t e;
        if (P > 1) {
          e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)P, nbat), dim3(nthreads), args, smem, st);
        } else {
          e = cudaLaunchKernel(fn, dim3(1, nbat), dim3(nthreads), args, smem, st);
        }
        if (e == cudaSuccess) {
          ++g_launches;
        } else {
          cudaGetLastError();   // e.g. cudaErrorCooperativeLaunchTooLarge
          ok = false;
          if (b0 > 0) {   // earlier chunks already ran: the multi-launch driver below must not redo them
            snprintf(g_err, sizeof(g_err), "jacobi_persistent (chunk %d): %s", b0, cudaGetErrorString(e));
            return (int)e;
          }
        }
      }
      if (ok) return 0;   // first launch refused: fall through to the multi-launch driver
    }
  }
  for (int sw = 0; sw < maxSweeps; ++sw) {
    a.sweep = sw;
    for (int r = 0; r < nbp - 1; ++r) {
      a.round = r;
      launch(dim3(nbp / 2, batch));
      int rc = check_launch("jacobi_kernel");
      if (rc) return rc;
    }
    // SYNC (multi-launch mode only): read this sweep's rotation counters back and stop once every matrix of
    // the batch has converged, instead of enqueueing no-op rounds up to maxSweeps.
    std::vector<int> hCntV((size_t)batch);   // pageable on purpose (no cudaMallocHost on a strand thread)
    int* hCnt = hCntV.data();
    MPDO_CUDA(cudaMemcpy2DAsync(hCnt, sizeof(int), work + sw, sizeof(int) * WORK_INTS, sizeof(int), batch,
                                cudaMemcpyDeviceToHost, st));
    MPDO_CUDA(cudaStreamSynchronize(st));
    bool any = false;
    for (int i = 0; i < batch; ++i) any |= (hCnt[i] != 0);
    if (!any) break;
  }
  return 0;
}
}  // namespace mpdo

extern "C" int mpdo_jacobi_rows(int batch, int n, int m, int mt, int ld, int64_t batchStride, void* Y, double tol,
                                int maxSweeps, int32_t* work, void* stream) {
  return mpdo::jacobi_rows_ranked(batch, n, m, mt, ld, batchStride, Y, tol, maxSweeps, work, nullptr, 0, stream);
}

extern "C" int mpdo_rows_finalize(int batch, int n, int m, int mz, int ld, int64_t batchStride, const void* Y,
                                  double* s, void* Yn, void* Z, int normalize, double zeroTol, void* stream) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 0;
  if (!Y || !s || m <= 0) return fail(MPDO_EINVAL, "mpdo_rows_finalize: bad argument");
  const size_t smem = (size_t)n * (sizeof(double) + sizeof(int));
  if (smem > 48 * 1024) return fail(MPDO_ENOSMEM, "mpdo_rows_finalize: n too large");
  if (n >= 192 && batch <= 4096) {
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((n + 7) / 8, batch);
    row_norms_kernel<<<grid, 256, 0, st>>>(n, m, ld, batchStride, (const double2*)Y, s);
    int rc = check_launch("row_norms_kernel");
    if (rc) return rc;
    if (Yn || Z) {
      rows_scatter_kernel<<<grid, 256, sizeof(double) * (size_t)n, st>>>(n, m, mz, ld, batchStride, (const double2*)Y, s,
                                                                        (double2*)Yn, (double2*)Z, normalize, zeroTol);
      rc = check_launch("rows_scatter_kernel");
      if (rc) return rc;
    }
    sort_norms_kernel<<<batch, 256, sizeof(double) * (size_t)n, st>>>(n, s, (normalize & 2) ? 1 : 0);
    return check_launch("sort_norms_kernel");
  }
  rows_finalize_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(n, m, mz, ld, batchStride, (const double2*)Y, s,
                                                                  (double2*)Yn, (double2*)Z, normalize, zeroTol);
  return check_launch("rows_finalize_kernel");
}

extern "C" int mpdo_decompose_rows(int batch, int n, int m, const void* L, void* Y, int32_t* work, double* s,
                                   void* Yn, void* Z, int normalize, double zeroTol, double tol, int maxSweeps,
                                   void* stream) {
  using namespace mpdo;
  if (batch <= 0 || n <= 0) return 0;
  if (!L || !Y || !work || !s || m <= 0) return fail(MPDO_EINVAL, "mpdo_decompose_rows: bad argument");
  if (batch > 65535) return fail(MPDO_EINVAL, "mpdo_decompose_rows: batch > 65535");
  const int mt = m + n;
  const long long total = (long long)n * mt;
  unsigned gx = (unsigned)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
  rows_setup_kernel<<<dim3(gx, batch), 256, 0, (cudaStream_t)stream>>>(n, m, (const double2*)L, (double2*)Y);
  int rc = check_launch("rows_setup_kernel");
  if (rc) return rc;
  rc = mpdo_jacobi_rows(batch, n, m, mt, mt, total, Y, tol, maxSweeps, work, stream);
  if (rc) return rc;
  return mpdo_rows_finalize(batch, n, m, n, mt, total, Y, s, Yn, Z, normalize, zeroTol, stream);
}
