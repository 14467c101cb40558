// Batched complex tensor contraction for MPDO site tensors (sm_100a).
//
//   C[b,i,j] = alpha * sum_k opA(A[b,i,k]) * opB(B[b,k,j]) + beta * C[b,i,j]
//
// Every logical axis is a composite index (mpdo_idxmap) over up to three axes of the operand, so the
// same kernel performs the two-site bond merge, the R-absorb of the QR sweep, Gram matrices over any
// grouping of T[l,s,a,r], the projected updates U^h*Theta of the bond/kappa truncation and the
// transfer-matrix steps of the readout without a transpose pass through HBM.
//
// Tiling: 64x64 output tile per CTA, K staged through shared memory in slabs of 16, 256 threads each
// owning a 4x4 interleaved micro-tile of complex accumulators (rows ty+16u, columns tx+16v, so both the
// shared-memory reads and the global stores of a warp are stride-1). Global loads are issued one tile
// ahead into registers; the loader picks the thread->element mapping that makes the contiguous axis of
// each operand the fastest-varying one. complex64 operands may be accumulated in fp32 (FFMA) or fp64
// (DFMA; B200 runs fp64 at half the fp32 rate, which is what makes fp64 Gram matrices affordable).
//
// fp64-accumulated contractions (Gram matrices, fp64 cores, every complex128 contraction) run on the fp64 tensor
// pipe: the same 64x64x16 shared-memory tiles are consumed by 8 warps (4 along M x 2 along N, a 16x32 complex
// warp tile) with mma.sync.m8n8k4.f64 (DMMA) - four real DMMAs per complex 8x8x4 block, the imaginary part of the A
// fragment negated once in registers. One DMMA retires 256 FMAs per warp instruction where DFMA retires 32, so the
// kernel spends its issue slots on the fp64 pipe instead of on shared-memory operand traffic. B200 retires DMMA on
// the same fp64 units as DFMA, so the gain is small (see the dispatch in mpdo_contract for the measured numbers).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace mpdo {

constexpr int BM = 64, BN = 64, BK = 16;
// MMA = 0: scalar tiles, 256 threads; 1: DMMA, 8 warps with 16x32 warp tiles; 2: DMMA, 16 warps with 16x16 warp tiles
__host__ __device__ constexpr int threads_of(int mma) { return mma == 2 ? 512 : 256; }
constexpr int TM = BM / 16, TN = BN / 16;

// D(8x8) += A(8x4, row) * B(4x8, col), fp64. Fragments: A: lane holds A[lane/4][lane%4]; B: lane holds B[lane%4][lane/4];
// C/D: lane holds C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <typename TA, typename TB, typename TC, typename R, int MMA>
__global__ void __launch_bounds__(threads_of(MMA), MMA ? 1 : 2) contract_kernel(const mpdo_contract_desc d, const TA* __restrict__ A,
                                                      const TB* __restrict__ B, TC* __restrict__ C, int tilesM,
                                                      int tilesN, int kChunk) {
  using CR = typename cplx<R>::type;
  // row stride in 16-byte units: 65 = 1 (mod 8) spreads the scalar kernel's column reads; 66 = 2 (mod 8) makes the
  // DMMA fragment reads (4 k-rows x 2 columns per quarter warp) hit eight distinct 16-byte bank groups
  constexpr int PAD = MMA ? 2 : 1;
  __shared__ CR As[BK][BM + PAD];
  __shared__ CR Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;

  long long t = blockIdx.x;
  int tm, tn;
  if (d.hermitian) {   // tiles on and below the diagonal, enumerated row by row: index = tm (tm + 1) / 2 + tn
    const int tri = tilesM * (tilesM + 1) / 2;
    const int ti = (int)(t % tri);
    t /= tri;
    tm = (int)((sqrtf(8.0f * (float)ti + 1.0f) - 1.0f) * 0.5f);
    while (tm * (tm + 1) / 2 > ti) --tm;
    while ((tm + 1) * (tm + 2) / 2 <= ti) ++tm;
    tn = ti - tm * (tm + 1) / 2;
  } else {
    tn = (int)(t % tilesN);
    t /= tilesN;
    tm = (int)(t % tilesM);
    t /= tilesM;
  }
  const int ksplit = d.ksplit > 1 ? d.ksplit : 1;
  const int ks = (int)(t % ksplit);
  const int b = (int)(t / ksplit);

  const int i0 = tm * BM, j0 = tn * BN;
  const int kBeg = ks * kChunk;
  const int kEnd = min(d.K, kBeg + kChunk);

  const TA* Ab = A + map_idx(d.Ab, b);
  const TB* Bb = B + map_idx(d.Bb, b);

  // ---- loader geometry -----------------------------------------------------------------------
  constexpr int NT = threads_of(MMA);
  constexpr int LA = BM * BK / NT;  // elements of the A tile per thread
  constexpr int LB = BN * BK / NT;
  const bool akf = d.a_kfast != 0;
  const bool bjf = d.b_jfast != 0;
  // A, k fastest: kk fixed per thread, rows ii0 + r*(NT/BK).  A, i fastest: ii fixed, k = kk0 + r*(NT/BM).
  const int a_kk = akf ? tid % BK : tid / BM;
  const int a_ii = akf ? tid / BK : tid % BM;
  const int b_kk = bjf ? tid / BN : tid % BK;
  const int b_jj = bjf ? tid % BN : tid / BK;

  long long offAi[LA];
  long long offBj[LB];
#pragma unroll
  for (int r = 0; r < LA; ++r) {
    int i = i0 + (akf ? a_ii + r * (NT / BK) : a_ii);
    offAi[r] = (i < d.M) ? map_idx(d.Ai, i) : -1;
  }
#pragma unroll
  for (int r = 0; r < LB; ++r) {
    int j = j0 + (bjf ? b_jj : b_jj + r * (NT / BK));
    offBj[r] = (j < d.N) ? map_idx(d.Bj, j) : -1;
  }

  TA ra[LA];   // prefetched in the operand's own type (half the registers for fp32 operands under fp64 accumulate)
  TB rb[LB];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int r = 0; r < LA; ++r) {
      int k = k0 + (akf ? a_kk : a_kk + r * (NT / BM));
      TA v;
      v.x = 0;
      v.y = 0;
      if (k < kEnd && offAi[r] >= 0) {
        v = Ab[offAi[r] + map_idx(d.Ak, k)];
      }
      ra[r] = v;
    }
#pragma unroll
    for (int r = 0; r < LB; ++r) {
      int k = k0 + (bjf ? b_kk + r * (NT / BN) : b_kk);
      TB v;
      v.x = 0;
      v.y = 0;
      if (k < kEnd && offBj[r] >= 0) {
        v = Bb[offBj[r] + map_idx(d.Bk, k)];
      }
      rb[r] = v;
    }
  };
  // Conjugation happens here, not in load_tile: anything that consumes a loaded value there makes the warp wait
  // for the load at once (ncu: a third of all stall samples sat on the sign flip right behind the LDG) instead of
  // after the slab's arithmetic.
  const R sgnA = d.conjA ? (R)-1 : (R)1, sgnB = d.conjB ? (R)-1 : (R)1;
  auto store_tile = [&]() {
#pragma unroll
    for (int r = 0; r < LA; ++r) {
      CR v = cconv<CR>(ra[r]);
      v.y *= sgnA;
      if (akf)
        As[a_kk][a_ii + r * (NT / BK)] = v;
      else
        As[a_kk + r * (NT / BM)][a_ii] = v;
    }
#pragma unroll
    for (int r = 0; r < LB; ++r) {
      CR v = cconv<CR>(rb[r]);
      v.y *= sgnB;
      if (bjf)
        Bs[b_kk + r * (NT / BN)][b_jj] = v;
      else
        Bs[b_kk][b_jj + r * (NT / BK)] = v;
    }
  };

  // ---- epilogue helper: one output element (alpha/beta, split-K atomics, Hermitian mirror) ------------
  using RC = typename real_of<TC>::type;
  TC* Cb = C + map_idx(d.Cb, b);
  const R alpha = (R)d.alpha;
  const R beta = (R)d.beta;
  auto emit = [&](int i, int j, long long oi, long long oj, R accRe, R accIm) {
    TC* p = Cb + oi + oj;
    R re = alpha * accRe, im = alpha * accIm;
    if (ksplit > 1) {
      atomicAdd(&p->x, (RC)re);
      atomicAdd(&p->y, (RC)im);
    } else {
      if (beta != (R)0) {
        TC old = *p;
        re += beta * (R)old.x;
        im += beta * (R)old.y;
      }
      TC o;
      o.x = (RC)re;
      o.y = (RC)im;
      *p = o;
    }
    if (d.hermitian && tm != tn) {   // mirror: C[j,i] = conj(C[i,j])
      TC* q = Cb + map_idx(d.Ci, j) + map_idx(d.Cj, i);
      if (ksplit > 1) {
        atomicAdd(&q->x, (RC)re);
        atomicAdd(&q->y, (RC)(-im));
      } else {
        TC o;
        o.x = (RC)re;
        o.y = (RC)(-im);
        *q = o;
      }
    }
  };

  if constexpr (MMA) {
    // ---- fp64 tensor pipe: warp tile 16 (M) x 8 NB (N) complex = 2 x NB DMMA blocks, re and im accumulators ------
    constexpr int NB = MMA == 2 ? 2 : 4;   // 16 warps: 4 x 4 warps of 16 x 16; 8 warps: 4 x 2 warps of 16 x 32
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int fr = lane >> 2, fk = lane & 3;
    double accRe[2][NB][2], accIm[2][NB][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < NB; ++v) {
        accRe[u][v][0] = accRe[u][v][1] = 0.0;
        accIm[u][v][0] = accIm[u][v][1] = 0.0;
      }
    if (kBeg < kEnd) {
      load_tile(kBeg);
      store_tile();
      __syncthreads();
      for (int k0 = kBeg; k0 < kEnd; k0 += BK) {
        const bool more = (k0 + BK) < kEnd;
        if (more) load_tile(k0 + BK);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
          CR a[2], bb[NB];
#pragma unroll
          for (int u = 0; u < 2; ++u) a[u] = As[k4 + fk][wm * 16 + u * 8 + fr];
#pragma unroll
          for (int v = 0; v < NB; ++v) bb[v] = Bs[k4 + fk][wn * (8 * NB) + v * 8 + fr];
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int v = 0; v < NB; ++v) {
              dmma(accRe[u][v], a[u].x, bb[v].x);
              dmma(accIm[u][v], a[u].x, bb[v].y);
            }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const double nai = -a[u].y;
#pragma unroll
            for (int v = 0; v < NB; ++v) {
              dmma(accRe[u][v], nai, bb[v].y);
              dmma(accIm[u][v], a[u].y, bb[v].x);
            }
          }
        }
        __syncthreads();
        if (more) {
          store_tile();
          __syncthreads();
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + wm * 16 + u * 8 + fr;
      if (i >= d.M) continue;
      const long long oi = map_idx(d.Ci, i);
#pragma unroll
      for (int v = 0; v < NB; ++v)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = j0 + wn * (8 * NB) + v * 8 + 2 * fk + e;
          if (j >= d.N) continue;
          emit(i, j, oi, map_idx(d.Cj, j), (R)accRe[u][v][e], (R)accIm[u][v][e]);
        }
    }
  } else {
    CR acc[TM][TN];
#pragma unroll
    for (int u = 0; u < TM; ++u)
#pragma unroll
      for (int v = 0; v < TN; ++v) {
        acc[u][v].x = 0;
        acc[u][v].y = 0;
      }

    if (kBeg < kEnd) {
      load_tile(kBeg);
      store_tile();
      __syncthreads();
      for (int k0 = kBeg; k0 < kEnd; k0 += BK) {
        const bool more = (k0 + BK) < kEnd;
        if (more) load_tile(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          CR a[TM], bb[TN];
#pragma unroll
          for (int u = 0; u < TM; ++u) a[u] = As[kk][ty + 16 * u];
#pragma unroll
          for (int v = 0; v < TN; ++v) bb[v] = Bs[kk][tx + 16 * v];
#pragma unroll
          for (int u = 0; u < TM; ++u)
#pragma unroll
            for (int v = 0; v < TN; ++v) cfma(acc[u][v], a[u], bb[v]);
        }
        __syncthreads();
        if (more) {
          store_tile();
          __syncthreads();
        }
      }
    }

    long long offCj[TN];
#pragma unroll
    for (int v = 0; v < TN; ++v) {
      int j = j0 + tx + 16 * v;
      offCj[v] = (j < d.N) ? map_idx(d.Cj, j) : -1;
    }
#pragma unroll
    for (int u = 0; u < TM; ++u) {
      int i = i0 + ty + 16 * u;
      if (i >= d.M) continue;
      long long oi = map_idx(d.Ci, i);
#pragma unroll
      for (int v = 0; v < TN; ++v) {
        if (offCj[v] < 0) continue;
        emit(i, j0 + tx + 16 * v, oi, offCj[v], acc[u][v].x, acc[u][v].y);
      }
    }
  }
}

template <typename TA, typename TB, typename TC, typename R, int MMA = 0>
static int launch_contract(const mpdo_contract_desc& d, const void* A, const void* B, void* C, cudaStream_t st) {
  const int tilesM = (d.M + BM - 1) / BM, tilesN = (d.N + BN - 1) / BN;
  const int ksplit = d.ksplit > 1 ? d.ksplit : 1;
  int kChunk = (d.K + ksplit - 1) / ksplit;
  kChunk = ((kChunk + BK - 1) / BK) * BK;
  const long long tilesMN = d.hermitian ? (long long)tilesM * (tilesM + 1) / 2 : (long long)tilesM * tilesN;
  long long grid = tilesMN * ksplit * d.batch;
  if (grid <= 0) return 0;
  if (grid > 2147483647LL) return fail(MPDO_EINVAL, "mpdo_contract: grid too large");
  {
    const double mnk = (double)d.M * d.N * d.K * d.batch;
    const double byts = (double)d.batch * ((double)d.M * d.K * sizeof(TA) + (double)d.K * d.N * sizeof(TB) +
                                           (double)d.M * d.N * sizeof(TC));
    // algorithmic cost of a complex contraction; a Hermitian result needs its M (M + 1) / 2 entries on
    // and below the diagonal only
    const double flops = d.hermitian ? 4.0 * mnk * (d.M + 1.0) / d.M : 8.0 * mnk;
    TimedLaunch timed(std::is_same<R, double>::value ? 3 : 0, flops, byts, st);   // class 3: fp64 accumulation
    contract_kernel<TA, TB, TC, R, MMA><<<(unsigned)grid, threads_of(MMA), 0, st>>>(d, (const TA*)A, (const TB*)B, (TC*)C, tilesM,
                                                                    tilesN, kChunk);
  }
  return check_launch("contract_kernel");
}

}  // namespace mpdo

extern "C" int mpdo_contract(const mpdo_contract_desc* dp, const void* A, const void* B, void* C, void* stream) {
  using namespace mpdo;
  if (!dp || !A || !B || !C) return fail(MPDO_EINVAL, "mpdo_contract: null argument");
  const mpdo_contract_desc& d = *dp;
  if (d.M < 0 || d.N < 0 || d.K < 0 || d.batch < 0) return fail(MPDO_EINVAL, "mpdo_contract: negative extent");
  if (d.M == 0 || d.N == 0 || d.batch == 0) return 0;
  if (d.hermitian && (d.M != d.N || d.beta != 0.0))
    return fail(MPDO_EINVAL, "mpdo_contract: hermitian needs M == N and beta == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int key = (d.dtypeA << 2) | (d.dtypeB << 1) | d.dtypeC;
  const bool f64 = d.acc64 || key != 0;
  if (!f64) {
    // tall complex64 applies with a small second operand go to the tcgen05 / TMA tile (tc_apply.cu); everything
    // else (composite-index views, short products) stays on the FFMA tiles
    const int rc = tc::try_apply(d, A, B, C, st);
    if (rc == 1) return 0;
    if (rc != 0) return rc;
    return launch_contract<float2, float2, float2, float>(d, A, B, C, st);
  }
  // Measured on B200 (tools/bench_contract.py, profiles/r1_contract_ncu.md): DMMA tiles 18.1 TFLOP/s on the kappa
  // Gram matrix (1024 x 1024 x 8192, complex64 in, fp64 accumulate) and 26.5 TFLOP/s on a complex128-operand apply
  // (512 x 8192 x 512), scalar DFMA tiles 16.5 and 22.0. B200 retires DMMA at about the DFMA rate, so the tensor form
  // only frees issue slots and registers. MPDO_NO_DMMA=1 forces the scalar kernel for A/B comparisons.
  static const bool noDmma = getenv("MPDO_NO_DMMA") != nullptr;
  static const bool dmma8 = getenv("MPDO_DMMA_8WARPS") != nullptr;   // A/B knob: 8-warp DMMA tiles
  if (!noDmma && dmma8) {
    switch (key) {
      case 0: return launch_contract<float2, float2, float2, double, 1>(d, A, B, C, st);
      case 1: return launch_contract<float2, float2, double2, double, 1>(d, A, B, C, st);
      case 2: return launch_contract<float2, double2, float2, double, 1>(d, A, B, C, st);
      case 3: return launch_contract<float2, double2, double2, double, 1>(d, A, B, C, st);
      case 4: return launch_contract<double2, float2, float2, double, 1>(d, A, B, C, st);
      case 5: return launch_contract<double2, float2, double2, double, 1>(d, A, B, C, st);
      case 6: return launch_contract<double2, double2, float2, double, 1>(d, A, B, C, st);
      case 7: return launch_contract<double2, double2, double2, double, 1>(d, A, B, C, st);
    }
  }
  if (!noDmma) {
    switch (key) {
      case 0: return launch_contract<float2, float2, float2, double, 2>(d, A, B, C, st);
      case 1: return launch_contract<float2, float2, double2, double, 2>(d, A, B, C, st);
      case 2: return launch_contract<float2, double2, float2, double, 2>(d, A, B, C, st);
      case 3: return launch_contract<float2, double2, double2, double, 2>(d, A, B, C, st);
      case 4: return launch_contract<double2, float2, float2, double, 2>(d, A, B, C, st);
      case 5: return launch_contract<double2, float2, double2, double, 2>(d, A, B, C, st);
      case 6: return launch_contract<double2, double2, float2, double, 2>(d, A, B, C, st);
      case 7: return launch_contract<double2, double2, double2, double, 2>(d, A, B, C, st);
    }
  }
  switch (key) {
    case 0: return launch_contract<float2, float2, float2, double>(d, A, B, C, st);
    case 1: return launch_contract<float2, float2, double2, double>(d, A, B, C, st);
    case 2: return launch_contract<float2, double2, float2, double>(d, A, B, C, st);
    case 3: return launch_contract<float2, double2, double2, double>(d, A, B, C, st);
    case 4: return launch_contract<double2, float2, float2, double>(d, A, B, C, st);
    case 5: return launch_contract<double2, float2, double2, double>(d, A, B, C, st);
    case 6: return launch_contract<double2, double2, float2, double>(d, A, B, C, st);
    case 7: return launch_contract<double2, double2, double2, double>(d, A, B, C, st);
  }
  return fail(MPDO_EINVAL, "mpdo_contract: bad dtype");
}
