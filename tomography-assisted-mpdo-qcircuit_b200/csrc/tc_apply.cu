// Blackwell tensor-core path of the complex64 "apply" contractions (sm_100a: TMA + tcgen05.mma + TMEM).
//
//   C[b, M, N] = alpha * op(A)[b, M, K] . op(B)[b, K, N]        complex64 in / out, A tall (M >> K, N), B small
//
// These are the products that apply a small core to a site tensor: Q = T . Linv^h in the QR sweep
// (reference TNNOptimizer.py:98-106), T_l . U sqrt(S) in the chi sweep (:126-133), A_lo . W in the gate split
// (Circuit.py:104-124). At chi >= 128 they are dense GEMMs with tens of GFLOP each.
//
// Formulation. Interleaved complex storage makes a complex matrix A[M, K] a REAL matrix At[M, 2K] as it lies in
// memory, and the complex product a real GEMM against an expanded small operand:
//   Ct[M, 2N] = At[M, 2K] . Bt[2K, 2N],   Bt[(k,0),(n,0)] = Br  Bt[(k,1),(n,0)] = -Bi  Bt[(k,0),(n,1)] = Bi  Bt[(k,1),(n,1)] = Br
// (signs of the (k,1) rows flip for conj(A)). At is consumed as is - no de-interleave pass over the big operand - and
// Ct is the interleaved complex result. Same 4 real multiplications per complex one as the SIMT kernel.
//
// Precision. tcgen05 kind::tf32 reads fp32 from shared memory and keeps 10 mantissa bits. Each operand is split
// x = hi + lo (hi = x rounded to tf32, lo = x - hi rounded to tf32) and the product accumulated in fp32 in TMEM as
// hi.hi + lo.hi + hi.lo (3xTF32): the dropped lo.lo term is <= 2^-22 relative per product. The small operand is split
// once by the preparation kernel; the big operand is split per tile inside the kernel by four converter warps
// (elementwise, so the swizzled TMA layout is preserved: hi overwrites the raw tile, lo goes to a sibling tile).
//
// Kernel anatomy (one CTA per 128-row tile of M x BN real columns, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor loads of the raw A tile [128 x 32 fp32] and the hi / lo tiles of Bt^T
//              [BN x 32], 128-byte swizzle, completing on full[s]
//   warps 2-5  converter: wait full[s], split the A tile, fence.proxy.async, arrive on conv[s]; after the main loop
//              the same warps are the epilogue: tcgen05.ld the fp32 accumulators (one TMEM lane = one output row) and
//              store the interleaved complex rows
//   warp 1     TMEM allocation and the MMA issuer: one elected lane issues 4 k-steps x 3 tcgen05.mma (M=128, N=BN,
//              K=8) per stage and tcgen05.commit-s the stage back to the producer (empty[s]) / the accumulator to the
//              epilogue (accum)
// Every mbarrier wait is bounded: a protocol error traps instead of hanging the device.
#include <cuda.h>

#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "engine.cuh"

namespace mpdo {
namespace tc {

constexpr int BM = 128;        // rows per CTA = TMEM lanes
constexpr int BKR = 32;        // real k per stage = one 128-byte swizzle row of fp32
constexpr int UMMA_K = 8;      // tf32
constexpr int A_TILE = BM * BKR * 4;   // 16 KB

static int initial_mode() {
  const char* e = getenv("MPDO_TC");      // 0 = off, 1 = on (default), 2 = on with the widest tiles
  if (!e) return 1;
  const int v = atoi(e);
  return v < 0 ? 0 : (v > 2 ? 2 : v);
}
static std::atomic<int> g_enabled{initial_mode()};

// Scratch for the prepared small operand: one grow-only device buffer per stream (consecutive launches on a stream
// are ordered, so the buffer can be reused without synchronisation). A stream-ordered pool allocation per call was
// measured to add occasional 50-350 ms stalls to a layer (pool growth while persistent kernels of other strands run).
struct StreamScratch {
  float* ptr = nullptr;
  size_t bytes = 0;
};
static std::mutex g_scratch_mu;
static std::unordered_map<cudaStream_t, StreamScratch> g_scratch;

static float* stream_scratch(cudaStream_t st, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  StreamScratch& s = g_scratch[st];
  if (s.bytes < bytes) {
    if (s.ptr) {
      cudaStreamSynchronize(st);
      cudaFree(s.ptr);
      s.ptr = nullptr;
      s.bytes = 0;
    }
    size_t want = bytes < (size_t)(8 << 20) ? (size_t)(8 << 20) : bytes + bytes / 2;
    if (cudaMalloc((void**)&s.ptr, want) != cudaSuccess) {
      cudaGetLastError();
      s.ptr = nullptr;
      return nullptr;
    }
    s.bytes = want;
  }
  return s.ptr;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && spins > (1u << 26)) __trap();   // seconds: a broken pipeline must not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major operand tile in the canonical 128-byte-swizzle layout (rows of 128 B, 8-row groups of 1024 B):
// start address >> 4 | LBO (ignored for swizzled K-major) | SBO = 1024 B | version 1 (sm_100) | SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct Params {
  int M, NR, KR;            // rows, real columns (2N), real k (2K)
  long long ldc;            // floats between rows of Ct
  long long strideCb;       // floats between batches of Ct
  float* C;
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int B_TILE = BN * BKR * 4;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int BYTES = STAGES * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
tc_apply_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                const __grid_constant__ CUtensorMap mapBl, Params p) {
  using S = Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
  const uint32_t bars = base + STAGES * S::STAGE;                 // full[S] | conv[S] | empty[S] | accum | tmem ptr
  auto full = [&](int s) { return bars + 8u * s; };
  auto conv = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  const uint32_t accum = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = accum + 8u;
  auto tileAh = [&](int s) { return base + s * S::STAGE; };
  auto tileAl = [&](int s) { return base + s * S::STAGE + A_TILE; };
  auto tileBh = [&](int s) { return base + s * S::STAGE + 2 * A_TILE; };
  auto tileBl = [&](int s) { return base + s * S::STAGE + 2 * A_TILE + S::B_TILE; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // column tile fastest: the CTAs that share a 128-row block of the tall operand are scheduled next to one another, so
  // the block is read from HBM once and served from L2 to its siblings (ncu, 65536 x 256 x 256: with the row tile
  // fastest every column tile re-read the whole 134 MB operand from DRAM - 532 MB read for 4 column tiles)
  const int ntn = (p.NR + BN - 1) / BN;
  const int m0 = (int)(blockIdx.x / ntn) * BM, n0 = (int)(blockIdx.x % ntn) * BN, b = blockIdx.y;
  const int nkb = (p.KR + BKR - 1) / BKR;
  // tcgen05 adds every MMA into its fp32 TMEM accumulator with truncation, so the rounding error of an accumulator
  // grows linearly with the number of MMAs chained into it (measured with one accumulator and three MMAs per k-step:
  // 5.6e-8 relative per k-step, 4.9e-6 of max |C| at K = 256, against 1.1e-6 for the FFMA tiles). Two measures:
  // the hi.hi products go to NMAIN "main" accumulators dealt round-robin over the k-blocks (chains NMAIN times
  // shorter, one truncation per k-step), and the two small cross terms (2^-11 of the main sum, so their truncation
  // does not matter) to NMAIN "cross" accumulators. All 512 TMEM columns are used; the epilogue adds the 2 NMAIN
  // accumulators in registers with ordinary rounded fp32 adds.
  constexpr int NMAIN = 256 / BN;
  constexpr uint32_t TMEM_COLS = 512;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(conv(s), 4);
      mbar_init(empty(s), 1);
    }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapBl) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        mbar_expect_tx(full(s), A_TILE + 2 * S::B_TILE);
        tma_load_3d(tileAh(s), &mapA, full(s), kb * BKR, m0, b);
        tma_load_3d(tileBh(s), &mapBh, full(s), kb * BKR, n0, b);
        tma_load_3d(tileBl(s), &mapBl, full(s), kb * BKR, n0, b);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: D = F32 | A, B = TF32 | both K-major | N >> 3 | M >> 4
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full(s), ph);
      mbar_wait(conv(s), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint64_t ah = smem_desc_k128(tileAh(s)), al = smem_desc_k128(tileAl(s));
        const uint64_t bh = smem_desc_k128(tileBh(s)), bl = smem_desc_k128(tileBl(s));
#pragma unroll
        for (int k = 0; k < BKR / UMMA_K; ++k) {
          const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);   // 32 bytes per k-step inside the swizzle row
          const uint32_t main_acc = tmem_base + (uint32_t)((kb % NMAIN) * BN);
          const uint32_t cross_acc = main_acc + (uint32_t)(NMAIN * BN);
          const uint32_t fresh = (kb >= NMAIN || k != 0) ? 1u : 0u;
          umma_tf32(cross_acc, al + adv, bh + adv, idesc, fresh);
          umma_tf32(cross_acc, ah + adv, bl + adv, idesc, 1u);
          umma_tf32(main_acc, ah + adv, bh + adv, idesc, fresh);
        }
        umma_commit(empty(s));                 // the stage's tiles are free once these MMAs have read them
        if (kb == nkb - 1) umma_commit(accum); // ... and the accumulator is complete
      }
      __syncwarp();
    }
  } else {
    // ---------------- converter warps (2..5): split the raw A tile into tf32 hi / lo ----------------
    const int t = threadIdx.x - 64;   // 0..127
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full(s), ph);
      const uint32_t src = tileAh(s), dst = tileAl(s);
#pragma unroll
      for (int i = 0; i < A_TILE / 16 / 128; ++i) {
        const uint32_t off = (uint32_t)(i * 128 + t) * 16u;
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + off) : "memory");
        uint32_t h0 = to_tf32(v.x), h1 = to_tf32(v.y), h2 = to_tf32(v.z), h3 = to_tf32(v.w);
        uint32_t l0 = to_tf32(v.x - __uint_as_float(h0)), l1 = to_tf32(v.y - __uint_as_float(h1));
        uint32_t l2 = to_tf32(v.z - __uint_as_float(h2)), l3 = to_tf32(v.w - __uint_as_float(h3));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(src + off), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + off), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(conv(s));
    }
    // ---------------- epilogue: TMEM -> registers -> interleaved complex rows ----------------
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                      // a warp may only touch TMEM lanes [32 q, 32 q + 32)
    const int row = m0 + q * 32 + lane;
    float* crow = p.C + (long long)b * p.strideCb + (long long)row * p.ldc;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      float sum[32];
      const int used = nkb < NMAIN ? nkb : NMAIN;      // main accumulators that received a k-block (cross likewise)
      for (int a = 0; a < 2 * used; ++a) {
        const int slot = a < used ? a : NMAIN + (a - used);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * BN + c);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = a == 0 ? __uint_as_float(r[j]) : sum[j] + __uint_as_float(r[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(sum[j]);
      if (row < p.M) {
        const int col = n0 + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col + j + 3 < p.NR) {
            float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            *reinterpret_cast<float4*>(crow + col + j) = o;
          } else {
            for (int e = 0; e < 4; ++e)
              if (col + j + e < p.NR) crow[col + j + e] = __uint_as_float(r[j + e]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- preparation of the small operand: Bt^T [batch, 2N, 2K] (K-major), split into tf32 hi / lo ---------------
__global__ void __launch_bounds__(256) prep_b_kernel(int K, int N, int batch, mpdo_idxmap Bb, mpdo_idxmap Bk, mpdo_idxmap Bj,
                                                     const float2* __restrict__ B, int conjA, int conjB, float alpha,
                                                     float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = (long long)batch * N * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const int n = (int)((idx / K) % N);
    const int bb = (int)(idx / ((long long)K * N));
    float2 v = B[map_idx(Bb, bb) + map_idx(Bk, k) + map_idx(Bj, n)];
    v.x *= alpha;
    v.y *= conjB ? -alpha : alpha;
    // row (n,0) = real part of the result, row (n,1) = imaginary part; columns (k,0), (k,1) multiply (Ar, Ai)
    const float e00 = v.x, e01 = conjA ? v.y : -v.y, e10 = v.y, e11 = conjA ? -v.x : v.x;
    const long long r0 = ((long long)bb * 2 * N + 2 * n) * (2LL * K) + 2 * k, r1 = r0 + 2LL * K;
    auto put = [&](long long at, float x) {
      const float h = __uint_as_float(to_tf32(x));
      hi[at] = h;
      lo[at] = __uint_as_float(to_tf32(x - h));
    };
    put(r0, e00);
    put(r0 + 1, e01);
    put(r1, e10);
    put(r1 + 1, e11);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)ptr;
    else
      cudaGetLastError();
  }
  return fn;
}

// fp32 matrix [batch, rows, cols] with cols contiguous -> 3-D map, box {32 floats, boxRows, 1}, 128-byte swizzle;
// out-of-range elements read as zero
static bool make_map(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long batch, long long ld,
                     long long strideB, int boxRows) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batch > 1 ? strideB : ld * rows) * 4};
  cuuint32_t box[3] = {(cuuint32_t)BKR, (cuuint32_t)boxRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool single_level(const mpdo_idxmap& m) { return m.d0 <= 0; }

template <int BN, int STAGES>
static int launch(const CUtensorMap& mA, const CUtensorMap& mBh, const CUtensorMap& mBl, const Params& p, int batch,
                  cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    MPDO_CUDA(cudaFuncSetAttribute(tc_apply_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
    configured = true;
  }
  dim3 grid((unsigned)(((p.M + BM - 1) / BM) * ((p.NR + BN - 1) / BN)), (unsigned)batch);
  tc_apply_kernel<BN, STAGES><<<grid, 192, S::BYTES, st>>>(mA, mBh, mBl, p);
  return check_launch("tc_apply_kernel");
}

// Returns 1 when the contraction was issued on the tensor-core path, 0 when it is not eligible (the caller falls
// back to the SIMT kernel), or an error code < 0 / > 1.
int try_apply(const mpdo_contract_desc& d, const void* A, const void* B, void* C, cudaStream_t st) {
  if (!g_enabled.load(std::memory_order_relaxed)) return 0;
  if (d.dtypeA != MPDO_C64 || d.dtypeB != MPDO_C64 || d.dtypeC != MPDO_C64 || d.acc64 || d.hermitian) return 0;
  if (d.beta != 0.0 || d.ksplit > 1) return 0;
  // the big operand: plain row-major matrices per batch (k contiguous), result likewise (j contiguous)
  if (!single_level(d.Ai) || !single_level(d.Ak) || !single_level(d.Ab) || d.Ak.s0 != 1) return 0;
  if (!single_level(d.Ci) || !single_level(d.Cj) || !single_level(d.Cb) || d.Cj.s0 != 1) return 0;
  if (d.K < 16 || d.N < 8 || (d.K & 1) || (d.N & 1) || (d.Ai.s0 & 1) || (d.Ci.s0 & 1)) return 0;
  if ((long long)d.M * d.K * d.N < (1LL << 24) || d.M < 4 * BM) return 0;     // small products stay on the SIMT tiles
  if (((uintptr_t)A & 15) || ((uintptr_t)C & 15) || (d.batch > 1 && ((d.Ab.s0 & 1) || (d.Cb.s0 & 1)))) return 0;
  if (d.batch > 65535) return 0;
  if (!encode_fn()) return 0;

  const int KR = 2 * d.K, NR = 2 * d.N;
  const size_t bElems = (size_t)d.batch * NR * KR;
  float* scratch = stream_scratch(st, 2 * bElems * sizeof(float));
  if (!scratch) return 0;
  float *bh = scratch, *bl = scratch + bElems;
  {
    const long long total = (long long)d.batch * d.N * d.K;
    const unsigned gx = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
    prep_b_kernel<<<gx, 256, 0, st>>>(d.K, d.N, d.batch, d.Bb, d.Bk, d.Bj, (const float2*)B, d.conjA, d.conjB, (float)d.alpha,
                                      bh, bl);
    int rc = check_launch("prep_b_kernel");
    if (rc) return rc;
  }
  // column tile: as wide as the result allows, but narrow enough that the accumulator chains stay short
  // (K / (4 NMAIN) k-steps per main accumulator: <= 32 up to K = 512)
  int BN = NR > 128 ? 256 : (NR > 64 ? 128 : 64);
  if (g_enabled.load(std::memory_order_relaxed) != 2) {   // mode 2 ("fast") keeps the widest tile whatever K
    if (d.K > 256) BN = 64;
    else if (d.K > 128 && BN > 128) BN = 128;
  }
  CUtensorMap mA, mBh, mBl;
  const bool ok = make_map(&mA, A, KR, d.M, d.batch, 2 * d.Ai.s0, 2 * d.Ab.s0, BM) &&
                  make_map(&mBh, bh, KR, NR, d.batch, KR, (long long)NR * KR, BN) &&
                  make_map(&mBl, bl, KR, NR, d.batch, KR, (long long)NR * KR, BN);
  if (!ok) return 0;
  Params p;
  p.M = d.M;
  p.NR = NR;
  p.KR = KR;
  p.ldc = 2 * d.Ci.s0;
  p.strideCb = 2 * d.Cb.s0;
  p.C = (float*)C;
  int rc;
  {
    const double mnk = (double)d.M * d.N * d.K * d.batch;
    TimedLaunch timed(4, 8.0 * mnk, 8.0 * d.batch * ((double)d.M * d.K + (double)d.K * d.N + (double)d.M * d.N), st);
    if (BN == 256)
      rc = launch<256, 2>(mA, mBh, mBl, p, d.batch, st);
    else if (BN == 128)
      rc = launch<128, 3>(mA, mBh, mBl, p, d.batch, st);
    else
      rc = launch<64, 4>(mA, mBh, mBl, p, d.batch, st);
  }
  return rc ? rc : 1;
}

}  // namespace tc
}  // namespace mpdo

extern "C" int mpdo_tc_enable(int mode) {
  return mpdo::tc::g_enabled.exchange(mode < 0 ? 0 : (mode > 2 ? 2 : mode));
}
