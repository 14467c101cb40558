// Shared helpers for the mpdo_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mpdo_b200.h"

namespace mpdo {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  ++g_launches;
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define MPDO_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      snprintf(mpdo::g_err, sizeof(mpdo::g_err), "%s: %s", #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                                     \
    }                                                                                     \
  } while (0)

// ---- complex helpers (interleaved re/im in float2 / double2) -------------------------------
template <typename R> struct cplx;
template <> struct cplx<float> { using type = float2; };
template <> struct cplx<double> { using type = double2; };

template <typename C> struct real_of;
template <> struct real_of<float2> { using type = float; };
template <> struct real_of<double2> { using type = double; };

template <typename CO, typename CI> __device__ __forceinline__ CO cconv(CI v) {
  CO o;
  o.x = (typename real_of<CO>::type)v.x;
  o.y = (typename real_of<CO>::type)v.y;
  return o;
}

// acc += a * b
template <typename C> __device__ __forceinline__ void cfma(C& acc, const C a, const C b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ long long map_idx(const mpdo_idxmap& m, int i) {
  if (m.d0 <= 0) return (long long)i * m.s0;
  int i0 = i % m.d0;
  int t = i / m.d0;
  if (m.d1 <= 0) return (long long)i0 * m.s0 + (long long)t * m.s1;
  int i1 = t % m.d1;
  int i2 = t / m.d1;
  return (long long)i0 * m.s0 + (long long)i1 * m.s1 + (long long)i2 * m.s2;
}

#ifdef __CUDACC__
// Device-wide barrier between the CTAs that work on one matrix (grid.x of them; the kernel is launched
// cooperatively, so they are co-resident). `bar` only ever increases: round `phase` completes at (phase+1)*nblk.
// A CTA that does not see the others arrive within `limit` polls (seconds at the default: never expected, it means the
// CTAs are not co-resident or one of them died) sets *errflag and TRAPS: the kernel aborts and every later CUDA call of
// the process reports the failure, instead of a factorisation that silently used half-updated data.
__device__ __forceinline__ void matrix_barrier(unsigned* bar, unsigned nblk, unsigned& phase, int* errflag,
                                               unsigned limit = 1u << 24) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // arrive: release at gpu scope (cumulative over the CTA's writes ordered by the bar.sync above), no return value
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    const unsigned target = (phase + 1u) * nblk;
    unsigned spins = 0, seen;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (seen >= target) break;
      if (++spins > limit) {  // refuse to hang the device, and refuse to continue with inconsistent data
        *errflag = 1;
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncthreads();
  ++phase;
}
#endif

// tc_apply.cu: tcgen05 / TMA path of the complex64 apply contractions. 1 = issued, 0 = not eligible, else error.
namespace tc {
int try_apply(const mpdo_contract_desc& d, const void* A, const void* B, void* C, cudaStream_t st);
}

// jacobi.cu: mpdo_jacobi_rows with an optional per-matrix count of non-zero leading rows (device memory)
int jacobi_rows_ranked(int batch, int n, int m, int mt, int ld, long long batchStride, void* Y, double tol,
                       int maxSweeps, int32_t* work, const int* rank, int rankStride, void* stream);

// ---- optional per-launch timing (bench.py's roofline leg): CUDA events recorded on the launching stream ----------
extern std::atomic<int> g_timing;
void timing_begin(int cls, double flops, double bytes, cudaStream_t st, void** token);
void timing_end(void* token, cudaStream_t st);

struct TimedLaunch {  // RAII: brackets the launches issued in its scope when timing is enabled
  void* token = nullptr;
  cudaStream_t st;
  TimedLaunch(int cls, double flops, double bytes, cudaStream_t s) : st(s) {
    if (g_timing.load(std::memory_order_relaxed)) timing_begin(cls, flops, bytes, s, &token);
  }
  ~TimedLaunch() {
    if (token) timing_end(token, st);
  }
};

}  // namespace mpdo
