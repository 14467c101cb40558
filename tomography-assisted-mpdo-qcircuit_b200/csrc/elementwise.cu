// HBM-bound helpers of the MPDO update path (sm_100a): single-qubit gate / Kraus absorption,
// row scaling of small cores, the reference's kept-rank rule, dtype casts; plus library info.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpdo {

thread_local char g_err[512] = "";
std::atomic<int> g_timing{0};

namespace {
struct TimingRec {
  cudaEvent_t e0, e1;
  int cls;
  double flops, bytes;
};
std::mutex g_tmutex;
std::vector<TimingRec*> g_trecs;
}  // namespace

void timing_begin(int cls, double flops, double bytes, cudaStream_t st, void** token) {
  TimingRec* r = new TimingRec;
  r->cls = cls;
  r->flops = flops;
  r->bytes = bytes;
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, st);
  *token = r;
}

void timing_end(void* token, cudaStream_t st) {
  TimingRec* r = (TimingRec*)token;
  cudaEventRecord(r->e1, st);
  std::lock_guard<std::mutex> lock(g_tmutex);
  g_trecs.push_back(r);
}

std::atomic<long long> g_launches{0};

// Tout[bl, p, g, e] = sum_s G[p, s, g] * T[bl, s, e]     (e = (a, r) flattened, bl = (batch, l))
// One thread per (bl, e): two coalesced loads, 2K coalesced stores. Algorithmic traffic:
// c * l * r * 2a * (1 + K) bytes per batch entry.
template <typename CT>
__global__ void __launch_bounds__(256) absorb_1q_kernel(long long total, int l, long long E, int K,
                                                        const CT* __restrict__ T, const CT* __restrict__ G,
                                                        long long gBatchStride, CT* __restrict__ Tout) {
  extern __shared__ unsigned char smraw[];
  CT* g = reinterpret_cast<CT*>(smraw);  // [2][2][K] of the first batch entry touched by this CTA
  const long long idx0 = (long long)blockIdx.x * blockDim.x;
  const long long perBatch = (long long)l * E;
  const long long bFirst = idx0 / perBatch;
  const long long idxLast = min(total - 1, idx0 + blockDim.x - 1);
  const long long bLast = idxLast / perBatch;
  const bool uniform = (gBatchStride == 0) || (bFirst == bLast);
  if (uniform) {
    const CT* gs = G + bFirst * gBatchStride;
    for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) g[i] = gs[i];
  }
  __syncthreads();
  const long long idx = idx0 + threadIdx.x;
  if (idx >= total) return;
  const long long bl = idx / E, e = idx % E;
  const CT* gp = uniform ? g : G + (idx / perBatch) * gBatchStride;
  const CT t0 = T[(bl * 2 + 0) * E + e];
  const CT t1 = T[(bl * 2 + 1) * E + e];
  CT* out = Tout + bl * 2 * K * E + e;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    for (int k = 0; k < K; ++k) {
      const CT g0 = gp[(p * 2 + 0) * K + k], g1 = gp[(p * 2 + 1) * K + k];
      CT o;
      o.x = g0.x * t0.x - g0.y * t0.y + g1.x * t1.x - g1.y * t1.y;
      o.y = g0.x * t0.y + g0.y * t0.x + g1.x * t1.y + g1.y * t1.x;
      out[((long long)p * K + k) * E] = o;
    }
  }
}

template <typename CX>
__global__ void __launch_bounds__(256) rowscale_kernel(int rows, int cols, int vRows, const double2* __restrict__ V,
                                                       const double* __restrict__ lam, int lamStride, double power,
                                                       double tol, int mode, CX* __restrict__ X) {
  const int b = blockIdx.y;
  const double lmax = lam[(long long)b * lamStride];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)rows * cols;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / cols), c = (int)(idx % cols);
    double lv = lam[(long long)b * lamStride + j];
    double f;
    const double thr = tol * lmax;
    if (mode == 0) {
      f = (lv > thr && lv > 0) ? pow(lv, power) : 0.0;
    } else {
      lv = fmax(lv, thr);
      f = lv > 0 ? pow(lv, power) : 0.0;
    }
    const double2 v = V[((long long)b * vRows + j) * cols + c];
    CX o;
    o.x = (typename real_of<CX>::type)(f * v.x);
    o.y = (typename real_of<CX>::type)(f * v.y);
    X[((long long)b * rows + j) * cols + c] = o;
  }
}

// decompositions.py:117-134, one thread per batch entry (n is a few hundred at most).
__global__ void rank_rule_kernel(int batch, int n, double* __restrict__ s, int sStride, int squared, int cap,
                                 double maxTruncErr, int relative, int f32, int32_t* __restrict__ keep,
                                 int zeroTail) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double* sv = s + (long long)b * sStride;
  auto sval = [&](int i) { return squared ? sqrt(fmax(sv[i], 0.0)) : sv[i]; };
  int numErr = cap;
  if (maxTruncErr >= 0 && n > 0) {
    if (f32) {
      double run = 0;  // torch CPU cumsum on fp32: fp64 running sum, rounded to fp32 per element
      for (int i = 0; i < n; ++i) {
        float x = (float)sval(i);
        run += (double)(x * x);
      }
      const float last = sqrtf((float)run);
      const float eps = relative ? (float)((float)maxTruncErr * (float)sval(0)) : (float)maxTruncErr;
      double acc = 0;
      for (int i = 0; i < n - 1; ++i) {
        float x = (float)sval(i);
        acc += (double)(x * x);
        const float ti = sqrtf((float)acc);
        if (last - ti <= eps) {
          numErr = i + 1;
          break;
        }
      }
    } else {
      double run = 0;
      for (int i = 0; i < n; ++i) run += sval(i) * sval(i);
      const double last = sqrt(run);
      const double eps = relative ? maxTruncErr * sval(0) : maxTruncErr;
      double acc = 0;
      for (int i = 0; i < n - 1; ++i) {
        acc += sval(i) * sval(i);
        if (last - sqrt(acc) <= eps) {
          numErr = i + 1;
          break;
        }
      }
    }
  }
  const int k = numErr < cap ? numErr : cap;
  keep[b] = k;
  if (zeroTail)
    for (int i = k; i < n; ++i) sv[i] = 0.0;
}

template <typename CI, typename CO>
__global__ void __launch_bounds__(256) cast_kernel(long long count, const CI* __restrict__ in, CO* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = cconv<CO>(in[i]);
}

}  // namespace mpdo

extern "C" int mpdo_absorb_1q(int dtype, int batch, int l, int a, int r, int K, const void* T, const void* G,
                              int64_t gBatchStride, void* Tout, void* stream) {
  using namespace mpdo;
  if (batch <= 0 || l <= 0 || a <= 0 || r <= 0) return 0;
  if (!T || !G || !Tout || K <= 0 || K > 64) return fail(MPDO_EINVAL, "mpdo_absorb_1q: bad argument");
  const long long E = (long long)a * r;
  const long long total = (long long)batch * l * E;
  const long long blocks = (total + 255) / 256;
  if (blocks > 2147483647LL) return fail(MPDO_EINVAL, "mpdo_absorb_1q: too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MPDO_C64) {
    absorb_1q_kernel<float2><<<(unsigned)blocks, 256, 4 * K * sizeof(float2), st>>>(
        total, l, E, K, (const float2*)T, (const float2*)G, gBatchStride, (float2*)Tout);
  } else {
    absorb_1q_kernel<double2><<<(unsigned)blocks, 256, 4 * K * sizeof(double2), st>>>(
        total, l, E, K, (const double2*)T, (const double2*)G, gBatchStride, (double2*)Tout);
  }
  return check_launch("absorb_1q_kernel");
}

extern "C" int mpdo_rowscale(int batch, int rows, int cols, int vRows, const void* V, const double* lam,
                             int lamStride, double power, double tol, int mode, int dtypeX, void* X, void* stream) {
  using namespace mpdo;
  if (batch <= 0 || rows <= 0 || cols <= 0) return 0;
  if (!V || !lam || !X || rows > vRows) return fail(MPDO_EINVAL, "mpdo_rowscale: bad argument");
  long long tot = (long long)rows * cols;
  unsigned gx = (unsigned)((tot + 255) / 256 > 1024 ? 1024 : (tot + 255) / 256);
  if (batch > 65535) return fail(MPDO_EINVAL, "mpdo_rowscale: batch > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtypeX == MPDO_C64)
    rowscale_kernel<float2><<<dim3(gx, batch), 256, 0, st>>>(rows, cols, vRows, (const double2*)V, lam, lamStride,
                                                            power, tol, mode, (float2*)X);
  else
    rowscale_kernel<double2><<<dim3(gx, batch), 256, 0, st>>>(rows, cols, vRows, (const double2*)V, lam, lamStride,
                                                             power, tol, mode, (double2*)X);
  return check_launch("rowscale_kernel");
}

extern "C" int mpdo_rank_rule(int batch, int n, double* s, int sStride, int squared, int cap, double maxTruncErr,
                              int relative, int f32, int32_t* keep, int zeroTail, void* stream) {
  using namespace mpdo;
  if (batch <= 0) return 0;
  if (!s || !keep) return fail(MPDO_EINVAL, "mpdo_rank_rule: null argument");
  rank_rule_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(batch, n, s, sStride, squared, cap,
                                                                         maxTruncErr, relative, f32, keep, zeroTail);
  return check_launch("rank_rule_kernel");
}

extern "C" int mpdo_cast(int dtypeIn, int dtypeOut, int64_t count, const void* in, void* out, void* stream) {
  using namespace mpdo;
  if (count <= 0) return 0;
  if (!in || !out) return fail(MPDO_EINVAL, "mpdo_cast: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned gx = (unsigned)((count + 255) / 256 > 148 * 16 ? 148 * 16 : (count + 255) / 256);
  if (dtypeIn == MPDO_C64 && dtypeOut == MPDO_C128)
    cast_kernel<float2, double2><<<gx, 256, 0, st>>>(count, (const float2*)in, (double2*)out);
  else if (dtypeIn == MPDO_C128 && dtypeOut == MPDO_C64)
    cast_kernel<double2, float2><<<gx, 256, 0, st>>>(count, (const double2*)in, (float2*)out);
  else if (dtypeIn == MPDO_C64)
    cast_kernel<float2, float2><<<gx, 256, 0, st>>>(count, (const float2*)in, (float2*)out);
  else
    cast_kernel<double2, double2><<<gx, 256, 0, st>>>(count, (const double2*)in, (double2*)out);
  return check_launch("cast_kernel");
}

extern "C" int mpdo_timing_enable(int on) {
  using namespace mpdo;
  std::lock_guard<std::mutex> lock(g_tmutex);
  for (TimingRec* r : g_trecs) {
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  g_trecs.clear();
  g_timing.store(on ? 1 : 0);
  return 0;
}

extern "C" int mpdo_timing_summary(int cls, double minFlops, double* seconds, double* flops, double* bytes,
                                   int64_t* launches, double* maxFlopsSeconds, double* maxFlops) {
  using namespace mpdo;
  MPDO_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_tmutex);
  double sec = 0, fl = 0, by = 0, bestF = -1, bestS = 0;
  int64_t n = 0;
  for (TimingRec* r : g_trecs) {
    if (r->cls != cls || r->flops < minFlops) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r->e0, r->e1) != cudaSuccess) continue;
    sec += ms * 1e-3;
    fl += r->flops;
    by += r->bytes;
    ++n;
    if (r->flops > bestF) {
      bestF = r->flops;
      bestS = ms * 1e-3;
    }
  }
  if (seconds) *seconds = sec;
  if (flops) *flops = fl;
  if (bytes) *bytes = by;
  if (launches) *launches = n;
  if (maxFlopsSeconds) *maxFlopsSeconds = bestS;
  if (maxFlops) *maxFlops = bestF < 0 ? 0 : bestF;
  return 0;
}

// Debug / test hook: a two-CTA launch in which CTA 1 never arrives at the device-wide barrier, with a short poll limit.
// The waiting CTA must trap (see matrix_barrier); the call returns the resulting CUDA error.
__global__ void barrier_timeout_kernel(unsigned* bar, int* flag) {
  unsigned phase = 0;
  if (blockIdx.x == 0) mpdo::matrix_barrier(bar, 2u, phase, flag, 1u << 12);
}

extern "C" int mpdo_debug_barrier_timeout(void* stream) {
  using namespace mpdo;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* bar = nullptr;
  MPDO_CUDA(cudaMalloc((void**)&bar, 16));
  MPDO_CUDA(cudaMemsetAsync(bar, 0, 16, st));
  barrier_timeout_kernel<<<2, 32, 0, st>>>(bar, reinterpret_cast<int*>(bar) + 2);
  int rc = check_launch("barrier_timeout_kernel");
  if (rc) return rc;
  MPDO_CUDA(cudaStreamSynchronize(st));   // returns the launch failure raised by the trap
  return 0;
}

extern "C" int mpdo_version(void) { return 100; }
extern "C" const char* mpdo_last_error(void) { return mpdo::g_err; }
extern "C" int64_t mpdo_launch_count(void) { return mpdo::g_launches.load(); }
extern "C" int mpdo_device_info(int* smCount, int* smemPerBlockOptin, int* ccMajor, int* ccMinor) {
  using namespace mpdo;
  int dev = 0;
  MPDO_CUDA(cudaGetDevice(&dev));
  if (smCount) MPDO_CUDA(cudaDeviceGetAttribute(smCount, cudaDevAttrMultiProcessorCount, dev));
  if (smemPerBlockOptin) MPDO_CUDA(cudaDeviceGetAttribute(smemPerBlockOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (ccMajor) MPDO_CUDA(cudaDeviceGetAttribute(ccMajor, cudaDevAttrComputeCapabilityMajor, dev));
  if (ccMinor) MPDO_CUDA(cudaDeviceGetAttribute(ccMinor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}
