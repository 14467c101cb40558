// Native engine of the MPDO update path: the kernel sequences of one gate split, one QR-sweep step, one
// chi-truncation step and one kappa truncation, issued from C++ on the caller's stream.
//
// The reference runs these steps through tensornetwork + LAPACK (Circuit.py:74-136, TNNOptimizer.py:87-197);
// MPDOSimulator/_engine/steps.py states the same sequences over the Python primitive wrappers and documents the
// mathematics (Gram-eig orthogonalisation, isometry shortcut, core split, top-kappa subspace iteration). This file
// is the production path: one C call per step instead of ~50 Python-level launches, which is what bounds a layer
// once the kernels themselves take tens of microseconds.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>
#include <atomic>

#include "engine.cuh"

namespace mpdo {
namespace eng {

// ---------------------------------------------------------------------------------------------------
// contraction descriptor from strided views (port of prims.py:_levels/_idxmap)
// ---------------------------------------------------------------------------------------------------
struct Level {
  long long n, s;
};

static int collapse(const Tn& t, int d0, int cnt, Level* lv) {
  int k = 0;
  for (int d = d0; d < d0 + cnt; ++d) {
    long long n = t.sh[d], s = t.st[d];
    if (n == 1) continue;
    if (k > 0 && lv[k - 1].s == s * n) {
      lv[k - 1].n *= n;
      lv[k - 1].s = s;
    } else {
      lv[k].n = n;
      lv[k].s = s;
      ++k;
    }
  }
  return k;
}

static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {nullptr};

cudaMemPool_t scratch_pool() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;   // Arena falls back to the default pool
    }
    unsigned long long keep = ~0ULL;
    int off = 0, on = 1;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowOpportunistic, &on);
    cudaMemPoolSetAttribute(pool, cudaMemPoolReuseFollowEventDependencies, &on);
    g_pools[dev] = pool;
    // Pre-grow the pool once. With cross-stream reuse limited to completed frees, whether an allocation finds a
    // cached block depends on how far the other strands' streams have run, so an un-grown pool keeps growing at
    // random moments of steady-state operation, and each growth maps fresh physical memory (measured: single bench
    // passes 2-3x slower than their neighbours). One block of MPDO_SCRATCH_PREWARM_MB (default 16 GiB, capped at a
    // quarter of the free device memory) is allocated and freed here; it stays cached (release threshold = max).
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
      size_t want = (size_t)16384 << 20;
      if (const char* e = getenv("MPDO_SCRATCH_PREWARM_MB")) want = (size_t)strtoull(e, nullptr, 10) << 20;
      if (want > freeB / 4) want = freeB / 4;
      if (want >= ((size_t)1 << 20)) {
        void* warm = nullptr;
        if (cudaMallocFromPoolAsync(&warm, want, pool, (cudaStream_t)0) == cudaSuccess) {
          cudaFreeAsync(warm, (cudaStream_t)0);
          cudaStreamSynchronize((cudaStream_t)0);
        } else {
          cudaGetLastError();
        }
      }
    }
  }
  return g_pools[dev];
}

void trim_all_pools() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (cudaMemPool_t p : g_pools)
    if (p) cudaMemPoolTrimTo(p, 0);
}

static bool to_map(const Level* lv, int k, mpdo_idxmap* m) {
  memset(m, 0, sizeof(*m));
  if (k == 0) return true;
  if (k == 1) {
    m->s0 = lv[0].s;
  } else if (k == 2) {
    m->d0 = (int)lv[1].n;
    m->s0 = lv[1].s;
    m->s1 = lv[0].s;
  } else if (k == 3) {
    m->d0 = (int)lv[2].n;
    m->d1 = (int)lv[1].n;
    m->s0 = lv[2].s;
    m->s1 = lv[1].s;
    m->s2 = lv[0].s;
  } else {
    return false;
  }
  return true;
}

static long long prod(const Tn& t, int d0, int cnt) {
  long long p = 1;
  for (int d = d0; d < d0 + cnt; ++d) p *= t.sh[d];
  return p;
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int contract(cudaStream_t st, const Tn& A, Roles ra, const Tn& B, Roles rb, const Tn& C, Roles rc, bool conjA,
             bool conjB, int acc64, double alpha, double beta, bool hermitian) {
  mpdo_contract_desc d;
  memset(&d, 0, sizeof(d));
  const long long M = prod(A, ra.nb, ra.n1), K = prod(A, ra.nb + ra.n1, ra.n2);
  const long long N = prod(B, rb.nb + rb.n1, rb.n2), batch = prod(A, 0, ra.nb);
  if (prod(B, rb.nb, rb.n1) != K || prod(C, rc.nb, rc.n1) != M || prod(C, rc.nb + rc.n1, rc.n2) != N ||
      prod(B, 0, rb.nb) != batch || prod(C, 0, rc.nb) != batch)
    return fail(MPDO_EINVAL, "engine contract: inconsistent extents");
  d.M = (int)M;
  d.N = (int)N;
  d.K = (int)K;
  d.batch = (int)batch;
  d.dtypeA = A.dt;
  d.dtypeB = B.dt;
  d.dtypeC = C.dt;
  d.conjA = conjA;
  d.conjB = conjB;
  const bool all64 = A.dt == MPDO_C64 && B.dt == MPDO_C64 && C.dt == MPDO_C64;
  d.acc64 = acc64 >= 0 ? acc64 : !all64;
  Level lab[MAXD], lai[MAXD], lak[MAXD], lbb[MAXD], lbk[MAXD], lbj[MAXD], lcb[MAXD], lci[MAXD], lcj[MAXD];
  const int nab = collapse(A, 0, ra.nb, lab), nai = collapse(A, ra.nb, ra.n1, lai),
            nak = collapse(A, ra.nb + ra.n1, ra.n2, lak);
  const int nbb = collapse(B, 0, rb.nb, lbb), nbk = collapse(B, rb.nb, rb.n1, lbk),
            nbj = collapse(B, rb.nb + rb.n1, rb.n2, lbj);
  const int ncb = collapse(C, 0, rc.nb, lcb), nci = collapse(C, rc.nb, rc.n1, lci),
            ncj = collapse(C, rc.nb + rc.n1, rc.n2, lcj);
  if (!to_map(lab, nab, &d.Ab) || !to_map(lai, nai, &d.Ai) || !to_map(lak, nak, &d.Ak) || !to_map(lbb, nbb, &d.Bb) ||
      !to_map(lbk, nbk, &d.Bk) || !to_map(lbj, nbj, &d.Bj) || !to_map(lcb, ncb, &d.Cb) || !to_map(lci, nci, &d.Ci) ||
      !to_map(lcj, ncj, &d.Cj))
    return fail(MPDO_EINVAL, "engine contract: an axis group needs more than 3 stride levels");
  auto inner = [](const Level* lv, int k) { return k ? lv[k - 1].s : (1LL << 60); };
  d.a_kfast = inner(lak, nak) <= inner(lai, nai);
  d.b_jfast = inner(lbj, nbj) <= inner(lbk, nbk);
  const long long tm_ = (M + 63) / 64;
  const long long tiles = (hermitian ? tm_ * (tm_ + 1) / 2 : tm_ * ((N + 63) / 64)) * batch;
  d.hermitian = hermitian ? 1 : 0;
  int ksplit = 1;
  const int sms = sm_count();
  // Gram matrices of the sweeps are one or two output tiles with a long K: a single CTA walking K alone is pure
  // latency (measured 90 us for 48x48x384), so K is spread over CTAs as soon as the grid is small.
  if (K >= 256 && tiles < sms && C.contiguous() && (beta == 0.0 || beta == 1.0))
    ksplit = (int)std::max(1LL, std::min((K + 63) / 64, (2LL * sms) / std::max(tiles, 1LL)));
  d.ksplit = ksplit;
  d.alpha = alpha;
  d.beta = beta;
  if (ksplit > 1 && beta == 0.0) MPDO_CUDA(cudaMemsetAsync(C.p, 0, (size_t)C.numel() * C.esz(), st));
  return mpdo_contract(&d, A.p, B.p, C.p, st);
}

// ---------------------------------------------------------------------------------------------------
// strided copy with dtype conversion / conjugation (permutes and casts that torch did in the Python engine)
// ---------------------------------------------------------------------------------------------------
struct CopyArgs {
  int nd;
  long long sh[MAXD], st[MAXD];
};

template <typename CI, typename CO>
__global__ void __launch_bounds__(256) copy_view_kernel(CopyArgs a, long long total, const CI* __restrict__ in,
                                                        CO* __restrict__ out, int conj) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long rem = idx, off = 0;
    for (int d = a.nd - 1; d >= 0; --d) {
      const long long c = rem % a.sh[d];
      rem /= a.sh[d];
      off += c * a.st[d];
    }
    CI v = in[off];
    CO o = cconv<CO>(v);
    if (conj) o.y = -o.y;
    out[idx] = o;
  }
}

static int copy_view(cudaStream_t st, const Tn& in, const Tn& out, bool conj = false) {  // out contiguous
  CopyArgs a;
  a.nd = in.nd;
  for (int i = 0; i < in.nd; ++i) {
    a.sh[i] = in.sh[i];
    a.st[i] = in.st[i];
  }
  const long long total = in.numel();
  if (total == 0) return 0;
  unsigned gx = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  if (in.dt == MPDO_C64 && out.dt == MPDO_C64)
    copy_view_kernel<float2, float2><<<gx, 256, 0, st>>>(a, total, (const float2*)in.p, (float2*)out.p, conj);
  else if (in.dt == MPDO_C64)
    copy_view_kernel<float2, double2><<<gx, 256, 0, st>>>(a, total, (const float2*)in.p, (double2*)out.p, conj);
  else if (out.dt == MPDO_C64)
    copy_view_kernel<double2, float2><<<gx, 256, 0, st>>>(a, total, (const double2*)in.p, (float2*)out.p, conj);
  else
    copy_view_kernel<double2, double2><<<gx, 256, 0, st>>>(a, total, (const double2*)in.p, (double2*)out.p, conj);
  return check_launch("copy_view_kernel");
}

// residual norms of the kept Ritz pairs, relative to the leading Ritz value: out[b] = max_j |R[b,j,:]| / theta[b,0]
__global__ void __launch_bounds__(256) residual_kernel(int k, int n, const double2* __restrict__ R,
                                                       const double* __restrict__ theta, int thetaStride,
                                                       double* __restrict__ out) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double worst = 0;
  for (int j = 0; j < k; ++j) {
    const double2* row = R + ((long long)b * k + j) * n;
    double a = 0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) a = fma(row[c].x, row[c].x, fma(row[c].y, row[c].y, a));
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
      worst = fmax(worst, sqrt(s));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b] = worst / fmax(theta[(long long)b * thetaStride], 1e-300);
}

// trace of G minus the kept eigenvalues -> norm of the discarded part (diagnostic output of the kappa step)
__global__ void discarded_kernel(int B, int n, int k, const double2* __restrict__ G, const double* __restrict__ theta,
                                 int thetaStride, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double tr = 0;
  for (int i = 0; i < n; ++i) tr += G[((long long)b * n + i) * n + i].x;
  for (int j = 0; j < k; ++j) tr -= theta[(long long)b * thetaStride + j];
  out[b] = sqrt(fmax(tr, 0.0));
}

// ---------------------------------------------------------------------------------------------------
// engine context and small-core helpers
// ---------------------------------------------------------------------------------------------------
struct Ctx {
  cudaStream_t st;
  Arena ar;
  int dt;      // dtype of the state tensors
  bool f32;
  int npass;   // 1: Gram-eig once (complex64), 2: twice + Jacobi SVD of the core (complex128)
  double null_tol = 1e-14, floor_tol = 1e-13, jtol = 1e-15, chol_rel = 1e-15;
  Ctx(cudaStream_t s, int dtype, int np) : st(s), ar(s), dt(dtype), f32(dtype == MPDO_C64), npass(np) {
    // complex64 states are stored in fp32: rows orthogonal to 1e-10 relative are far below what the state can
    // represent, and Jacobi converges quadratically, so this saves the last sweep or two of every decomposition
    if (f32) {
      static const double jt32 = getenv("MPDO_JTOL32") ? atof(getenv("MPDO_JTOL32")) : 1e-10;   // tuning knob
      jtol = jt32;
    }
    static const bool noChol = getenv("MPDO_NO_CHOLQR") != nullptr;   // debugging knob
    use_chol = !noChol;
  }
  bool use_chol = true;
  // Large batches (parameter sweeps, cfg4): the factorisations that need co-resident CTA groups (pivoted Cholesky
  // and persistent Jacobi beyond one CTA's shared memory) walk the batch in co-resident chunks, so a batch takes the
  // same Cholesky-QR / preconditioned routes as a single circuit. MPDO_BATCH_CLASSIC=1 restores the first policy
  // of the round (classic Jacobi on [G | I] for B >= 32) for A/B measurements.
  void set_batch(long long B) {
    static const bool classic = getenv("MPDO_BATCH_CLASSIC") != nullptr;
    if (classic && B >= 32) {
      use_chol = false;
      precondition = false;
    }
  }
  bool precondition = true;
};

#define EC(call)            \
  do {                      \
    int rc_ = (call);       \
    if (rc_ != 0) return rc_; \
  } while (0)
#define ARENA_OK(c)                  \
  do {                               \
    if ((c).ar.err) return (c).ar.err; \
  } while (0)

static int decompose(Ctx& c, const Tn& L, bool wantRows, double** s, Tn* Yn, Tn* Z) {
  const long long B = L.sh[0], n = L.sh[1], m = L.sh[2];
  Tn Y = c.ar.alloc(MPDO_C128, {B, n, m + n});
  int32_t* work = (int32_t*)c.ar.raw(sizeof(int32_t) * 48 * (size_t)B);
  *s = c.ar.reals(B * n);
  *Z = c.ar.alloc(MPDO_C128, {B, n, n});
  if (wantRows) *Yn = c.ar.alloc(MPDO_C128, {B, n, m});
  ARENA_OK(c);
  const double tol = std::max(c.jtol, 4.4e-16 * sqrt((double)m));
  return mpdo_decompose_rows((int)B, (int)n, (int)m, L.p, Y.p, work, *s, wantRows ? Yn->p : nullptr, Z->p,
                             wantRows ? 1 : 0, wantRows ? 1e-300 : 0.0, tol, 30, c.st);
}

// G = Vh^h diag(lam) Vh (Hermitian PSD, complex128 contiguous [B,n,n])
static int eigh(Ctx& c, const Tn& G, double** lam, Tn* Vh) {
  if (c.npass == 1) {
    // complex64 states: null directions are zeroed downstream anyway (null_tol), so the rank-revealing
    // preconditioned route applies. The two-pass complex128 path needs the complete basis of its first pass.
    const int B = (int)G.sh[0], n = (int)G.sh[1];
    void* scratch = c.ar.raw((size_t)mpdo_eigh_psd_scratch_bytes(B, n));
    *lam = c.ar.reals((long long)B * n);
    *Vh = c.ar.alloc(MPDO_C128, {(long long)B, (long long)n, (long long)n});
    ARENA_OK(c);
    return mpdo_eigh_psd(B, n, G.p, scratch, *lam, Vh->p, c.precondition ? (c.f32 ? 2 : 1) : 0, c.chol_rel,
                         std::max(c.jtol, 4.4e-16 * sqrt((double)n)), 30, c.st);
  }
  Tn none;
  return decompose(c, G, false, lam, &none, Vh);
}

// Pivoted Cholesky G = Lh^h Lh with the left inverse Linv (mpdo_chol_psd). MPDO_ENOSMEM: shape not schedulable,
// the caller uses the eigen route instead.
static int chol(Ctx& c, const Tn& G, Tn* Lh, Tn* Linv) {
  const long long B = G.sh[0], n = G.sh[1];
  void* scratch = c.ar.raw((size_t)mpdo_chol_psd_scratch_bytes((int)B, (int)n));
  *Lh = c.ar.alloc(MPDO_C128, {B, n, n});
  *Linv = c.ar.alloc(MPDO_C128, {B, n, n});
  ARENA_OK(c);
  return mpdo_chol_psd((int)B, (int)n, G.p, scratch, Lh->p, Linv->p, nullptr, c.null_tol, c.st);
}

static int rowscale(Ctx& c, const Tn& V, const double* lam, int lamStride, int rows, double power, double tol, int mode,
                    int dtX, Tn* X) {
  const long long B = V.sh[0], vrows = V.sh[1], cols = V.sh[2];
  *X = c.ar.alloc(dtX, {B, (long long)rows, cols});
  ARENA_OK(c);
  return mpdo_rowscale((int)B, rows, (int)cols, (int)vrows, V.p, lam, lamStride, power, tol, mode, dtX, X->p, c.st);
}

// kept rank (batch maximum) by the reference rule; zero-tails sv in place. SYNC when max_err >= 0.
static int keep_rank(Ctx& c, double* sv, int B, int n, bool squared, int cap, double max_err, bool relative, int* k,
                     int* rows_out = nullptr) {
  if (cap < 0 || cap > n) cap = n;
  if (max_err < 0) {
    *k = cap;
    if (rows_out)
      for (int i = 0; i < B; ++i) rows_out[i] = cap;
    return 0;
  }
  int32_t* dk = (int32_t*)c.ar.raw(sizeof(int32_t) * (size_t)B);
  ARENA_OK(c);
  EC(mpdo_rank_rule(B, n, sv, n, squared, cap, max_err, relative, c.f32, dk, 1, c.st));
  // (pageable on purpose: cudaMallocHost from a strand thread while persistent kernels of other strands are running
  // was measured to stall the caller for 40-150 ms - a single slow step per process - and the read-back is a few bytes)
  std::vector<int32_t> hkv((size_t)B);
  int32_t* hk = hkv.data();
  MPDO_CUDA(cudaMemcpyAsync(hk, dk, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, c.st));
  MPDO_CUDA(cudaStreamSynchronize(c.st));
  int best = 1;
  for (int i = 0; i < B; ++i) best = std::max(best, (int)hk[i]);
  if (rows_out)
    for (int i = 0; i < B; ++i) rows_out[i] = (int)hk[i];
  *k = best;
  return 0;
}

// views with the two matrix axes swapped: [b.. | rows.. | cols..] -> [b.. | cols.. | rows..]
static Tn swap_groups(const Tn& X, Roles r) {
  Tn t = X;
  int pos = r.nb;
  for (int d = 0; d < r.n2; ++d, ++pos) {
    t.sh[pos] = X.sh[r.nb + r.n1 + d];
    t.st[pos] = X.st[r.nb + r.n1 + d];
  }
  for (int d = 0; d < r.n1; ++d, ++pos) {
    t.sh[pos] = X.sh[r.nb + d];
    t.st[pos] = X.st[r.nb + d];
  }
  return t;
}

static Tn like_contig(Ctx& c, const Tn& X, int dt) {  // fresh contiguous tensor with X's logical shape
  Tn t;
  t.dt = dt;
  t.nd = X.nd;
  long long s = 1;
  for (int d = X.nd - 1; d >= 0; --d) {
    t.sh[d] = X.sh[d];
    t.st[d] = s;
    s *= X.sh[d];
  }
  t.p = (char*)c.ar.raw((size_t)s * t.esz());
  return t;
}

// G[b,c,c'] = sum_rows conj(X[b,rows,c]) X[b,rows,c']
static int gram_cols(Ctx& c, const Tn& X, Roles r, Tn* G) {
  const long long B = prod(X, 0, r.nb), n = prod(X, r.nb + r.n1, r.n2);
  *G = c.ar.alloc(MPDO_C128, {B, n, n});
  ARENA_OK(c);
  return contract(c.st, swap_groups(X, r), {r.nb, r.n2, r.n1}, X, r, *G, {1, 1, 1}, true, false, 1, 1.0, 0.0, true);
}

// G[b,i,i'] = sum_cols M[b,i,cols] conj(M[b,i',cols])
static int gram_rows(Ctx& c, const Tn& M, Roles r, Tn* G) {
  const long long B = prod(M, 0, r.nb), n = prod(M, r.nb, r.n1);
  *G = c.ar.alloc(MPDO_C128, {B, n, n});
  ARENA_OK(c);
  return contract(c.st, M, r, swap_groups(M, r), {r.nb, r.n2, r.n1}, *G, {1, 1, 1}, false, true, 1, 1.0, 0.0, true);
}

static Tn transposed(const Tn& X) { return X.permute({0, 2, 1}); }  // [B,a,b] -> [B,b,a] view

// Tall view X -> Q = Alast . Xs^h (isometry, zero columns for null directions), X = Q . R
static int orth_cols(Ctx& c, const Tn& X, Roles r, Tn* Alast, Tn* Xs, Tn* R) {
  Tn G, Vh;
  double* lam;
  EC(gram_cols(c, X, r, &G));
  if (c.npass == 1 && c.use_chol) {   // Cholesky-QR: Q = X . Linv^h, R = Lh (no eigen-decomposition needed)
    const int rc = chol(c, G, R, Xs);
    if (rc == 0) {
      *Alast = X;
      return 0;
    }
    if (rc != MPDO_ENOSMEM) return rc;
  }
  EC(eigh(c, G, &lam, &Vh));
  const int n = (int)G.sh[1];
  if (c.npass == 1) {
    EC(rowscale(c, Vh, lam, n, n, -0.5, c.null_tol, 0, MPDO_C128, Xs));
    EC(rowscale(c, Vh, lam, n, n, 0.5, c.null_tol, 0, MPDO_C128, R));
    *Alast = X;
    return 0;
  }
  Tn Xs1, R1, G2, Vh2, R2;
  double* lam2;
  EC(rowscale(c, Vh, lam, n, n, -0.5, c.floor_tol, 1, MPDO_C128, &Xs1));
  EC(rowscale(c, Vh, lam, n, n, 0.5, c.floor_tol, 1, MPDO_C128, &R1));
  Tn A1 = like_contig(c, X, c.dt);
  ARENA_OK(c);
  EC(contract(c.st, X, r, transposed(Xs1), {1, 1, 1}, A1, r, false, true));
  EC(gram_cols(c, A1, r, &G2));
  EC(eigh(c, G2, &lam2, &Vh2));
  EC(rowscale(c, Vh2, lam2, n, n, -0.5, c.null_tol, 0, MPDO_C128, Xs));
  EC(rowscale(c, Vh2, lam2, n, n, 0.5, c.null_tol, 0, MPDO_C128, &R2));
  *R = c.ar.alloc(MPDO_C128, {G.sh[0], (long long)n, (long long)n});
  ARENA_OK(c);
  EC(contract(c.st, R2, {1, 1, 1}, R1, {1, 1, 1}, *R, {1, 1, 1}));
  *Alast = A1;
  return 0;
}

// Wide view M -> Qt = F . Mlast (orthonormal or zero rows), M = Lh^h . Qt
static int orth_rows(Ctx& c, const Tn& M, Roles r, Tn* Mlast, Tn* F, Tn* Lh, double** lam_out = nullptr,
                     Tn* Uh_out = nullptr) {
  Tn G, Uh;
  double* lam;
  EC(gram_rows(c, M, r, &G));
  if (c.npass == 1 && c.use_chol && !lam_out && !Uh_out) {   // Qt = Linv . M, M = Lh^h . Qt
    const int rc = chol(c, G, Lh, F);
    if (rc == 0) {
      *Mlast = M;
      return 0;
    }
    if (rc != MPDO_ENOSMEM) return rc;
  }
  EC(eigh(c, G, &lam, &Uh));
  const int n = (int)G.sh[1];
  if (lam_out) *lam_out = lam;
  if (Uh_out) *Uh_out = Uh;
  if (c.npass == 1) {
    EC(rowscale(c, Uh, lam, n, n, -0.5, c.null_tol, 0, MPDO_C128, F));
    EC(rowscale(c, Uh, lam, n, n, 0.5, c.null_tol, 0, MPDO_C128, Lh));
    *Mlast = M;
    return 0;
  }
  Tn F1, R1, G2, Uh2, R2;
  double* lam2;
  EC(rowscale(c, Uh, lam, n, n, -0.5, c.floor_tol, 1, MPDO_C128, &F1));
  EC(rowscale(c, Uh, lam, n, n, 0.5, c.floor_tol, 1, MPDO_C128, &R1));
  Tn M1 = like_contig(c, M, c.dt);
  ARENA_OK(c);
  EC(contract(c.st, F1, {1, 1, 1}, M, r, M1, r));
  EC(gram_rows(c, M1, r, &G2));
  EC(eigh(c, G2, &lam2, &Uh2));
  EC(rowscale(c, Uh2, lam2, n, n, -0.5, c.null_tol, 0, MPDO_C128, F));
  EC(rowscale(c, Uh2, lam2, n, n, 0.5, c.null_tol, 0, MPDO_C128, &R2));
  *Lh = c.ar.alloc(MPDO_C128, {G.sh[0], (long long)n, (long long)n});
  ARENA_OK(c);
  EC(contract(c.st, R2, {1, 1, 1}, R1, {1, 1, 1}, *Lh, {1, 1, 1}));
  *Mlast = M1;
  return 0;
}

// Truncatable SVD of a wide view M = U diag(s) Vh: singular values (squared in one-pass mode) and, once the kept
// rank k is known, right(k) with sqrt(S_k) Vh_k = right . Mlast and left(k)[j,i] = sqrt(s_j) conj(U[i,j]).
struct WideSvd {
  Tn Mlast, Uh, Wh, F;  // one pass: Uh (eigenvectors); two passes: Uh = left vectors, Wh, F
  double* sv = nullptr;
  bool squared = false;
  int n = 0;
};

static int svd_wide(Ctx& c, const Tn& M, Roles r, WideSvd* w) {
  if (c.npass == 1) {
    Tn G;
    EC(gram_rows(c, M, r, &G));
    EC(eigh(c, G, &w->sv, &w->Uh));
    w->Mlast = M;
    w->squared = true;
    w->n = (int)G.sh[1];
    return 0;
  }
  Tn Lh, Uh_, Wh_;
  EC(orth_rows(c, M, r, &w->Mlast, &w->F, &Lh));
  EC(decompose(c, Lh, true, &w->sv, &Wh_, &Uh_));  // Lh = Uh_^h diag(s) Wh_  ->  L = Lh^h = Wh_^h diag(s) Uh_
  w->Uh = Wh_;                                     // Uh_L
  w->Wh = Uh_;                                     // Wh_L
  w->squared = false;
  w->n = (int)Lh.sh[1];
  return 0;
}

static int wide_right(Ctx& c, const WideSvd& w, int k, int dt, Tn* out) {
  if (c.npass == 1) return rowscale(c, w.Uh, w.sv, w.n, k, -0.25, c.null_tol, 0, dt, out);
  Tn ws;
  EC(rowscale(c, w.Wh, w.sv, w.n, k, 0.5, 0.0, 0, MPDO_C128, &ws));
  *out = c.ar.alloc(dt, {w.F.sh[0], (long long)k, w.F.sh[2]});
  ARENA_OK(c);
  return contract(c.st, ws, {1, 1, 1}, w.F, {1, 1, 1}, *out, {1, 1, 1});
}

static int wide_left(Ctx& c, const WideSvd& w, int k, int dt, Tn* out) {
  if (c.npass == 1) return rowscale(c, w.Uh, w.sv, w.n, k, 0.25, c.null_tol, 0, dt, out);
  return rowscale(c, w.Uh, w.sv, w.n, k, 0.5, 0.0, 0, dt, out);
}

// Leading k eigenpairs of Hermitian PSD G by block subspace iteration with Rayleigh-Ritz, run to residual
// convergence (steps.py: Engine.eigh_topk). Vt[b,j,:] = components of eigenvector j. converged = 0 -> caller
// falls back to the full decomposition.
static int eigh_topk(Ctx& c, const Tn& G, int k, double** theta_out, Tn* Vt, int* converged) {
  const long long B = G.sh[0], n = G.sh[1];
  const int blk = (int)std::min<long long>(n, std::max(2 * k, k + 28));
  const double tol = c.f32 ? 1e-10 : 1e-12;
  // deterministic start block (a fixed pseudo-random pattern; any generic block works)
  std::vector<double> host((size_t)blk * n * 2);
  unsigned long long sd = 0x9E3779B97F4A7C15ULL;
  for (size_t i = 0; i < host.size(); ++i) {
    sd ^= sd << 13;
    sd ^= sd >> 7;
    sd ^= sd << 17;
    host[i] = ((double)(sd >> 11) / 9007199254740992.0) * 2.0 - 1.0;
  }
  Tn om = c.ar.alloc(MPDO_C128, {1, (long long)blk, n});
  ARENA_OK(c);
  MPDO_CUDA(cudaMemcpyAsync(om.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, c.st));
  MPDO_CUDA(cudaStreamSynchronize(c.st));  // `host` goes out of scope below; the copy is tiny
  Tn Yr = om.expand(0, B);
  Tn Gt = transposed(G);
  Tn Zr = c.ar.alloc(MPDO_C128, {B, (long long)blk, n});
  double* res = c.ar.reals(B);
  ARENA_OK(c);
  EC(contract(c.st, Yr, {1, 1, 1}, Gt, {1, 1, 1}, Zr, {1, 1, 1}));
  std::vector<double> hresv((size_t)B);   // pageable on purpose, see keep_rank
  double* hres = hresv.data();
  *converged = 0;
  // Near-degenerate clusters at the cut (e.g. the equal-weight error branches of a chi-matrix gate) make the
  // iteration stall; the rank-revealing full decomposition is cheap enough that a short leash is the better policy.
  const int maxIter = c.f32 ? 6 : 12;
  double prevWorst = 1e300;
  for (int it = 0; it < maxIter; ++it) {
    Tn H, Uh, Fo, Bm, Wh, ts;
    double *lam, *theta;
    EC(gram_rows(c, Zr, {1, 1, 1}, &H));
    EC(eigh(c, H, &lam, &Uh));
    EC(rowscale(c, Uh, lam, blk, blk, -0.5, c.null_tol, 0, MPDO_C128, &Fo));
    Tn Y2 = c.ar.alloc(MPDO_C128, {B, (long long)blk, n});
    Tn Z2 = c.ar.alloc(MPDO_C128, {B, (long long)blk, n});
    Bm = c.ar.alloc(MPDO_C128, {B, (long long)blk, (long long)blk});
    ARENA_OK(c);
    EC(contract(c.st, Fo, {1, 1, 1}, Zr, {1, 1, 1}, Y2, {1, 1, 1}));
    EC(contract(c.st, Y2, {1, 1, 1}, Gt, {1, 1, 1}, Z2, {1, 1, 1}));
    EC(contract(c.st, Z2, {1, 1, 1}, transposed(Y2), {1, 1, 1}, Bm, {1, 1, 1}, false, true));
    EC(eigh(c, Bm, &theta, &Wh));
    Tn Y3 = c.ar.alloc(MPDO_C128, {B, (long long)blk, n});
    Tn Z3 = c.ar.alloc(MPDO_C128, {B, (long long)blk, n});
    ARENA_OK(c);
    EC(contract(c.st, Wh, {1, 1, 1}, Y2, {1, 1, 1}, Y3, {1, 1, 1}));
    EC(contract(c.st, Wh, {1, 1, 1}, Z2, {1, 1, 1}, Z3, {1, 1, 1}));
    // residual rows R = Z3[:k] - theta * Y3[:k]  (scale Wh[:k] by theta, multiply by Y2, subtract from Wh[:k] Z2)
    Tn Rr = c.ar.alloc(MPDO_C128, {B, (long long)k, n});
    ARENA_OK(c);
    EC(rowscale(c, Wh, theta, blk, k, 1.0, 0.0, 0, MPDO_C128, &ts));
    EC(contract(c.st, Wh.narrow(1, 0, k), {1, 1, 1}, Z2, {1, 1, 1}, Rr, {1, 1, 1}));
    EC(contract(c.st, ts, {1, 1, 1}, Y2, {1, 1, 1}, Rr, {1, 1, 1}, false, false, -1, -1.0, 1.0));
    residual_kernel<<<(unsigned)B, 256, 0, c.st>>>(k, (int)n, (const double2*)Rr.p, theta, blk, res);
    EC(check_launch("residual_kernel"));
    MPDO_CUDA(cudaMemcpyAsync(hres, res, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, c.st));
    MPDO_CUDA(cudaStreamSynchronize(c.st));
    double worst = 0;
    for (long long i = 0; i < B; ++i) worst = std::max(worst, hres[i]);
    Yr = Y3;
    Zr = Z3;
    // stalled (less than 20x per iteration): the cut sits inside a cluster wider than the block; stop paying for it
    if (it >= 1 && worst > tol && worst > 0.05 * prevWorst) return 0;
    prevWorst = worst;
    if (worst <= tol) {
      static const bool trace = getenv("MPDO_TRACE") != nullptr;
      if (trace) fprintf(stderr, "[mpdo] eigh_topk n=%lld k=%d blk=%d B=%lld iterations=%d\n", n, k, blk, B, it + 1);
      *converged = 1;
      *theta_out = theta;   // stride blk
      *Vt = Y3;             // [B, blk, n]; the first k rows are the kept vectors
      return 0;
    }
  }
  return 0;
}

}  // namespace eng
}  // namespace mpdo

using namespace mpdo;
using namespace mpdo::eng;

// ===================================================================================================
// C entry points
// ===================================================================================================

extern "C" int mpdo_trim_pools(void) {
  trim_all_pools();
  return 0;
}

extern "C" int mpdo_pool_stats(int64_t* reservedBytes, int64_t* reservedHighBytes) {
  cudaMemPool_t pool = scratch_pool();
  if (!pool) return fail(MPDO_EINVAL, "mpdo_pool_stats: no scratch pool on this device");
  unsigned long long cur = 0, high = 0;
  MPDO_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &cur));
  MPDO_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &high));
  if (reservedBytes) *reservedBytes = (int64_t)cur;
  if (reservedHighBytes) *reservedHighBytes = (int64_t)high;
  return 0;
}

extern "C" int mpdo_qr_step(int dtype, int npass, int B, int l, int a, int r, const void* Ti, int a2, int r2,
                            const void* Tn_in, void* Q_out, void* Tn_out, void* stream) {
  Ctx c((cudaStream_t)stream, dtype, npass);
  c.set_batch(B);
  Tn T = Tn::contig((void*)Ti, dtype, {B, l, 2, a, r});
  Tn Tnx = Tn::contig((void*)Tn_in, dtype, {B, r, 2, a2, r2});
  Tn Q = Tn::contig(Q_out, dtype, {B, l, 2, a, r});
  Tn To = Tn::contig(Tn_out, dtype, {B, r, 2, a2, r2});
  Tn Alast, Xs, R, Xsd, Rd;
  EC(orth_cols(c, T, {1, 3, 1}, &Alast, &Xs, &R));
  Xsd = c.ar.alloc(dtype, {B, r, r});
  Rd = c.ar.alloc(dtype, {B, r, r});
  ARENA_OK(c);
  EC(copy_view(c.st, Xs, Xsd));
  EC(copy_view(c.st, R, Rd));
  EC(contract(c.st, Alast, {1, 3, 1}, transposed(Xsd), {1, 1, 1}, Q, {1, 3, 1}, false, true));
  return contract(c.st, Rd, {1, 1, 1}, Tnx, {1, 1, 3}, To, {1, 1, 3});
}

// chi truncation step: k = min(chi, l), further reduced by the relative-error rule when max_err >= 0 (SYNC: the kept
// rank is read back; outputs are written densely with the kept rank into the caller's cap-sized buffers).
// Tl [B,lp,2,ap,l], Tr [B,l,2,a,r].
extern "C" int mpdo_bond_svd_step(int dtype, int npass, int B, int lp, int ap, int l, const void* Tl, int a, int r,
                                  const void* Tr, int k, double max_err, int* k_out, void* Tl_out, void* Tr_out,
                                  double* disc_out, void* stream) {
  Ctx c((cudaStream_t)stream, dtype, npass);
  c.set_batch(B);
  Tn TL = Tn::contig((void*)Tl, dtype, {B, lp, 2, ap, l});
  Tn TR = Tn::contig((void*)Tr, dtype, {B, l, 2, a, r});
  WideSvd w;
  EC(svd_wide(c, TR, {1, 1, 3}, &w));
  if (disc_out)  // singular values (squared in one-pass mode), all l of them, for the caller's truncation record
    MPDO_CUDA(cudaMemcpyAsync(disc_out, w.sv, sizeof(double) * (size_t)B * w.n, cudaMemcpyDeviceToDevice, c.st));
  if (max_err >= 0) EC(keep_rank(c, w.sv, B, w.n, w.squared, k, max_err, true, &k));   // decompositions.py:117-134
  if (k_out) *k_out = k;
  Tn TLo = Tn::contig(Tl_out, dtype, {B, lp, 2, ap, k});
  Tn TRo = Tn::contig(Tr_out, dtype, {B, k, 2, a, r});
  Tn right, left;
  EC(wide_right(c, w, k, dtype, &right));
  EC(contract(c.st, right, {1, 1, 1}, w.Mlast, {1, 1, 3}, TRo, {1, 1, 3}));
  EC(wide_left(c, w, k, dtype, &left));
  return contract(c.st, TL, {1, 3, 1}, transposed(left), {1, 1, 1}, TLo, {1, 3, 1}, false, true);
}

// ---------------------------------------------------------------------------------------------------
// bondTruncate through left environments (complex64 states, fixed chi): see MPDOSimulator/_engine/steps.py
// bond_truncate_env for the mathematics. The left-to-right pass is a chain of contractions (E_{i+1} = T_i^h E_i T_i,
// the very Gram matrix the QR sweep forms); the factorisations E_i = C_i^h C_i do not depend on one another and run on
// side streams while the chain continues; the right-to-left pass (mpdo_bond_env_step) is the one sequential chain of
// decompositions left.
// ---------------------------------------------------------------------------------------------------
constexpr int N_SIDE = 8;
static std::mutex g_side_mu;
static cudaStream_t g_side[64][N_SIDE] = {{nullptr}};

static cudaStream_t side_stream(int i) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_side_mu);
  cudaStream_t& s = g_side[dev][i % N_SIDE];
  if (!s && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    s = nullptr;
  }
  return s;
}

// E = C^h C with the left inverse Ci (written to the caller's buffer) and the product M = C . Tnext of the factor with
// the site right of the bond, all on stream `on`
static int env_factor(int dtype, int B, int n, const Tn& E, const Tn& Tnext, void* Ci_out, void* M_out,
                      cudaStream_t on) {
  Ctx cs(on, dtype, 1);
  cs.set_batch(B);
  Tn Cf = cs.ar.alloc(MPDO_C128, {(long long)B, (long long)n, (long long)n});
  Tn Cd = cs.ar.alloc(dtype, {(long long)B, (long long)n, (long long)n});
  ARENA_OK(cs);
  int rc = MPDO_ENOSMEM;
  if (cs.use_chol) {
    void* scratch = cs.ar.raw((size_t)mpdo_chol_psd_scratch_bytes(B, n));
    ARENA_OK(cs);
    rc = mpdo_chol_psd(B, n, E.p, scratch, Cf.p, Ci_out, nullptr, cs.null_tol, on);
    if (rc != 0 && rc != MPDO_ENOSMEM) return rc;
  }
  if (rc == MPDO_ENOSMEM) {   // shape not schedulable for the Cholesky kernels: eigen route
    double* lam;
    Tn Vh, Cis;
    EC(eigh(cs, E, &lam, &Vh));
    EC(rowscale(cs, Vh, lam, n, n, 0.5, cs.null_tol, 0, MPDO_C128, &Cf));
    EC(rowscale(cs, Vh, lam, n, n, -0.5, cs.null_tol, 0, MPDO_C128, &Cis));
    MPDO_CUDA(cudaMemcpyAsync(Ci_out, Cis.p, sizeof(double2) * (size_t)B * n * n, cudaMemcpyDeviceToDevice, on));
  }
  EC(copy_view(on, Cf, Cd));
  Tn Mo = Tn::contig(M_out, dtype, {Tnext.sh[0], Tnext.sh[1], Tnext.sh[2], Tnext.sh[3], Tnext.sh[4]});
  return contract(on, Cd, {1, 1, 1}, Tnext, {1, 1, 3}, Mo, {1, 1, 3});
}

extern "C" int mpdo_env_sweep(int dtype, int B, int nsites, const int* l, const int* a, const int* r,
                              const void* const* T, void* const* Ci_out, void* const* M_out, void* stream) {
  if (nsites < 2) return 0;
  if (!l || !a || !r || !T || !Ci_out || !M_out) return fail(MPDO_EINVAL, "mpdo_env_sweep: null argument");
  if (l[0] != 1) return fail(MPDO_EINVAL, "mpdo_env_sweep: the first site must have a trivial left bond");
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c(st, dtype, 1);
  c.set_batch(B);
  static const bool serial = getenv("MPDO_ENV_SERIAL") != nullptr;   // A/B knob: factorisations on the caller's stream
  bool used[N_SIDE] = {false};
  Tn E;
  int rc = 0;
  for (int i = 0; i + 1 < nsites && rc == 0; ++i) {
    if (i > 0 && l[i] != r[i - 1]) {
      rc = fail(MPDO_EINVAL, "mpdo_env_sweep: bond dimensions of neighbouring sites differ");
      break;
    }
    Tn Ti = Tn::contig((void*)T[i], dtype, {(long long)B, (long long)l[i], 2, (long long)a[i], (long long)r[i]});
    Tn En;
    if (i == 0) {
      rc = gram_cols(c, Ti, {1, 3, 1}, &En);   // E_0 = 1
    } else {
      Ctx cx(st, dtype, 1);   // the wide intermediate is released as soon as the two contractions are queued
      Tn X = cx.ar.alloc(MPDO_C128, {(long long)B, (long long)l[i], 2, (long long)a[i], (long long)r[i]});
      En = c.ar.alloc(MPDO_C128, {(long long)B, (long long)r[i], (long long)r[i]});
      if (cx.ar.err || c.ar.err) {
        rc = cx.ar.err ? cx.ar.err : c.ar.err;
        break;
      }
      rc = contract(st, E, {1, 1, 1}, Ti, {1, 1, 3}, X, {1, 1, 3});
      if (rc == 0) rc = contract(st, Ti.permute({0, 4, 1, 2, 3}), {1, 1, 3}, X, {1, 3, 1}, En, {1, 1, 1}, true, false);
    }
    if (rc) break;
    E = En;
    // cooperative factorisations (n > 256: device-wide barrier) keep to the caller's stream
    cudaStream_t side = (serial || r[i] > 256) ? nullptr : side_stream(i);
    if (side) {
      cudaEvent_t ev;
      MPDO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      cudaEventRecord(ev, st);
      cudaStreamWaitEvent(side, ev, 0);
      cudaEventDestroy(ev);
      used[i % N_SIDE] = true;
    }
    if (l[i + 1] != r[i]) {
      rc = fail(MPDO_EINVAL, "mpdo_env_sweep: bond dimensions of neighbouring sites differ");
      break;
    }
    Tn Tnext = Tn::contig((void*)T[i + 1], dtype,
                          {(long long)B, (long long)l[i + 1], 2, (long long)a[i + 1], (long long)r[i + 1]});
    rc = env_factor(dtype, B, r[i], E, Tnext, Ci_out[i + 1], M_out[i + 1], side ? side : st);
  }
  for (int s = 0; s < N_SIDE; ++s) {   // join (also on failure: the arena of `c` is freed on `st` after this)
    if (!used[s]) continue;
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      cudaStreamSynchronize(side_stream(s));
      continue;
    }
    cudaEventRecord(ev, side_stream(s));
    cudaStreamWaitEvent(st, ev, 0);
    cudaEventDestroy(ev);
  }
  return rc;
}

// One step of the right-to-left chi truncation on the un-canonicalised state: M0 [B,l,2,a,r0] = C . T (the site right
// of the bond times the factor of the bond's environment, from mpdo_env_sweep), W [B,r0,rw] (state dtype; NULL at the
// last site, then rw = r0), Ci [B,l,l] complex128.
// M = M0 . W = U S V^h  ->  T_out = sqrt(S_k) V_k^h [B,k,2,a,rw], W_out = Ci^h U_k sqrt(S_k) [B,l,k],
// sv_out (optional) the l squared singular values.
extern "C" int mpdo_bond_env_step(int dtype, int B, int l, int a, int r0, const void* M0, int rw, const void* W,
                                  const void* Ci, int k, void* T_out, void* W_out, double* sv_out, void* stream) {
  Ctx c((cudaStream_t)stream, dtype, 1);
  c.set_batch(B);
  if (k < 1 || k > l) return fail(MPDO_EINVAL, "mpdo_bond_env_step: kept rank out of range");
  const long long Bn = B;
  Tn M = Tn::contig((void*)M0, dtype, {Bn, l, 2, a, r0});
  if (W) {
    Tn MW = c.ar.alloc(dtype, {Bn, (long long)l, 2, (long long)a, (long long)rw});
    ARENA_OK(c);
    EC(contract(c.st, M, {1, 3, 1}, Tn::contig((void*)W, dtype, {Bn, r0, rw}), {1, 1, 1}, MW, {1, 3, 1}));
    M = MW;
  } else {
    rw = r0;
  }
  WideSvd w;
  EC(svd_wide(c, M, {1, 1, 3}, &w));
  if (sv_out)
    MPDO_CUDA(cudaMemcpyAsync(sv_out, w.sv, sizeof(double) * (size_t)B * w.n, cudaMemcpyDeviceToDevice, c.st));
  Tn right, UL;
  EC(wide_right(c, w, k, dtype, &right));
  EC(contract(c.st, right, {1, 1, 1}, M, {1, 1, 3}, Tn::contig(T_out, dtype, {Bn, k, 2, a, rw}), {1, 1, 3}));
  EC(wide_left(c, w, k, MPDO_C128, &UL));   // UL[j,i] = sqrt(s_j) conj(U[i,j])
  return contract(c.st, transposed(Tn::contig((void*)Ci, MPDO_C128, {Bn, l, l})), {1, 1, 1}, transposed(UL),
                  {1, 1, 1}, Tn::contig(W_out, dtype, {Bn, l, k}), {1, 1, 1}, true, true);
}

// kappa truncation: k = min(kappa, a), further reduced by the relative-error rule when max_err >= 0 (SYNC; the output
// is written densely with the kept rank). disc_out[b] = norm of the discarded part.
extern "C" int mpdo_kappa_truncate(int dtype, int B, int l, int a, int r, const void* T, int k, double max_err,
                                   int* k_out, void* T_out, double* disc_out, void* stream) {
  Ctx c((cudaStream_t)stream, dtype, 1);
  c.set_batch(B);
  Tn Tv = Tn::contig((void*)T, dtype, {B, l, 2, a, r});
  Tn G;
  EC(gram_cols(c, Tv.permute({0, 1, 2, 4, 3}), {1, 3, 1}, &G));
  Tn Tperm = Tv.permute({0, 3, 1, 2, 4});   // [b | a | l,s,r]
  double* theta = nullptr;
  int thetaStride = a;
  bool done = false;
  // The subspace iteration pays off when the spectrum has a gap after the kept values; on workloads where the cut sits
  // in a cluster (the equal-weight error branches of a chi-matrix gate) it stalls every time and costs two wasted
  // iterations (~2.2 ms of 2.7 at a = 64 on the headline workload). Remember the outcome per (a, k): after a stall the
  // next 63 calls with that signature go straight to the full rank-revealing decomposition and ONE caller probes again
  // (the sites of a layer arrive here from a dozen strand threads at once: letting every one of them probe cost 7 ms
  // per layer); further stalls lengthen the pause (63, 4095, 65535 calls), a converged probe resets it. A probe on a
  // cold process was also measured to cost far more than its kernels (25-110 ms once, first process on a machine).
  // Both routes are exact solvers run to convergence, so this only moves time.
  static std::atomic<int> topkSkip[64], topkStalls[64];
  const unsigned slot = (((unsigned)a * 31u + (unsigned)k) * 17u + (unsigned)std::min(B, 64)) & 63u;
  std::atomic<int>& skip = topkSkip[slot];
  std::atomic<int>& stalls = topkStalls[slot];
  static const bool noTopk = getenv("MPDO_NO_TOPK") != nullptr;   // A/B knob: always the full decomposition
  // (for one or a few circuits the full decomposition of a Gram matrix below order 256 is a fraction of a millisecond
  // and the iteration cannot beat it; a probe that lands in a steady-state step only costs - measured 15-95 ms with
  // its first-time scratch shapes)
  bool tryTopk = !noTopk && max_err < 0 && a >= 64 && a >= 8 * k && (a >= 256 || B >= 16);
  bool claimed = false;
  if (tryTopk && stalls.load(std::memory_order_relaxed) > 0) {   // this signature stalled before: paused / one prober
    int cur = skip.load(std::memory_order_relaxed);
    for (;;) {
      if (cur > 0) {            // paused: one call less to wait
        if (skip.compare_exchange_weak(cur, cur - 1, std::memory_order_relaxed)) {
          tryTopk = false;
          break;
        }
      } else if (cur == 0) {    // claim the probe (-1 while it runs)
        if (skip.compare_exchange_weak(cur, -1, std::memory_order_relaxed)) {
          claimed = true;
          break;
        }
      } else {                  // another thread is probing right now
        tryTopk = false;
        break;
      }
    }
  }
  if (tryTopk) {
    Tn Vt;
    int conv = 0;
    const int rcTop = eigh_topk(c, G, k, &theta, &Vt, &conv);
    if (rcTop != 0) {
      if (claimed) skip.store(0, std::memory_order_relaxed);
      return rcTop;
    }
    if (conv) {
      if (claimed) {   // the signature converges again: everybody may use the iteration
        stalls.store(0, std::memory_order_relaxed);
        skip.store(0, std::memory_order_relaxed);
      }
    } else if (claimed) {
      const int n = stalls.fetch_add(1, std::memory_order_relaxed);
      skip.store(n <= 1 ? 63 : (n == 2 ? 4095 : 65535), std::memory_order_relaxed);
    } else {           // first stall of a signature that converged so far (several callers may get here at once)
      int expected = 0;
      if (stalls.compare_exchange_strong(expected, 1, std::memory_order_relaxed)) skip.store(63, std::memory_order_relaxed);
    }
    if (conv) {
      thetaStride = (int)Vt.sh[1];
      Tn Vk = c.ar.alloc(dtype, {(long long)B, (long long)k, (long long)a});
      ARENA_OK(c);
      EC(copy_view(c.st, Vt.narrow(1, 0, k), Vk));
      Tn Operm = Tn::contig(T_out, dtype, {B, l, 2, k, r}).permute({0, 3, 1, 2, 4});   // [b | k | l,s,r]
      EC(contract(c.st, Vk, {1, 1, 1}, Tperm, {1, 1, 3}, Operm, {1, 1, 3}));
      done = true;
    }
  }
  if (!done) {
    Tn Vh;
    EC(eigh(c, G, &theta, &Vh));
    thetaStride = a;
    if (max_err >= 0) {   // the eigenvalues are the squared singular values of T over the inner index
      double* sv = c.ar.reals((long long)B * a);
      ARENA_OK(c);
      MPDO_CUDA(cudaMemcpyAsync(sv, theta, sizeof(double) * (size_t)B * a, cudaMemcpyDeviceToDevice, c.st));
      EC(keep_rank(c, sv, B, a, true, k, max_err, true, &k));
    }
    Tn Vk = c.ar.alloc(dtype, {(long long)B, (long long)k, (long long)a});
    ARENA_OK(c);
    EC(copy_view(c.st, Vh.narrow(1, 0, k), Vk));
    Tn Operm = Tn::contig(T_out, dtype, {B, l, 2, k, r}).permute({0, 3, 1, 2, 4});
    EC(contract(c.st, Vk, {1, 1, 1}, Tperm, {1, 1, 3}, Operm, {1, 1, 3}, true, false));
  }
  if (k_out) *k_out = k;
  if (disc_out) {
    discarded_kernel<<<(B + 127) / 128, 128, 0, c.st>>>(B, a, k, (const double2*)G.p, theta, thetaStride, disc_out);
    EC(check_launch("discarded_kernel"));
  }
  return 0;
}

// Two-qubit gate absorption + split (Circuit.py:74-136). G [Bg,2,2,2,2,K] in (lo, hi) order, same dtype as the state.
// The kept rank is data dependent: alloc(which, count, user) is called once the rank is known and must return device
// memory for count complex elements (which = 0: T_lo' [B,l,2,a0,k]; 1: T_hi' [B,k,2,K*a1,r]). SYNC (rank read-back).
// ranks_out (optional, host memory, B ints): the kept rank of every batch entry (k_out is their maximum).
extern "C" int mpdo_split_2q(int dtype, int npass, int B, int l, int a0, int m, const void* Tlo, int a1, int r,
                             const void* Thi, int Bg, int K, const void* G, double max_err, mpdo_alloc_fn alloc,
                             void* user, int* k_out, int* ranks_out, void* stream) {
  Ctx c((cudaStream_t)stream, dtype, npass);
  c.set_batch(B);
  const long long Bn = B;
  Tn TLO = Tn::contig((void*)Tlo, dtype, {Bn, l, 2, a0, m});
  Tn THI = Tn::contig((void*)Thi, dtype, {Bn, m, 2, a1, r});
  Tn Gv = Tn::contig((void*)G, dtype, {Bg, 2, 2, 2, 2, K});

  // left factor: rows (l,a0), cols (s0,m)
  Tn Xlo = TLO.permute({0, 1, 3, 2, 4});  // [B,l,a0,2,m]
  Tn Alo, Xs_lo, Rp;
  bool orthLo = (long long)l * a0 > 2LL * m;
  long long x;
  if (orthLo) {
    EC(orth_cols(c, Xlo, {1, 2, 2}, &Alo, &Xs_lo, &Rp));
    x = 2LL * m;
  } else {
    x = (long long)l * a0;
    Rp = c.ar.alloc(MPDO_C128, {Bn, x, 2LL * m});
    ARENA_OK(c);
    EC(copy_view(c.st, Xlo, Rp));
  }
  // right factor: rows (m,s1), cols (a1,r)
  Tn Mhi, F_hi, Lh_hi;
  bool orthHi = (long long)a1 * r > 2LL * m;
  long long y;
  static const bool oldSplit = getenv("MPDO_SPLIT_VIA_CORE") != nullptr;   // A/B knob: always form the core
  if (c.npass == 1 && orthHi && !oldSplit) {
    // ---- complex64 states, wide right factor: the core is never formed ------------------------------------------
    // Theta = (Q' x 1) . C . (1 x Qt') and the split only needs (i) the Gram matrix of the core over its columns and
    // (ii) sqrt(S) Vh = S^-1/2 U^h Theta. With Lam = T_hi T_hi^h over (a1, r) [the 2m x 2m Gram matrix of the right
    // site] and Gam[p0 s0 s1 ; p0' s0' s1'] = sum_{p1,g} G conj(G) [8 x 8, gate only]:
    //   C C^h [(x,p0),(x',p0')] = sum R'[x,s0,m] conj(R'[x',s0',m']) Lam[(m,s1),(m',s1')] Gam[p0 s0 s1 ; p0' s0' s1']
    // which is ~0.2 GFLOP of small contractions instead of the 2x x 2Ky core (268 MB for a fused rzz at chi = 64),
    // its 17 GFLOP fp64 Gram matrix and the 25 + 13 GFLOP of Zc = right . C and Zc . F_hi; no factorisation of the
    // right site is needed at all. The kept right factor is rebuilt from the site itself:
    //   T_hi'[j,p1,(g,a1),r] = sum_{m,s1} Y[j,p1,g,m,s1] T_hi[m,s1,a1,r],  Y = sum_{x,p0,s0} W[j,x,p0] R'[x,s0,m] G[p0,p1,s0,s1,g]
    // with W = S^-1/4 U^h. Same sqrt(S) | sqrt(S) split, same rank rule on the same singular values.
    Tn Lam;
    EC(gram_rows(c, THI, {1, 2, 2}, &Lam));                                  // [B, (m,s1), (m',s1')]
    const long long mm = m;
    Tn T1 = c.ar.alloc(MPDO_C128, {Bn, 2, x * 2, 2 * mm});                   // [b, s1, (x,s0), (m',s1')]
    ARENA_OK(c);
    EC(contract(c.st, Rp.view({Bn, 1, x * 2, mm}).expand(1, 2), {2, 1, 1},
                Lam.view({Bn, mm, 2, 2 * mm}).permute({0, 2, 1, 3}), {2, 1, 1}, T1, {2, 1, 1}));
    Tn T2s = c.ar.alloc(MPDO_C128, {Bn, x, x, 2, 2, 2, 2});                  // [b, x, x', s0, s1, s0', s1']
    ARENA_OK(c);
    EC(contract(c.st, T1.view({Bn, 2, x, 2, mm, 2}).permute({0, 5, 1, 2, 3, 4}), {2, 3, 1},
                Rp.view({Bn, 1, x, 2, mm}).expand(1, 2).permute({0, 1, 4, 2, 3}), {2, 1, 2},
                T2s.permute({0, 6, 4, 1, 3, 2, 5}), {2, 3, 2}, false, true));
    Tn Gc = c.ar.alloc(MPDO_C128, {(long long)Bg, 2, 2, 2, 2, (long long)K});
    Tn Gam = c.ar.alloc(MPDO_C128, {(long long)Bg, 8, 8});                   // [bg, (p0,s0,s1), (p0',s0',s1')]
    ARENA_OK(c);
    EC(copy_view(c.st, Gv, Gc));
    EC(contract(c.st, Gc.permute({0, 1, 3, 4, 2, 5}), {1, 3, 2}, Gc.permute({0, 2, 5, 1, 3, 4}), {1, 2, 3}, Gam, {1, 1, 1},
                false, true));
    Tn GG = c.ar.alloc(MPDO_C128, {Bn, x, 2, x, 2});                         // [b, (x,p0), (x',p0')]
    ARENA_OK(c);
    {
      Tn GamV = Gam.view({(long long)Bg, 2, 2, 2, 2, 2, 2}).permute({0, 1, 4, 2, 3, 5, 6});   // [bg | p0,p0' | s0,s1,s0',s1']
      if (Bg == 1) GamV = GamV.expand(0, Bn);
      EC(contract(c.st, GamV, {1, 2, 4}, T2s.view({Bn, x * x, 16}).permute({0, 2, 1}), {1, 1, 1},
                  GG.permute({0, 2, 4, 1, 3}), {1, 2, 2}));
    }
    double* lam;
    Tn Uh;
    EC(eigh(c, GG.view({Bn, 2 * x, 2 * x}), &lam, &Uh));
    int k = 0;
    EC(keep_rank(c, lam, B, (int)(2 * x), true, -1, max_err, false, &k, ranks_out));
    Tn UL, Wr;
    EC(rowscale(c, Uh, lam, (int)(2 * x), k, 0.25, c.null_tol, 0, MPDO_C128, &UL));     // sqrt(s_j) conj(U[(x,p0), j])
    EC(rowscale(c, Uh, lam, (int)(2 * x), k, -0.25, c.null_tol, 0, MPDO_C128, &Wr));
    Tn V = c.ar.alloc(MPDO_C128, {Bn, (long long)k, 2, 2, mm});               // [b, j, p0, s0, m]
    Tn Gq = c.ar.alloc(MPDO_C128, {(long long)Bg, 2, (long long)K, 2, 2, 2});  // [bg, p1, g, s1, p0, s0]
    Tn Y = c.ar.alloc(dtype, {Bn, (long long)k, 2, (long long)K, mm, 2});      // [b, j, p1, g, m, s1]
    ARENA_OK(c);
    EC(contract(c.st, Wr.view({Bn, (long long)k, x, 2}).permute({0, 3, 1, 2}), {2, 1, 1},
                Rp.view({Bn, 1, x, 2 * mm}).expand(1, 2), {2, 1, 1}, V.permute({0, 2, 1, 3, 4}), {2, 1, 2}));
    EC(copy_view(c.st, Gv.permute({0, 2, 5, 4, 1, 3}), Gq));
    {
      Tn GqV = Gq.view({(long long)Bg, 1, 4LL * K, 4});
      if (Bg == 1) GqV = GqV.expand(0, Bn);
      GqV = GqV.expand(1, k);
      EC(contract(c.st, GqV, {2, 1, 1}, V.view({Bn, (long long)k, 4, mm}), {2, 1, 1}, Y.permute({0, 1, 2, 3, 5, 4}),
                  {2, 3, 1}));
    }
    *k_out = k;
    void* plo = alloc(0, (int64_t)Bn * l * 2 * a0 * k, user);
    void* phi = alloc(1, (int64_t)Bn * k * 2 * K * a1 * r, user);
    if (!plo || !phi) return fail(MPDO_EINVAL, "mpdo_split_2q: output allocation failed");
    Tn Tlo_n = Tn::contig(plo, dtype, {Bn, l, 2, a0, (long long)k});
    Tn Thi_n = Tn::contig(phi, dtype, {Bn, (long long)k, 2, (long long)K * a1, r});
    Tn ULv = UL.view({Bn, (long long)k, x, 2});
    if (orthLo) {
      Tn W = c.ar.alloc(dtype, {Bn, 2LL * m, 2, (long long)k});
      ARENA_OK(c);
      EC(contract(c.st, transposed(Xs_lo), {1, 1, 1}, ULv.permute({0, 2, 3, 1}), {1, 1, 2}, W.view({Bn, 2LL * m, 2LL * k}),
                  {1, 1, 1}, true, true));
      EC(contract(c.st, Alo, {1, 2, 2}, W.view({Bn, 2, (long long)m, 2, (long long)k}), {1, 2, 2},
                  Tlo_n.permute({0, 1, 3, 2, 4}), {1, 2, 2}));
    } else {
      Tn src = UL.view({Bn, (long long)k, (long long)l, (long long)a0, 2}).permute({0, 2, 4, 3, 1});
      EC(copy_view(c.st, src, Tlo_n, true));
    }
    // the one big product of the split: [(j,p1,g) x (m,s1)] . [(m,s1) x (a1,r)], complex64 (tensor-core tile)
    return contract(c.st, Y.view({Bn, (long long)k * 2 * K, mm, 2}), {1, 1, 2}, THI, {1, 2, 2},
                    Thi_n.view({Bn, (long long)k * 2 * K, (long long)a1, (long long)r}), {1, 1, 2});
  }
  if (orthHi) {
    EC(orth_rows(c, THI, {1, 2, 2}, &Mhi, &F_hi, &Lh_hi));
    y = 2LL * m;
  } else {
    y = (long long)a1 * r;
  }
  // D[b,x,s0,s1,y] = sum_m R'[x,s0,m] L'[m,s1,y]
  Tn D = c.ar.alloc(MPDO_C128, {Bn, x, 2, 2, y});
  ARENA_OK(c);
  {
    Tn Rp5 = Rp.view({Bn, x, 2, 1, (long long)m}).expand(3, 2).permute({0, 2, 3, 1, 4});  // [b,s0,s1 | x | m]
    Tn Dv = D.permute({0, 2, 3, 1, 4});                                                      // [b,s0,s1 | x | y]
    if (orthHi) {
      // L'[(m,s1),y] = conj(Lh_hi[y,(m,s1)])
      Tn Lv = Lh_hi.view({Bn, y, (long long)m, 1, 2}).expand(3, 2).permute({0, 3, 4, 2, 1});  // [b,s0,s1 | m | y]
      EC(contract(c.st, Rp5, {3, 1, 1}, Lv, {3, 1, 1}, Dv, {3, 1, 1}, false, true));
    } else {
      Tn Lv = THI.view({Bn, (long long)m, 1, 2, y}).expand(2, 2).permute({0, 2, 3, 1, 4});     // [b,s0,s1 | m | y]
      EC(contract(c.st, Rp5, {3, 1, 1}, Lv, {3, 1, 1}, Dv, {3, 1, 1}));
    }
  }
  // Cm[b,x,p0,p1,g,y] = sum_{s0,s1} G[p0,p1,s0,s1,g] D[b,x,s0,s1,y]
  Tn Gp = c.ar.alloc(MPDO_C128, {(long long)Bg, 2, 2, (long long)K, 2, 2});
  Tn Cm = c.ar.alloc(MPDO_C128, {Bn, x, 2, 2, (long long)K, y});
  ARENA_OK(c);
  EC(copy_view(c.st, Gv.permute({0, 1, 2, 5, 3, 4}), Gp));
  {
    Tn GpE = Gp.view({(long long)Bg, 1, 4LL * K, 4});
    if (Bg == 1) GpE = GpE.expand(0, Bn);
    GpE = GpE.expand(1, x);
    EC(contract(c.st, GpE, {2, 1, 1}, D.view({Bn, x, 4, y}), {2, 1, 1}, Cm.view({Bn, x, 4LL * K, y}), {2, 1, 1}));
  }
  // SVD of the core as rows (x,p0) x cols (p1,g,y)
  const long long nrow = 2 * x, ncol = 2LL * K * y;
  Tn Cv = Cm.view({Bn, nrow, ncol});
  const int zdt = orthHi ? MPDO_C128 : dtype;
  Tn UL, Zc;
  int k = 0;
  if (nrow <= ncol) {
    WideSvd w;
    EC(svd_wide(c, Cv, {1, 1, 1}, &w));
    EC(keep_rank(c, w.sv, B, w.n, w.squared, -1, max_err, false, &k, ranks_out));
    EC(wide_left(c, w, k, MPDO_C128, &UL));
    Tn right;
    EC(wide_right(c, w, k, MPDO_C128, &right));
    Zc = c.ar.alloc(zdt, {Bn, (long long)k, ncol});
    ARENA_OK(c);
    EC(contract(c.st, right, {1, 1, 1}, w.Mlast, {1, 1, 1}, Zc, {1, 1, 1}));
  } else {
    // tall core: decompose the column side
    Tn Alast, Xs, R, Uh_, Wh_, su, XU;
    double* s;
    EC(orth_cols(c, Cv, {1, 1, 1}, &Alast, &Xs, &R));
    EC(decompose(c, R, true, &s, &Wh_, &Uh_));  // R = Uh_^h diag(s) Wh_
    EC(keep_rank(c, s, B, (int)ncol, false, -1, max_err, false, &k, ranks_out));
    EC(rowscale(c, Uh_, s, (int)ncol, k, 0.5, 0.0, 0, MPDO_C128, &su));
    XU = c.ar.alloc(MPDO_C128, {Bn, (long long)k, ncol});
    UL = c.ar.alloc(MPDO_C128, {Bn, (long long)k, nrow});
    ARENA_OK(c);
    EC(contract(c.st, su, {1, 1, 1}, Xs, {1, 1, 1}, XU, {1, 1, 1}));
    EC(contract(c.st, XU, {1, 1, 1}, transposed(Alast), {1, 1, 1}, UL, {1, 1, 1}, false, true));
    EC(rowscale(c, Wh_, s, (int)ncol, k, 0.5, 0.0, 0, zdt, &Zc));
  }
  *k_out = k;
  void* plo = alloc(0, (int64_t)Bn * l * 2 * a0 * k, user);
  void* phi = alloc(1, (int64_t)Bn * k * 2 * K * a1 * r, user);
  if (!plo || !phi) return fail(MPDO_EINVAL, "mpdo_split_2q: output allocation failed");
  Tn Tlo_n = Tn::contig(plo, dtype, {Bn, l, 2, a0, (long long)k});
  Tn Thi_n = Tn::contig(phi, dtype, {Bn, (long long)k, 2, (long long)K * a1, r});

  // Tlo'[b,l,p0,a0,j] = sum_x Q'[(l,a0),x] conj(UL[j,(x,p0)])
  Tn ULv = UL.view({Bn, (long long)k, x, 2});
  if (orthLo) {
    // W[b,c,(p0,j)] = sum_x conj(Xs_lo[x,c]) conj(UL[j,x,p0]);   Tlo' = Alo . W
    Tn W = c.ar.alloc(dtype, {Bn, 2LL * m, 2, (long long)k});
    ARENA_OK(c);
    EC(contract(c.st, transposed(Xs_lo), {1, 1, 1}, ULv.permute({0, 2, 3, 1}), {1, 1, 2}, W.view({Bn, 2LL * m, 2LL * k}),
                {1, 1, 1}, true, true));
    EC(contract(c.st, Alo, {1, 2, 2}, W.view({Bn, 2, (long long)m, 2, (long long)k}), {1, 2, 2},
                Tlo_n.permute({0, 1, 3, 2, 4}), {1, 2, 2}));
  } else {
    // Tlo'[b,l,p0,a0,j] = conj(UL[j,(l,a0),p0])
    Tn src = UL.view({Bn, (long long)k, (long long)l, (long long)a0, 2}).permute({0, 2, 4, 3, 1});
    EC(copy_view(c.st, src, Tlo_n, true));
  }
  // Thi'[b,j,p1,(g,a1),r] = sum_y Zc[j,p1,g,y] Qt'[y,(a1,r)]
  if (orthHi) {
    Tn ZF = c.ar.alloc(dtype, {Bn, (long long)k * 2 * K, 2LL * m});
    ARENA_OK(c);
    EC(contract(c.st, Zc.view({Bn, (long long)k * 2 * K, y}), {1, 1, 1}, F_hi, {1, 1, 1}, ZF, {1, 1, 1}));
    EC(contract(c.st, ZF.view({Bn, (long long)k * 2 * K, (long long)m, 2}), {1, 1, 2}, Mhi, {1, 2, 2},
                Thi_n.view({Bn, (long long)k * 2 * K, (long long)a1, (long long)r}), {1, 1, 2}));
  } else {
    EC(copy_view(c.st, Zc, Thi_n.view({Bn, (long long)k, ncol})));
  }
  return 0;
}
