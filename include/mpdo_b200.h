/*
 * mpdo_b200.h — C ABI of libmpdo_b200.so: the B200 (sm_100a) kernels behind the
 * MPDOSimulator noisy-gate update path.
 *
 * The reference (WeiguoMa/Tomography-assisted-MPDO-QCircuit) is pure Python; the only
 * seam it exposes for this path is the TensorNetwork-pytorch backend it asks users to
 * overwrite (README.md:17):
 *     decompositions.svd(torch, tensor, pivot_axis, max_singular_values,
 *                        max_truncation_error, relative)      decompositions.py:51-146
 *     decompositions.qr (torch, tensor, pivot_axis, ...)      decompositions.py:149-195
 * plus the tensornetwork 0.4.6 contraction calls reached from Circuit.py / TNNOptimizer.py
 * (tn.contract, tn.contract_between, tn.contractors.optimal/auto -> torch.tensordot).
 * Each entry point below names the reference call site(s) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and no entry
 *     point synchronises unless documented ("SYNC");
 *   - complex data are interleaved (re, im); dtype 0 = complex64, 1 = complex128;
 *   - tensors are row-major; every entry point carries a leading batch dimension (independent
 *     circuits / brick pairs);
 *   - return value: 0 on success, a negative MPDO_E* code on argument errors, or the positive
 *     cudaError_t of the failing runtime call. mpdo_last_error() gives a message.
 */
#ifndef MPDO_B200_H_
#define MPDO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPDO_C64 0
#define MPDO_C128 1

#define MPDO_EINVAL (-1)  /* bad argument */
#define MPDO_ENOSMEM (-2) /* problem does not fit the kernel's shared-memory tiling */

/* Composite index: a logical index i in [0, prod) addresses
 *   (i % d0) * s0 + ((i / d0) % d1) * s1 + (i / (d0*d1)) * s2     (elements, not bytes)
 * d0 <= 0 means single level (i * s0); d1 <= 0 means two levels. This is how one kernel
 * serves every axis grouping of T[l,s,a,r] without a transpose pass. */
typedef struct {
  int32_t d0, d1;
  int64_t s0, s1, s2;
} mpdo_idxmap;

/* Batched complex contraction  C[b,i,j] = alpha * sum_k opA(A[b,i,k]) * opB(B[b,k,j]) + beta * C[b,i,j]
 * with op = identity or complex conjugate, and every logical axis (b, i, j, k) a composite
 * index over up to three tensor axes of the operand.
 * Replaces: tn.contract / tn.contract_between / tn.contractors.optimal (torch.tensordot) at
 *   Circuit.py:104,171 (gate absorption), TNNOptimizer.py:105 (R absorbed into the right
 *   neighbour), :126 (two-site bond merge), :191 (U*S), and the `@` products inside
 *   decompositions.py:39,43,46. */
typedef struct {
  int32_t M, N, K, batch;
  int32_t dtypeA, dtypeB, dtypeC; /* MPDO_C64 / MPDO_C128 */
  int32_t conjA, conjB;
  int32_t acc64;                  /* 1: accumulate in fp64 (always when any operand is c128) */
  int32_t a_kfast, b_jfast;       /* loader hints: which axis of A / B is contiguous */
  int32_t ksplit;                 /* >1: split K over CTAs, atomically accumulated (C must be pre-zeroed or beta==1) */
  int32_t hermitian;              /* 1: the result is Hermitian (a Gram matrix; M == N, beta == 0): only the tiles on and
                                     below the diagonal are computed, the rest is written as the conjugate mirror */
  double alpha, beta;
  mpdo_idxmap Ab, Ai, Ak;
  mpdo_idxmap Bb, Bk, Bj;
  mpdo_idxmap Cb, Ci, Cj;
} mpdo_contract_desc;

int mpdo_contract(const mpdo_contract_desc* d, const void* A, const void* B, void* C, void* stream);

/* Single-qubit gate / Kraus absorption into a site tensor (inner index grows by K, new index major):
 *   Tout[b,l,p,(g,a),r] = sum_s G[b?,p,s,g] * T[b,l,s,a,r]
 * G is [2,2,K] row-major per batch entry (gBatchStride = 0 shares one gate across the batch).
 * Replaces: Circuit.py:138-178 (_apply_single_qubit_gate: tn.contract + tn.flatten_edges). */
int mpdo_absorb_1q(int dtype, int batch, int l, int a, int r, int K, const void* T, const void* G,
                   int64_t gBatchStride, void* Tout, void* stream);

/* One-sided (Hestenes) Jacobi on the rows of Y[b] (n rows, leading dimension ld, complex128):
 * rows are mixed by a unitary J until the first m entries of every pair of rows are orthogonal;
 * the mixing is applied to all mt >= m entries (entries m..mt-1 carry an accumulator, e.g. the
 * identity, which ends up holding J). In place. `work` must hold batch*48 int32, 8-byte aligned (scratch).
 * Converges when every |<y_p,y_q>| <= tol*|y_p||y_q|, at most maxSweeps (<= 32) sweeps.
 * Replaces: torch.linalg.svd (LAPACK gesdd) at decompositions.py:45,113 for the small cores. */
int mpdo_jacobi_rows(int batch, int n, int m, int mt, int ld, int64_t batchStride, void* Y, double tol,
                     int maxSweeps, int32_t* work, void* stream);

/* After mpdo_jacobi_rows: row norms (over the first m entries) sorted descending into s[b,n], and
 * optionally the sorted rows themselves:
 *   Yn[b,j,0..m)  = row / norm  if normalize (0 where norm <= zeroTol * maxnorm), else row      (may be NULL)
 *   Z [b,j,0..mz) = entries m..m+mz of the row (the accumulator)                                  (may be NULL)
 * Yn and Z are complex128, dense row-major. */
int mpdo_rows_finalize(int batch, int n, int m, int mz, int ld, int64_t batchStride, const void* Y, double* s,
                       void* Yn, void* Z, int normalize, double zeroTol, void* stream);

/* One call for the two small-core decompositions of the path: builds Y = [L | I] (L dense [b,n,m], Y scratch
 * [b,n,m+n], both complex128), runs mpdo_jacobi_rows on it and mpdo_rows_finalize into s / Yn / Z.
 *   SVD of L = Uh^h diag(s) Wh :  Yn = Wh (normalize = 1), Z = Uh
 *   eigen-decomposition of a Hermitian PSD G (m = n): rows of J.G are lam_j v_j^h, so s = lam, Z = Vh, Yn = NULL.
 * SYNC when the matrix needs more than one block pair (the sweep loop then polls convergence from the host).
 * Replaces: torch.linalg.svd at decompositions.py:45,113 (and, through the Gram matrices, torch.linalg.qr :187). */
int mpdo_decompose_rows(int batch, int n, int m, const void* L, void* Y, int32_t* work, double* s, void* Yn,
                        void* Z, int normalize, double zeroTol, double tol, int maxSweeps, void* stream);

/* Eigen-decomposition of Hermitian positive-semidefinite matrices G[b] (n x n, complex128, dense):
 *   G = Vh^h diag(lam) Vh,  lam[b,n] descending, Vh[b,n,n] rows = eigenvectors (conjugated).
 * precondition = 1: rank-revealing pivoted Cholesky G = L L^h, then one-sided Jacobi on the rows of L^h
 *   (Drmac-Veselic preconditioning: a handful of sweeps on rows of length n even for graded spectra). Directions
 *   whose pivot falls below rel * max diag(G) are treated as the null space: lam = 0 and zero rows of Vh.
 * precondition = 2: the same with a blocked Cholesky WITHOUT pivoting (rows ordered once by decreasing diagonal, two
 *   cluster barriers per 16-column panel instead of one per pivot: 0.64 -> ~0.2 ms at n = 192) where the kernel takes
 *   the shape (80 < n <= 256), the pivoted kernels otherwise. Meant for Gram matrices of fp32 data: a numerically
 *   null pivot is skipped, which without full pivoting can drop couplings of up to sqrt(rel) relative size - below
 *   fp32 resolution, not below fp64's. On the Gram matrices of the sweeps the Jacobi phase needs the same number of
 *   sweeps as after the pivoted factorisation.
 * precondition = 0: Jacobi on [G | I] (mpdo_decompose_rows); complete orthonormal basis.
 * `scratch`: mpdo_eigh_psd_scratch_bytes(batch, n) bytes of device memory, 256-byte aligned.
 * Replaces: the LAPACK eigen/SVD work behind torch.linalg.svd / torch.linalg.qr at decompositions.py:45,113,187
 * for every Gram matrix of the truncation path (TNNOptimizer.py:143-215 sweeps, Tools.py gate split). */
int64_t mpdo_eigh_psd_scratch_bytes(int batch, int n);
int mpdo_eigh_psd(int batch, int n, const void* G, void* scratch, double* lam, void* Vh, int precondition,
                  double rel, double tol, int maxSweeps, void* stream);

/* Rank-revealing pivoted Cholesky factorisation of Hermitian PSD matrices G[b] (n x n, complex128, dense):
 *   G = Lh^h Lh,  Lh[b,k,:] = conj(column k of L), columns in pivot order, rows >= rank[b] are zero;
 *   Linv (optional) is the matching left inverse: Linv . Lh^h = diag(1 (rank times), 0, ...), rows >= rank zero.
 * The factorisation stops when the largest remaining diagonal falls below rel * max diag(G).
 * This is all an orthogonalisation needs (Cholesky-QR): for a tall X with G = X^h X,  Q = X . Linv^h is an isometry
 * on the numerical range and X = Q . Lh; for a wide M with G = M M^h,  Qt = Linv . M and M = Lh^h . Qt.
 * `scratch`: mpdo_chol_psd_scratch_bytes(batch, n) bytes, 256-byte aligned. rank: [batch] int32 on the device or NULL.
 * Returns MPDO_ENOSMEM when batch * ceil(n / rows per CTA) CTAs cannot be co-resident (use mpdo_eigh_psd then).
 * Replaces: torch.linalg.qr (LAPACK geqrf/ungqr) at decompositions.py:187 for the left-to-right sweep
 * (TNNOptimizer.py:143-170) and for the two site factorisations of the gate split. */
int64_t mpdo_chol_psd_scratch_bytes(int batch, int n);
int mpdo_chol_psd(int batch, int n, const void* G, void* scratch, void* Lh, void* Linv, int32_t* rank, double rel,
                  void* stream);

/* X[b,j,c] = f(lam[b,j]) * V[b,j,c] for j < rows, c < cols, with f(x) = x^power and
 *   mode 0: f = 0 where lam[b,j] <= tol*lam[b,0]      (drop numerically null directions)
 *   mode 1: lam clamped from below at tol*lam[b,0]    (floor, for the first pass of a two-pass orthogonalisation)
 * V is complex128 [b, vRows, cols] (first `rows` rows used); X has dtype `dtypeX`. */
int mpdo_rowscale(int batch, int rows, int cols, int vRows, const void* V, const double* lam, int lamStride,
                  double power, double tol, int mode, int dtypeX, void* X, void* stream);

/* The reference's kept-rank rule (decompositions.py:117-134) evaluated on device for sorted
 * singular values s[b,0..n) (or their squares when squared = 1, as produced by a Gram eigen-solve):
 * keep[b] = min(cap, first idx+1 with ||s|| - ||s[:idx+1]|| <= eps), eps = maxTruncErr * s[0] if
 * relative. f32 = 1 reproduces the complex64 arithmetic (fp32 values, cumsum rounded to fp32 per
 * element). maxTruncErr < 0 means "None". zeroTail = 1 zeroes s[b, keep[b]..n) in place so that a
 * batch padded to a common rank carries exact zeros in the discarded directions. */
int mpdo_rank_rule(int batch, int n, double* s, int sStride, int squared, int cap, double maxTruncErr,
                   int relative, int f32, int32_t* keep, int zeroTail, void* stream);

/* ---- step-level entry points: one call per step of the update path (the kernel sequences of
 * MPDOSimulator/_engine/steps.py issued from C++). Site tensors are dense row-major T[B, l, 2, a, r]; npass = 1 for
 * complex64 states (one Gram-eig pass), 2 for complex128 (two passes + Jacobi SVD of the core). ------------------ */

/* One step of the left-to-right QR sweep: Ti [B,l,2,a,r] -> Q (isometry, same shape); Tn [B,r,2,a2,r2] <- R.Tn.
 * Replaces: tn.split_node_qr + tn.contract_between at TNNOptimizer.py:98-106 (decompositions.qr :149-195). */
int mpdo_qr_step(int dtype, int npass, int B, int l, int a, int r, const void* Ti, int a2, int r2, const void* Tn_in,
                 void* Q_out, void* Tn_out, void* stream);

/* One step of the right-to-left bond truncation to k = min(chi, l) singular values: Tl [B,lp,2,ap,l] (left
 * isometric), Tr [B,l,2,a,r] -> Tl.U sqrt(S) [B,lp,2,ap,k'], sqrt(S) Vh [B,k',2,a,r]; sv_out (optional, [B,l] doubles)
 * receives the singular values (their squares when npass = 1). max_err < 0: k' = k. max_err >= 0: the reference's
 * relative rule (decompositions.py:117-134 with relative=True: ||s|| - ||s[:k']|| <= max_err * s[0], batch maximum)
 * lowers k' further; SYNC (rank read-back); the outputs are written densely with k' into buffers sized for k, and
 * *k_out = k'.
 * Replaces: tn.contract_between + tn.split_node at TNNOptimizer.py:126-133 (decompositions.svd :51-146). */
int mpdo_bond_svd_step(int dtype, int npass, int B, int lp, int ap, int l, const void* Tl, int a, int r, const void* Tr,
                       int k, double max_err, int* k_out, void* Tl_out, void* Tr_out, double* sv_out, void* stream);

/* bondTruncate (QR sweep + chi sweep) in environment form, for complex64 states and a fixed chi: same singular values
 * and the same sqrt(S) | sqrt(S) split at every bond as mpdo_qr_step + mpdo_bond_svd_step, but only ONE sequential
 * chain of decompositions.
 * mpdo_env_sweep: sites T[i] [B, l[i], 2, a[i], r[i]] (l[0] = 1, l[i+1] = r[i]), i = 0 .. nsites-1. For every bond
 * j = 1 .. nsites-1 it forms the left environment E_j = A_j^h A_j (A_j = the block of sites 0 .. j-1; the chain
 * E_{i+1} = T_i^h E_i T_i is two contractions per site and is the Gram matrix the QR sweep would factor), factors it
 * rank-revealingly, E_j = C^h C, and writes the left inverse (Ci . C^h = 1 on the numerical range) to Ci_out[j]
 * ([B, l[j], l[j]] complex128) and the product C . T[j] to M_out[j] ([B, l[j], 2, a[j], r[j]], state dtype); entries 0
 * of the pointer arrays are unused. The factorisations and products of different bonds are independent of one another
 * and run on the library's side streams while the chain continues; the caller's stream waits for them before the
 * call's work counts as done.
 * mpdo_bond_env_step: one step of the right-to-left sweep on the un-canonicalised state. M0 = M_out[j] [B,l,2,a,r0],
 * W [B,r0,rw] (state dtype) what the previous step returned for the right index (NULL at the last site), Ci the
 * left inverse of bond j. With M = M0.W = U S V^h the reference's two-site matrix is Q.M for an isometry Q:
 * T_out = sqrt(S_k) V_k^h [B,k,2,a,rw], W_out = Ci^h U_k sqrt(S_k) [B,l,k] (to be applied to the right index of site
 * j-1: by the next step, or directly for site 0), sv_out (optional, [B,l]) the squared singular values.
 * Replaces: tn.split_node_qr / tn.contract_between / tn.split_node at TNNOptimizer.py:98-106 and :126-133
 * (bondTruncate, TNNOptimizer.py:72-84). */
int mpdo_env_sweep(int dtype, int B, int nsites, const int* l, const int* a, const int* r, const void* const* T,
                   void* const* Ci_out, void* const* M_out, void* stream);
int mpdo_bond_env_step(int dtype, int B, int l, int a, int r0, const void* M0, int rw, const void* W, const void* Ci,
                       int k, void* T_out, void* W_out, double* sv_out, void* stream);

/* Inner-index truncation T [B,l,2,a,r] -> U.S [B,l,2,k',r], k = min(kappa, a) (k = a for kappa = None); max_err and
 * k_out as above (TNNOptimizer.py:189 passes max_truncation_err with relative=True). disc_out (optional, [B] doubles)
 * = norm of the discarded part. SYNC when the top-k subspace iteration is used (max_err < 0, a >= 64 and a >= 8k) or
 * max_err >= 0.
 * Replaces: tn.split_node_full_svd + contract_between at TNNOptimizer.py:186-197. */
int mpdo_kappa_truncate(int dtype, int B, int l, int a, int r, const void* T, int k, double max_err, int* k_out,
                        void* T_out, double* disc_out, void* stream);

/* Two-qubit gate absorption and split: T_lo [B,l,2,a0,m], T_hi [B,m,2,a1,r], G [Bg,2,2,2,2,K] as
 * [p_lo,p_hi,s_lo,s_hi,g] (Bg = 1 or B, same dtype as the state) -> T_lo' [B,l,2,a0,k], T_hi' [B,k,2,K*a1,r] with the
 * kept rank k from the reference rule ||s|| - ||s[:k]|| <= max_err. The rank is data dependent: `alloc(which, count,
 * user)` is called once it is known and returns device memory for `count` complex elements (which = 0: T_lo',
 * 1: T_hi'). SYNC (rank read-back). ranks_out (optional, HOST memory, B ints) receives the kept rank of every batch
 * entry; *k_out is their maximum and entries with a smaller rank are padded with exact zeros.
 * Replaces: tn.contractors.optimal + tn.flatten_edges + tn.split_node at Circuit.py:104-124. */
typedef void* (*mpdo_alloc_fn)(int which, int64_t count, void* user);
int mpdo_split_2q(int dtype, int npass, int B, int l, int a0, int m, const void* Tlo, int a1, int r, const void* Thi,
                  int Bg, int K, const void* G, double max_err, mpdo_alloc_fn alloc, void* user, int* k_out,
                  int* ranks_out, void* stream);

/* Elementwise dtype conversion between complex64 and complex128 (count complex elements). */
int mpdo_cast(int dtypeIn, int dtypeOut, int64_t count, const void* in, void* out, void* stream);

/* Optional per-launch timing for the roofline leg of bench.py: while enabled, every contraction (class 0: fp32
 * accumulation, class 3: fp64 accumulation), Jacobi (class 1) and pivoted-Cholesky (class 2) launch is bracketed by CUDA
 * events on its own stream. mpdo_timing_enable(on) clears earlier records.
 * mpdo_timing_summary (SYNC) sums, over the recorded launches of one class with at least minFlops algorithmic flops:
 * device seconds, algorithmic flops (8*M*N*K per complex contraction, 4*(M+1)*N*K for a Hermitian result) and bytes,
 * the launch count, and the duration / flops of the largest launch. */
int mpdo_timing_enable(int on);
int mpdo_timing_summary(int cls, double minFlops, double* seconds, double* flops, double* bytes, int64_t* launches,
                        double* maxFlopsSeconds, double* maxFlops);

/* Tensor-core path of mpdo_contract for complex64 applies (tcgen05.mma kind::tf32 with a 3xTF32 split, TMA operand
 * tiles, accumulators in TMEM; csrc/tc_apply.cu). Eligible: all-complex64, fp32 accumulation, beta = 0, A and C plain
 * row-major matrices per batch entry (k / j contiguous), M >= 512, K >= 16 and even, N >= 8 and even; op(B) may be any
 * composite-index view. mode 0 routes those products to the FFMA tiles again (A/B measurements, parity tests);
 * mode 1 (default) picks the column tile so that the fp32 accumulator chains in TMEM stay short (tcgen05 accumulates
 * with truncation: relative error <= 7e-7 measured up to K = 512, like the FFMA tiles); mode 2 always uses the widest
 * tile (about 1.35x faster at K >= 256, relative error 1.3e-6 .. 1.8e-6). Returns the previous mode. */
int mpdo_tc_enable(int mode);

/* Hands the scratch memory cached by the step-level entry points (one stream-ordered pool per calling thread) back
 * to the driver. The step functions return cudaErrorMemoryAllocation (2) only after trying this themselves; a caller
 * that shares the device with another caching allocator (torch) releases that cache and retries. */
int mpdo_trim_pools(void);
/* Bytes the scratch pool of the current device holds from the driver right now / at its high-water mark (the pool is
 * pre-grown once at first use, MPDO_SCRATCH_PREWARM_MB, default 16 GiB; a high-water mark above that means the
 * workload outgrew the pre-grown block). Either pointer may be NULL. */
int mpdo_pool_stats(int64_t* reservedBytes, int64_t* reservedHighBytes);

/* Test hook for the bounded device-wide barrier of the persistent factorisation kernels: launches two CTAs of which
 * one never arrives. The waiting CTA gives up after a short poll limit and traps, so the call returns a CUDA error
 * (and the context is unusable afterwards: call it from a throw-away process). A barrier that times out in
 * production (limit 2^24 polls) fails the same way - loudly - rather than returning a half-updated factorisation. */
int mpdo_debug_barrier_timeout(void* stream);

/* Library / device information. */
int mpdo_version(void);
const char* mpdo_last_error(void);
int mpdo_device_info(int* smCount, int* smemPerBlockOptin, int* ccMajor, int* ccMinor);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t mpdo_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MPDO_B200_H_ */
