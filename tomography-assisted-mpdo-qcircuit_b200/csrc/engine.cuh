// Host-side building blocks of the native update-path engine (engine.cu): strided tensor views over device
// memory, a stream-ordered scratch arena, and typed wrappers over the kernels' C entry points.
#pragma once
#include <string.h>

#include <initializer_list>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpdo {
namespace eng {

constexpr int MAXD = 8;

struct Tn {  // strided view of interleaved-complex data on the device (sizes / strides in elements)
  char* p = nullptr;
  int dt = MPDO_C128;
  int nd = 0;
  long long sh[MAXD] = {0};
  long long st[MAXD] = {0};

  size_t esz() const { return dt == MPDO_C64 ? 8 : 16; }
  long long numel() const {
    long long n = 1;
    for (int i = 0; i < nd; ++i) n *= sh[i];
    return n;
  }
  bool contiguous() const {
    long long s = 1;
    for (int i = nd - 1; i >= 0; --i) {
      if (sh[i] != 1 && st[i] != s) return false;
      s *= sh[i];
    }
    return true;
  }
  static Tn contig(void* ptr, int dtype, std::initializer_list<long long> shape) {
    Tn t;
    t.p = (char*)ptr;
    t.dt = dtype;
    t.nd = (int)shape.size();
    int i = 0;
    for (long long s : shape) t.sh[i++] = s;
    long long s = 1;
    for (int d = t.nd - 1; d >= 0; --d) {
      t.st[d] = s;
      s *= t.sh[d];
    }
    return t;
  }
  Tn permute(std::initializer_list<int> perm) const {
    Tn t = *this;
    int i = 0;
    for (int d : perm) {
      t.sh[i] = sh[d];
      t.st[i] = st[d];
      ++i;
    }
    return t;
  }
  Tn view(std::initializer_list<long long> shape) const {  // only for contiguous data
    return contig(p, dt, shape);
  }
  Tn expand(int dim, long long n) const {  // dim must have extent 1
    Tn t = *this;
    t.sh[dim] = n;
    t.st[dim] = 0;
    return t;
  }
  Tn narrow(int dim, long long start, long long len) const {
    Tn t = *this;
    t.p = p + (size_t)(start * st[dim]) * esz();
    t.sh[dim] = len;
    return t;
  }
};

struct Roles {
  int nb, n1, n2;
};

// One stream-ordered scratch pool per device, shared by every calling thread. Strands run on different streams from
// different threads; with the driver's default pool a buffer freed on one stream and reused on another makes the
// second stream wait for the first (the allocator orders the reuse after the free), which serialises the strands.
// This pool forbids that kind of reuse (cudaMemPoolReuseAllowInternalDependencies = 0): a freed block goes to another
// stream only once the work before its free has completed (opportunistic reuse), otherwise the pool grows. Freed
// scratch stays cached (release threshold = max) until mpdo_trim_pools() or an allocation failure trims it.
cudaMemPool_t scratch_pool();
void trim_all_pools();

struct Arena {  // scratch buffers freed (stream-ordered) when the step returns
  cudaStream_t st;
  std::vector<void*> bufs;
  int err = 0;
  cudaMemPool_t pool = nullptr;
  explicit Arena(cudaStream_t s) : st(s), pool(scratch_pool()) {}
  ~Arena() {
    for (void* b : bufs) cudaFreeAsync(b, st);
  }
  void* raw(size_t bytes) {
    void* ptr = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = pool ? cudaMallocFromPoolAsync(&ptr, bytes, pool, st) : cudaMallocAsync(&ptr, bytes, st);
    if (e == cudaErrorMemoryAllocation) {   // take back what the other strands' pools cache, then retry once
      cudaGetLastError();
      cudaStreamSynchronize(st);
      trim_all_pools();
      e = pool ? cudaMallocFromPoolAsync(&ptr, bytes, pool, st) : cudaMallocAsync(&ptr, bytes, st);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      snprintf(g_err, sizeof(g_err), "cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
      err = (int)e;
      return nullptr;
    }
    bufs.push_back(ptr);
    return ptr;
  }
  Tn alloc(int dtype, std::initializer_list<long long> shape) {
    Tn t = Tn::contig(nullptr, dtype, shape);
    t.p = (char*)raw((size_t)t.numel() * t.esz());
    return t;
  }
  double* reals(long long n) { return (double*)raw(sizeof(double) * (size_t)n); }
};

int contract(cudaStream_t st, const Tn& A, Roles ra, const Tn& B, Roles rb, const Tn& C, Roles rc, bool conjA = false,
             bool conjB = false, int acc64 = -1, double alpha = 1.0, double beta = 0.0, bool hermitian = false);

}  // namespace eng
}  // namespace mpdo
