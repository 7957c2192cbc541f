// libhmx.so: host side of the C ABI declared in include/hmx.h.
//
// The micro cell kernel is compiled per coefficient program (hommx_b200/native.py) and
// arrives as a cubin image; it is loaded with the driver API, obtained through
// cudaGetDriverEntryPoint so that the library has no link-time dependency on libcuda.so
// (it must load -- and export its symbols -- on a machine without a GPU).
// The macro gather, halo pack/unpack and the peak microbenchmarks are compiled in.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hmx.h"
#include "hmx_cell_common.cuh"

namespace {

thread_local std::string g_create_error;

// ---- fixed kernels ------------------------------------------------------------------
// CSR value slot s = sum of its sources in fixed order (deterministic replacement of
// MatSetValues(ADD_VALUES), hmm.py:325-330).  HBM-bound streaming gather.
__global__ void __launch_bounds__(256) hmx_gather_csr(long long nnz, const long long* __restrict__ ptr,
                                                      const int* __restrict__ src, const double* __restrict__ S,
                                                      double* __restrict__ vals) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nnz; s += (long long)gridDim.x * blockDim.x) {
    const long long b = ptr[s], e = ptr[s + 1];
    double acc = 0.0;
    for (long long j = b; j < e; ++j) acc += S[src[j]];
    vals[s] = acc;
  }
}
__global__ void __launch_bounds__(256) hmx_halo_pack(long long n, const long long* __restrict__ slots,
                                                     const double* __restrict__ vals, double* __restrict__ buf) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
    buf[j] = vals[slots[j]];
}
__global__ void __launch_bounds__(256) hmx_halo_unpack(long long n, const long long* __restrict__ slots,
                                                       double* __restrict__ vals, const double* __restrict__ buf) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
    vals[slots[j]] = buf[j];
}
// ---- macro system on the device (Dirichlet lifting hmm.py:453-480, Krylov solve :482-483) ----
// b -= A u_bc on free rows; constrained rows/columns zeroed, unit diagonal, b = value (zeroRowsColumns).
__global__ void __launch_bounds__(256) hmx_lift(long long n, const long long* __restrict__ ptr, const int* __restrict__ idx,
                                                double* __restrict__ vals, const signed char* __restrict__ mask,
                                                const double* __restrict__ ubc, double* __restrict__ b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool fixed = mask[i] != 0;
    double acc = 0.0;
    for (long long k = ptr[i]; k < ptr[i + 1]; ++k) {
      const int j = idx[k];
      if (fixed) {
        vals[k] = (j == i) ? 1.0 : 0.0;
      } else if (mask[j]) {
        acc += vals[k] * ubc[j];
        vals[k] = 0.0;
      }
    }
    b[i] = fixed ? ubc[i] : b[i] - acc;
  }
}
// Macro element matrices from homogenised tensors at SEVERAL macro quadrature points per cell (a higher-order macro
// rule; the reference evaluates the barycentre only, hmm.py:349-352):  S_loc = |T| C^T (sum_q w_q A_hom(x_q)) C  --
// the P1 macro gradients / strains C are constant per cell, so the rule acts on the tensor alone.  One thread per entry.
template <int D, int KIND>
__global__ void __launch_bounds__(256) hmx_macro_elements(long long n_cells, const int* __restrict__ cell_nodes,
                                                          const double* __restrict__ node_xyz, int nq,
                                                          const double* __restrict__ wq, const double* __restrict__ A_pts,
                                                          double* __restrict__ S_loc) {
  constexpr int MV = KIND == 0 ? D : D * (D + 1) / 2, NB = KIND == 0 ? D + 1 : (D + 1) * D;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells; c += (long long)gridDim.x * blockDim.x) {
    double verts[(D + 1) * 3], Ah[MV * MV];
    for (int v = 0; v <= D; ++v)
      for (int k = 0; k < 3; ++k) verts[v * 3 + k] = node_xyz[(long long)cell_nodes[c * (D + 1) + v] * 3 + k];
    for (int e = 0; e < MV * MV; ++e) {
      double a = 0.0;
      for (int q = 0; q < nq; ++q) a += wq[q] * A_pts[(c * nq + q) * MV * MV + e];  // fixed order
      Ah[e] = a;
    }
    hmx::macro_element_matrix<D, KIND>(verts, Ah, S_loc + c * NB * NB);
  }
}
// Jacobi-PCG building blocks.  Scalars live in a small device array sc[]: 0 rz, 3 rz0; every dot product is reduced in
// two FIXED-ORDER stages (block partials bp[blockIdx.x], then the same sum of the partials in every consumer block), so
// the macro solve is bitwise reproducible like the assembly that feeds it (no floating-point atomics).
__device__ __forceinline__ double hmx_block_sum_fixed(double v) {  // every thread of the block returns the same total
  __shared__ double sh[33];
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  __syncthreads();  // (sh may still be read from the previous call)
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    sh[32] = s;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ double hmx_sum_partials(const double* __restrict__ bp, int g) {
  double v = 0.0;
  for (int i = threadIdx.x; i < g; i += blockDim.x) v += bp[i];
  return hmx_block_sum_fixed(v);
}
__global__ void __launch_bounds__(256) hmx_pcg_init(long long n, const long long* __restrict__ ptr, const int* __restrict__ idx,
                                                    const double* __restrict__ vals, const double* __restrict__ b,
                                                    double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                                                    double* __restrict__ dinv, double* __restrict__ bp) {
  double part = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double d = 0.0;
    for (long long k = ptr[i]; k < ptr[i + 1]; ++k)
      if (idx[k] == i) d = vals[k];
    const double di = d != 0.0 ? 1.0 / d : 0.0;
    dinv[i] = di;
    x[i] = 0.0;
    r[i] = b[i];
    p[i] = di * b[i];
    part += b[i] * di * b[i];
  }
  part = hmx_block_sum_fixed(part);
  if (threadIdx.x == 0) bp[blockIdx.x] = part;
}
// sc[0] = sc[3] = sum of the init partials (one block)
__global__ void __launch_bounds__(256) hmx_pcg_start(const double* __restrict__ bp, int g, double* __restrict__ sc) {
  const double s = hmx_sum_partials(bp, g);
  if (threadIdx.x == 0) {
    sc[0] = s;
    sc[3] = s;
  }
}
__global__ void __launch_bounds__(256) hmx_pcg_spmv(long long n, const long long* __restrict__ ptr, const int* __restrict__ idx,
                                                    const double* __restrict__ vals, const double* __restrict__ p,
                                                    double* __restrict__ y, double* __restrict__ bp_pap) {
  double part = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (long long k = ptr[i]; k < ptr[i + 1]; ++k) acc += vals[k] * p[idx[k]];
    y[i] = acc;
    part += p[i] * acc;
  }
  part = hmx_block_sum_fixed(part);
  if (threadIdx.x == 0) bp_pap[blockIdx.x] = part;
}
__global__ void __launch_bounds__(256) hmx_pcg_update(long long n, double* __restrict__ x, double* __restrict__ r,
                                                      const double* __restrict__ p, const double* __restrict__ y,
                                                      const double* __restrict__ dinv, const double* __restrict__ sc,
                                                      const double* __restrict__ bp_pap, double* __restrict__ bp_rz) {
  const double pAp = hmx_sum_partials(bp_pap, (int)gridDim.x);
  const double alpha = pAp > 0.0 ? sc[0] / pAp : 0.0;
  double part = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * y[i];
    r[i] = ri;
    part += ri * dinv[i] * ri;
  }
  part = hmx_block_sum_fixed(part);
  if (threadIdx.x == 0) bp_rz[blockIdx.x] = part;
}
__global__ void __launch_bounds__(256) hmx_pcg_direction(long long n, const double* __restrict__ r, double* __restrict__ p,
                                                         const double* __restrict__ dinv, const double* __restrict__ sc,
                                                         const double* __restrict__ bp_rz) {
  const double rz_new = hmx_sum_partials(bp_rz, (int)gridDim.x);
  const double beta = sc[0] > 0.0 ? rz_new / sc[0] : 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = dinv[i] * r[i] + beta * p[i];
}
__global__ void __launch_bounds__(256) hmx_pcg_rotate(double* sc, const double* __restrict__ bp_rz, int g) {  // rz <- rz_new
  const double rz_new = hmx_sum_partials(bp_rz, g);
  if (threadIdx.x == 0) sc[0] = rz_new;
}
// FP64 peak: 8 independent DFMA chains per thread, 1024 threads/SM
__global__ void __launch_bounds__(256) hmx_dfma_peak(double* out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v0 = fma(v0, a, b);
      v1 = fma(v1, a, b);
      v2 = fma(v2, a, b);
      v3 = fma(v3, a, b);
      v4 = fma(v4, a, b);
      v5 = fma(v5, a, b);
      v6 = fma(v6, a, b);
      v7 = fma(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}
__global__ void __launch_bounds__(256) hmx_copy(const double2* __restrict__ a, double2* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    b[i] = a[i];
}

// ---- driver API entry points ----------------------------------------------------------
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*OccupancyMaxActiveClusters)(int*, CUfunction, const CUlaunchConfig*) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
  std::string why;
};

template <class F>
bool load_entry(const char* name, F& fn, std::string& why) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || p == nullptr) {
    why = std::string("cudaGetDriverEntryPoint(") + name + "): " + cudaGetErrorString(e);
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

Driver& driver() {
  static Driver d;
  if (!d.ok && d.why.empty()) {
    d.ok = load_entry("cuModuleLoadData", d.ModuleLoadData, d.why) && load_entry("cuModuleUnload", d.ModuleUnload, d.why) &&
           load_entry("cuModuleGetFunction", d.ModuleGetFunction, d.why) &&
           load_entry("cuModuleGetGlobal", d.ModuleGetGlobal, d.why) &&
           load_entry("cuFuncSetAttribute", d.FuncSetAttribute, d.why) && load_entry("cuLaunchKernel", d.LaunchKernel, d.why) &&
           load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", d.OccupancyMaxActiveBlocksPerMultiprocessor, d.why) &&
           load_entry("cuOccupancyMaxActiveClusters", d.OccupancyMaxActiveClusters, d.why) &&
           load_entry("cuGetErrorString", d.GetErrorString, d.why);
  }
  return d;
}

// Makes `device` current for the scope of one entry point and restores the caller's device afterwards (the library
// must not change the calling thread's current device behind its back).
struct DeviceGuard {
  int prev = -1;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) status = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct hmx_handle {
  int dim = 0, kind = 0, n_micro = 0, nq = 0, device = 0;
  double rtol = 1e-8, atol = 1e-10;
  int max_it = 10000;
  int info[8] = {0};
  // element-list kernel: the micro mesh on the device
  DevBuf mm_buf[11], mm_struct;
  int natoms = 1;  // hmx_info[9] of the kernel image: atoms of the coefficient program (element-list scratch)
  long long mm_scratch = 0;  // per-CTA scratch doubles for this mesh
  int cluster = 1;       // CTAs per thread-block cluster of the cell kernel (hmx_info[8]; 1: ordinary launch)
  int max_clusters = 0;  // clusters of the cell kernel that are resident at once
  int grid_override = 0;
  CUmodule module = nullptr;
  CUfunction fn = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  DevBuf qp, qw, scratch, work;
  // staging for the host-pointer entry points
  DevBuf d_x, d_A, d_it, d_res, d_cells, d_xyz, d_ptr, d_src, d_vals, d_S;
  DevBuf m_work;  // macro PCG: r, p, y, dinv, scalars
  // macro load vector (hmx_macro_load_dev): the module of the last f, its quadrature table
  CUmodule load_module = nullptr;
  CUfunction load_fn = nullptr;
  unsigned long long load_hash = 0;
  int load_info[4] = {0, 0, 0, 0};
  DevBuf l_qp, l_qw;
  mutable std::string err;
  int ntypes() const { return dim == 2 ? 2 : 6; }
  int m() const { return kind == HMX_POISSON ? dim : dim * (dim + 1) / 2; }
  int nb() const { return kind == HMX_POISSON ? dim + 1 : (dim + 1) * dim; }
};

namespace {

int fail(const hmx_t* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define HMX_CUDA(h, call)                                                                        \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(h, HMX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#define HMX_CU(h, call)                                                              \
  do {                                                                               \
    CUresult r_ = (call);                                                            \
    if (r_ != CUDA_SUCCESS) {                                                        \
      const char* s_ = nullptr;                                                      \
      driver().GetErrorString(r_, &s_);                                              \
      return fail(h, HMX_ERR_CUDA, "%s: %s", #call, s_ ? s_ : "unknown driver error"); \
    }                                                                                \
  } while (0)

int grid_1d(long long n, int threads, int sms) {
  long long g = (n + threads - 1) / threads;
  long long cap = (long long)sms * 8;
  return (int)std::max<long long>(1, std::min(g, cap));
}

int launch_cell(hmx_t* h, long long n_pts, const double* x_pts, const int* cell_nodes, const double* node_xyz,
                double* A_hom, double* S_loc, int* iters, double* resid, double* chi = nullptr) {
  if (n_pts == 0) return HMX_OK;
  const int per_sm = std::max(1, h->info[5]);
  const int sms = h->info[6];
  long long grid = (long long)per_sm * sms;
  if (h->grid_override > 0) grid = h->grid_override;
  grid = std::max<long long>(1, std::min<long long>(grid, n_pts));
  if (h->cluster > 1) {  // one macro point per cluster: a whole number of clusters, at most the resident ones
    long long ncl = h->grid_override > 0 ? std::max<long long>(1, h->grid_override / h->cluster) : h->max_clusters;
    ncl = std::max<long long>(1, std::min<long long>(ncl, n_pts));
    grid = ncl * h->cluster;
  }
  const size_t scratch_doubles = (size_t)(h->mm_scratch > 0 ? h->mm_scratch : h->info[7]) * (size_t)grid;
  if (scratch_doubles) HMX_CUDA(h, h->scratch.reserve(scratch_doubles * sizeof(double)));
  hmx::CellParams P;
  P.n_pts = n_pts;
  P.x_pts = x_pts;
  P.cell_nodes = cell_nodes;
  P.node_xyz = node_xyz;
  P.A_hom = A_hom;
  P.S_loc = S_loc;
  P.iters = iters;
  P.resid = resid;
  P.qp = h->qp.as<double>();
  P.qw = h->qw.as<double>();
  P.scratch = h->scratch.as<double>();
  P.chi = chi;
  P.work = h->work.as<unsigned long long>();
  P.nq = h->nq;
  P.max_it = h->max_it;
  P.rtol = h->rtol;
  P.atol = h->atol;
  P.mesh = h->mm_scratch > 0 ? h->mm_struct.as<hmx::MicroMesh>() : nullptr;
  void* args[] = {&P};
  HMX_CU(h, driver().LaunchKernel(h->fn, (unsigned)grid, 1, 1, (unsigned)h->info[1], 1, 1, (unsigned)h->info[0],
                                  (CUstream)h->stream, args, nullptr));
  return HMX_OK;
}

}  // namespace

extern "C" {

int32_t hmx_abi_version(void) { return HMX_ABI_VERSION; }

const char* hmx_last_error(const hmx_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int hmx_create(hmx_t** out, const hmx_desc* d) {
  if (!out || !d) return fail(nullptr, HMX_ERR_ARG, "hmx_create: null argument");
  *out = nullptr;
  if (d->dim != 2 && d->dim != 3) return fail(nullptr, HMX_ERR_ARG, "Topology should be 3D or 2D");  // hmm.py:104-105
  if (d->kind != HMX_POISSON && d->kind != HMX_ELASTICITY) return fail(nullptr, HMX_ERR_ARG, "unknown problem kind %d", d->kind);
  const hmx_micro_mesh* um = d->micro_mesh;
  if (!um && d->n_micro < 2) return fail(nullptr, HMX_ERR_ARG, "the periodic micro mesh needs at least 2 cells per axis");
  if (um && (um->n_elem < 1 || um->n_nodes < 1 || um->nnzb < um->n_nodes || !um->elem_nodes || !um->elem_grad || !um->elem_vol ||
             !um->elem_yq || !um->row_ptr || !um->col || !um->blk_ptr || !um->blk_src || !um->node_ptr || !um->node_src || !um->diag))
    return fail(nullptr, HMX_ERR_ARG, "incomplete micro mesh description");
  if (d->nq < 1 || (!um && !d->qp) || !d->qw) return fail(nullptr, HMX_ERR_ARG, "quadrature table missing");
  if (!d->kernel_image || d->kernel_image_size == 0)
    return fail(nullptr, HMX_ERR_KERNEL, "no cell kernel image: the CUDA path has no fallback");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, HMX_ERR_CUDA, "no CUDA device available (the hot path has no CPU fallback)");
  }
  if (d->device < 0 || d->device >= ndev) return fail(nullptr, HMX_ERR_ARG, "device %d out of range (%d devices)", d->device, ndev);
  DeviceGuard guard(d->device);
  HMX_CUDA(nullptr, guard.status);
  HMX_CUDA(nullptr, cudaFree(nullptr));  // make the primary context current for the driver API
  Driver& drv = driver();
  if (!drv.ok) return fail(nullptr, HMX_ERR_CUDA, "%s", drv.why.c_str());

  hmx_t* h = new hmx_t;
  h->dim = d->dim;
  h->kind = d->kind;
  h->n_micro = um ? 0 : d->n_micro;
  h->nq = d->nq;
  h->device = d->device;
  h->rtol = d->rtol;
  h->atol = d->atol;
  h->max_it = d->max_it > 0 ? d->max_it : 10000;
  auto bail = [&](int code) {
    g_create_error = h->err;
    hmx_destroy(h);
    return code;
  };
  {
    CUresult r = drv.ModuleLoadData(&h->module, d->kernel_image);
    if (r != CUDA_SUCCESS) {
      const char* s = nullptr;
      drv.GetErrorString(r, &s);
      fail(h, HMX_ERR_KERNEL, "cuModuleLoadData: %s (image not built for this GPU?)", s ? s : "?");
      return bail(HMX_ERR_KERNEL);
    }
    r = drv.ModuleGetFunction(&h->fn, h->module, "hmx_cell");
    if (r != CUDA_SUCCESS) {
      fail(h, HMX_ERR_KERNEL, "kernel image has no entry 'hmx_cell'");
      return bail(HMX_ERR_KERNEL);
    }
    CUdeviceptr gp = 0;
    size_t gs = 0;
    r = drv.ModuleGetGlobal(&gp, &gs, h->module, "hmx_info");
    if (r != CUDA_SUCCESS || gs < 8 * sizeof(int)) {
      fail(h, HMX_ERR_KERNEL, "kernel image has no 'hmx_info' table");
      return bail(HMX_ERR_KERNEL);
    }
    int ki[12] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
    if (cudaMemcpy(ki, (const void*)gp, std::min(gs, sizeof ki), cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail(h, HMX_ERR_CUDA, "reading hmx_info failed: %s", cudaGetErrorString(cudaGetLastError()));
      return bail(HMX_ERR_CUDA);
    }
    // ki: 0 smem bytes, 1 threads, 2 nrhs, 3 dim, 4 kind, 5 n_micro, 6 scratch doubles per CTA, 7 reserved
    if (ki[3] != d->dim || ki[4] != d->kind || ki[5] != h->n_micro) {
      fail(h, HMX_ERR_KERNEL, "kernel image was built for dim=%d kind=%d n=%d, descriptor says dim=%d kind=%d n=%d", ki[3], ki[4],
           ki[5], d->dim, d->kind, d->n_micro);
      return bail(HMX_ERR_KERNEL);
    }
    h->info[0] = ki[0];
    h->info[1] = ki[1];
    h->info[2] = ki[2];
    h->info[3] = h->m();
    h->info[4] = h->nb();
    h->info[7] = ki[6];
    r = drv.FuncSetAttribute(h->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, ki[0]);
    if (r != CUDA_SUCCESS) {
      fail(h, HMX_ERR_KERNEL, "cell kernel needs %d bytes of shared memory per CTA: not available on this device", ki[0]);
      return bail(HMX_ERR_KERNEL);
    }
    int per_sm = 0;
    r = drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, h->fn, ki[1], (size_t)ki[0]);
    if (r != CUDA_SUCCESS || per_sm < 1) {
      fail(h, HMX_ERR_KERNEL, "cell kernel cannot be resident on an SM (threads=%d smem=%d)", ki[1], ki[0]);
      return bail(HMX_ERR_KERNEL);
    }
    h->info[5] = per_sm;
    h->cluster = ki[8] > 1 ? ki[8] : 1;
    h->natoms = ki[9] > 0 ? ki[9] : 1;
    if (h->cluster > 1) {
      // the kernel carries its cluster size (__cluster_dims__): a plain launch of a multiple of it forms the clusters
      if (h->cluster > 8) {
        r = drv.FuncSetAttribute(h->fn, CU_FUNC_ATTRIBUTE_NON_PORTABLE_CLUSTER_SIZE_ALLOWED, 1);
        if (r != CUDA_SUCCESS) {
          fail(h, HMX_ERR_KERNEL, "cell kernel needs clusters of %d CTAs: not available on this device", h->cluster);
          return bail(HMX_ERR_KERNEL);
        }
      }
      CUlaunchAttribute attr;
      attr.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
      attr.value.clusterDim.x = (unsigned)h->cluster;
      attr.value.clusterDim.y = 1;
      attr.value.clusterDim.z = 1;
      CUlaunchConfig cfg;
      std::memset(&cfg, 0, sizeof cfg);
      cfg.gridDimX = (unsigned)h->cluster;
      cfg.gridDimY = cfg.gridDimZ = 1;
      cfg.blockDimX = (unsigned)ki[1];
      cfg.blockDimY = cfg.blockDimZ = 1;
      cfg.sharedMemBytes = (unsigned)ki[0];
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int ncl = 0;
      r = drv.OccupancyMaxActiveClusters(&ncl, h->fn, &cfg);
      if (r != CUDA_SUCCESS || ncl < 1) {
        fail(h, HMX_ERR_KERNEL, "no cluster of %d CTAs (threads=%d smem=%d) can be resident on this device", h->cluster, ki[1], ki[0]);
        return bail(HMX_ERR_KERNEL);
      }
      h->max_clusters = ncl;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d->device) != cudaSuccess) {
      fail(h, HMX_ERR_CUDA, "cudaGetDeviceProperties failed");
      return bail(HMX_ERR_CUDA);
    }
    h->info[6] = prop.multiProcessorCount;
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    fail(h, HMX_ERR_CUDA, "cudaStreamCreate failed");
    return bail(HMX_ERR_CUDA);
  }
  h->own_stream = true;
  const size_t nqp = um ? 1 : (size_t)h->ntypes() * d->nq * d->dim;
  if (h->qp.reserve(nqp * sizeof(double)) != cudaSuccess || h->qw.reserve(d->nq * sizeof(double)) != cudaSuccess ||
      (!um && cudaMemcpy(h->qp.p, d->qp, nqp * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) ||
      cudaMemcpy(h->qw.p, d->qw, d->nq * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    fail(h, HMX_ERR_CUDA, "uploading the quadrature table failed: %s", cudaGetErrorString(cudaGetLastError()));
    return bail(HMX_ERR_CUDA);
  }
  if (um) {  // the general micro mesh: copy every array, then the struct of device pointers
    const int nv = d->dim + 1;
    const size_t E = (size_t)um->n_elem, N = (size_t)um->n_nodes, Z = (size_t)um->nnzb;
    const size_t nblk = (size_t)um->blk_ptr[Z], nnode = (size_t)um->node_ptr[N];
    const void* src[11] = {um->elem_nodes, um->elem_grad, um->elem_vol, um->elem_yq, um->row_ptr, um->col,
                           um->blk_ptr,    um->blk_src,   um->node_ptr, um->node_src, um->diag};
    const size_t bytes[11] = {E * nv * sizeof(int), E * nv * d->dim * sizeof(double), E * sizeof(double),
                              E * d->nq * d->dim * sizeof(double), (N + 1) * sizeof(int), Z * sizeof(int),
                              (Z + 1) * sizeof(int), nblk * sizeof(int), (N + 1) * sizeof(int), nnode * sizeof(int), N * sizeof(int)};
    if (nblk != E * nv * nv || nnode != E * nv) {
      fail(h, HMX_ERR_ARG, "micro mesh: the contribution lists must hold every (element, vertex[, vertex]) once");
      return bail(HMX_ERR_ARG);
    }
    for (int k = 0; k < 11; ++k)
      if (h->mm_buf[k].reserve(std::max<size_t>(bytes[k], 8)) != cudaSuccess ||
          cudaMemcpy(h->mm_buf[k].p, src[k], bytes[k], cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(h, HMX_ERR_CUDA, "uploading the micro mesh failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(HMX_ERR_CUDA);
      }
    hmx::MicroMesh mm;
    mm.n_elem = um->n_elem;
    mm.n_nodes = um->n_nodes;
    mm.nnzb = um->nnzb;
    mm.nq = d->nq;
    mm.elem_nodes = h->mm_buf[0].as<int>();
    mm.elem_grad = h->mm_buf[1].as<double>();
    mm.elem_vol = h->mm_buf[2].as<double>();
    mm.elem_yq = h->mm_buf[3].as<double>();
    mm.row_ptr = h->mm_buf[4].as<int>();
    mm.col = h->mm_buf[5].as<int>();
    mm.blk_ptr = h->mm_buf[6].as<int>();
    mm.blk_src = h->mm_buf[7].as<int>();
    mm.node_ptr = h->mm_buf[8].as<int>();
    mm.node_src = h->mm_buf[9].as<int>();
    mm.diag = h->mm_buf[10].as<int>();
    if (h->mm_struct.reserve(sizeof mm) != cudaSuccess || cudaMemcpy(h->mm_struct.p, &mm, sizeof mm, cudaMemcpyHostToDevice) != cudaSuccess) {
      fail(h, HMX_ERR_CUDA, "uploading the micro mesh failed: %s", cudaGetErrorString(cudaGetLastError()));
      return bail(HMX_ERR_CUDA);
    }
    const int bs = d->kind == HMX_POISSON ? 1 : d->dim;
    h->mm_scratch = hmx::element_list_scratch(um->n_elem, um->n_nodes, um->nnzb, std::max(1, h->natoms), bs, h->info[2]);
    h->info[7] = (int)std::min<long long>(h->mm_scratch, 2147483647LL);
  }
  if (h->work.reserve(sizeof(unsigned long long)) != cudaSuccess || cudaMemset(h->work.p, 0, sizeof(unsigned long long)) != cudaSuccess) {
    fail(h, HMX_ERR_CUDA, "allocating the work counter failed");
    return bail(HMX_ERR_CUDA);
  }
  *out = h;
  return HMX_OK;
}

void hmx_destroy(hmx_t* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (DevBuf& b : h->mm_buf) b.release();
  h->mm_struct.release();
  for (DevBuf* b : {&h->qp, &h->qw, &h->scratch, &h->work, &h->d_x, &h->d_A, &h->d_it, &h->d_res, &h->d_cells, &h->d_xyz, &h->d_ptr,
                    &h->d_src, &h->d_vals, &h->d_S, &h->m_work, &h->l_qp, &h->l_qw})
    b->release();
  if (h->load_module && driver().ok) driver().ModuleUnload(h->load_module);
  if (h->module && driver().ok) driver().ModuleUnload(h->module);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int hmx_set_stream(hmx_t* h, void* s) {
  if (!h) return HMX_ERR_ARG;
  if (h->own_stream && h->stream) {
    cudaStreamSynchronize(h->stream);
    cudaStreamDestroy(h->stream);
  }
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  return HMX_OK;
}

int hmx_set_tolerances(hmx_t* h, double rtol, double atol, int32_t max_it) {
  if (!h) return HMX_ERR_ARG;
  if (rtol < 0 || atol < 0 || max_it < 0) return fail(h, HMX_ERR_ARG, "tolerances must be non-negative");
  h->rtol = rtol;
  h->atol = atol;
  if (max_it > 0) h->max_it = max_it;
  return HMX_OK;
}

int hmx_set_grid(hmx_t* h, int32_t n) {
  if (!h || n < 0) return HMX_ERR_ARG;
  h->grid_override = n;
  return HMX_OK;
}

int hmx_cluster_info(const hmx_t* h, int32_t info[2]) {
  if (!h || !info) return HMX_ERR_ARG;
  info[0] = h->cluster;
  info[1] = h->cluster > 1 ? h->max_clusters : 0;
  return HMX_OK;
}

int hmx_kernel_info(const hmx_t* h, int32_t info[8]) {
  if (!h || !info) return HMX_ERR_ARG;
  for (int i = 0; i < 8; ++i) info[i] = h->info[i];
  return HMX_OK;
}

int hmx_sync(hmx_t* h) {
  if (!h) return HMX_ERR_ARG;
  HMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return HMX_OK;
}

int hmx_rhs_iterations(hmx_t* h, int64_t* total, int32_t reset) {
  if (!h || !total) return HMX_ERR_ARG;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  unsigned long long v = 0;
  HMX_CUDA(h, cudaMemcpyAsync(&v, h->work.p, sizeof v, cudaMemcpyDeviceToHost, h->stream));
  if (reset) HMX_CUDA(h, cudaMemsetAsync(h->work.p, 0, sizeof v, h->stream));
  HMX_CUDA(h, cudaStreamSynchronize(h->stream));
  *total = (int64_t)v;
  return HMX_OK;
}

int hmx_gather_csr_dev(hmx_t* h, int64_t nnz, const int64_t* gather_ptr, const int32_t* gather_src, const double* S_loc,
                       double* csr_vals) {
  if (!h) return HMX_ERR_ARG;
  if (nnz < 0 || (nnz > 0 && (!gather_ptr || !csr_vals))) return fail(h, HMX_ERR_ARG, "hmx_gather_csr: null CSR buffer");
  if (nnz == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  hmx_gather_csr<<<grid_1d(nnz, 256, h->info[6]), 256, 0, h->stream>>>(nnz, (const long long*)gather_ptr, gather_src, S_loc, csr_vals);
  HMX_CUDA(h, cudaGetLastError());
  return HMX_OK;
}

int hmx_cell_tensors_dev(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, int32_t* iters, double* resid) {
  if (!h) return HMX_ERR_ARG;
  if (n_pts < 0 || (n_pts > 0 && (!x_pts || !A_hom))) return fail(h, HMX_ERR_ARG, "hmx_cell_tensors: null buffer");
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  return launch_cell(h, n_pts, x_pts, nullptr, nullptr, A_hom, nullptr, iters, resid);
}

int hmx_cell_correctors_dev(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, double* chi) {
  if (!h) return HMX_ERR_ARG;
  if (n_pts < 0 || (n_pts > 0 && (!x_pts || !chi))) return fail(h, HMX_ERR_ARG, "hmx_cell_correctors: null buffer");
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  return launch_cell(h, n_pts, x_pts, nullptr, nullptr, A_hom, nullptr, nullptr, nullptr, chi);
}

int hmx_cell_tensors(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, int32_t* iters, double* resid) {
  if (!h) return HMX_ERR_ARG;
  if (n_pts < 0 || (n_pts > 0 && (!x_pts || !A_hom))) return fail(h, HMX_ERR_ARG, "hmx_cell_tensors: null buffer");
  if (n_pts == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  const size_t mm = (size_t)h->m() * h->m();
  HMX_CUDA(h, h->d_x.reserve(n_pts * 3 * sizeof(double)));
  HMX_CUDA(h, h->d_A.reserve(n_pts * mm * sizeof(double)));
  HMX_CUDA(h, h->d_it.reserve(n_pts * sizeof(int)));
  HMX_CUDA(h, h->d_res.reserve(n_pts * sizeof(double)));
  HMX_CUDA(h, cudaMemcpyAsync(h->d_x.p, x_pts, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = launch_cell(h, n_pts, h->d_x.as<double>(), nullptr, nullptr, h->d_A.as<double>(), nullptr, h->d_it.as<int>(),
                       h->d_res.as<double>());
  if (rc != HMX_OK) return rc;
  HMX_CUDA(h, cudaMemcpyAsync(A_hom, h->d_A.p, n_pts * mm * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (iters) HMX_CUDA(h, cudaMemcpyAsync(iters, h->d_it.p, n_pts * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (resid) HMX_CUDA(h, cudaMemcpyAsync(resid, h->d_res.p, n_pts * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return HMX_OK;
}

int hmx_assemble_macro_dev(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, int64_t n_nodes, const double* node_xyz,
                           int64_t nnz, const int64_t* gather_ptr, const int32_t* gather_src, double* csr_vals, double* S_loc,
                           int32_t* iters, double* resid) {
  if (!h) return HMX_ERR_ARG;
  if (n_cells < 0 || nnz < 0 || n_nodes < 0) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: negative size");
  if (n_cells > 0 && (!cell_nodes || !node_xyz)) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: null mesh buffer");
  if (nnz > 0 && (!gather_ptr || !csr_vals || (n_cells > 0 && !gather_src)))
    return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: null CSR buffer");
  const size_t nb2 = (size_t)h->nb() * h->nb();
  if ((double)n_cells * (double)nb2 > 2147483647.0)
    return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: n_cells*n_b^2 exceeds the int32 range of gather_src; shard the cells");
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  double* S = S_loc;
  if (!S && n_cells > 0) {
    HMX_CUDA(h, h->d_S.reserve(n_cells * nb2 * sizeof(double)));
    S = h->d_S.as<double>();
  }
  int rc = launch_cell(h, n_cells, nullptr, cell_nodes, node_xyz, nullptr, S, iters, resid);
  if (rc != HMX_OK) return rc;
  return hmx_gather_csr_dev(h, nnz, gather_ptr, gather_src, S, csr_vals);
}

int hmx_assemble_macro(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, int64_t n_nodes, const double* node_xyz, int64_t nnz,
                       const int64_t* gather_ptr, const int32_t* gather_src, double* csr_vals, double* S_loc, int32_t* iters,
                       double* resid) {
  if (!h) return HMX_ERR_ARG;
  if (n_cells < 0 || nnz < 0 || n_nodes < 0) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: negative size");
  if (n_cells > 0 && (!cell_nodes || !node_xyz)) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: null mesh buffer");
  if (nnz > 0 && (!gather_ptr || !csr_vals)) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: null CSR buffer");
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  const int nv = h->dim + 1;
  const size_t nb2 = (size_t)h->nb() * h->nb();
  std::vector<int64_t> last(1, 0);
  int64_t nsrc = 0;
  if (nnz > 0) nsrc = gather_ptr[nnz];
  if (nsrc < 0 || (double)nsrc > (double)n_cells * (double)nb2)
    return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: gather_ptr[nnz]=%lld exceeds n_cells*n_b^2", (long long)nsrc);
  if (nsrc > 0 && !gather_src) return fail(h, HMX_ERR_ARG, "hmx_assemble_macro: null gather_src");
  HMX_CUDA(h, h->d_cells.reserve(std::max<size_t>(1, n_cells * nv * sizeof(int))));
  HMX_CUDA(h, h->d_xyz.reserve(std::max<size_t>(1, n_nodes * 3 * sizeof(double))));
  HMX_CUDA(h, h->d_ptr.reserve((nnz + 1) * sizeof(int64_t)));
  HMX_CUDA(h, h->d_src.reserve(std::max<size_t>(1, nsrc * sizeof(int))));
  HMX_CUDA(h, h->d_vals.reserve(std::max<size_t>(1, nnz * sizeof(double))));
  HMX_CUDA(h, h->d_S.reserve(std::max<size_t>(1, n_cells * nb2 * sizeof(double))));
  HMX_CUDA(h, h->d_it.reserve(std::max<size_t>(1, n_cells * sizeof(int))));
  HMX_CUDA(h, h->d_res.reserve(std::max<size_t>(1, n_cells * sizeof(double))));
  if (n_cells > 0) {
    HMX_CUDA(h, cudaMemcpyAsync(h->d_cells.p, cell_nodes, n_cells * nv * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    HMX_CUDA(h, cudaMemcpyAsync(h->d_xyz.p, node_xyz, n_nodes * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (nnz > 0) {
    HMX_CUDA(h, cudaMemcpyAsync(h->d_ptr.p, gather_ptr, (nnz + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    if (nsrc > 0) HMX_CUDA(h, cudaMemcpyAsync(h->d_src.p, gather_src, nsrc * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  int rc = hmx_assemble_macro_dev(h, n_cells, h->d_cells.as<int>(), n_nodes, h->d_xyz.as<double>(), nnz, h->d_ptr.as<int64_t>(),
                                  h->d_src.as<int>(), h->d_vals.as<double>(), h->d_S.as<double>(), h->d_it.as<int>(),
                                  h->d_res.as<double>());
  if (rc != HMX_OK) return rc;
  if (nnz > 0) HMX_CUDA(h, cudaMemcpyAsync(csr_vals, h->d_vals.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (n_cells > 0) {
    if (S_loc) HMX_CUDA(h, cudaMemcpyAsync(S_loc, h->d_S.p, n_cells * nb2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (iters) HMX_CUDA(h, cudaMemcpyAsync(iters, h->d_it.p, n_cells * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (resid) HMX_CUDA(h, cudaMemcpyAsync(resid, h->d_res.p, n_cells * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  HMX_CUDA(h, cudaStreamSynchronize(h->stream));
  return HMX_OK;
}

int hmx_halo_pack_dev(hmx_t* h, const double* csr_vals, const int64_t* slots, int64_t n, double* buf) {
  if (!h) return HMX_ERR_ARG;
  if (n < 0 || (n > 0 && (!csr_vals || !slots || !buf))) return fail(h, HMX_ERR_ARG, "hmx_halo_pack: null buffer");
  if (n == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  hmx_halo_pack<<<grid_1d(n, 256, h->info[6]), 256, 0, h->stream>>>(n, (const long long*)slots, csr_vals, buf);
  HMX_CUDA(h, cudaGetLastError());
  return HMX_OK;
}

int hmx_halo_unpack_dev(hmx_t* h, double* csr_vals, const int64_t* slots, int64_t n, const double* buf) {
  if (!h) return HMX_ERR_ARG;
  if (n < 0 || (n > 0 && (!csr_vals || !slots || !buf))) return fail(h, HMX_ERR_ARG, "hmx_halo_unpack: null buffer");
  if (n == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  hmx_halo_unpack<<<grid_1d(n, 256, h->info[6]), 256, 0, h->stream>>>(n, (const long long*)slots, csr_vals, buf);
  HMX_CUDA(h, cudaGetLastError());
  return HMX_OK;
}

namespace {
// ncclAllReduce of the NCCL library the HOST already uses (the communicator is the host's): looked up among the loaded
// objects first, then libnccl.so.2 by name.  libhmx.so has no link-time dependency on NCCL.
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
nccl_allreduce_fn find_nccl_allreduce(std::string& why) {
  static nccl_allreduce_fn fn = nullptr;
  if (fn) return fn;
  void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
  if (!sym) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (lib) sym = dlsym(lib, "ncclAllReduce");
  }
  if (!sym) why = "ncclAllReduce not found: load the NCCL library the communicator belongs to before calling hmx_halo_sum_dev";
  fn = reinterpret_cast<nccl_allreduce_fn>(sym);
  return fn;
}
}  // namespace

int hmx_halo_sum_dev(hmx_t* h, void* nccl_comm, double* csr_vals, const int64_t* slots, int64_t n, double* buf) {
  if (!h) return HMX_ERR_ARG;
  if (!nccl_comm) return fail(h, HMX_ERR_ARG, "hmx_halo_sum: null communicator");
  if (n < 0 || (n > 0 && (!csr_vals || !slots || !buf))) return fail(h, HMX_ERR_ARG, "hmx_halo_sum: null buffer");
  std::string why;
  nccl_allreduce_fn allreduce = find_nccl_allreduce(why);
  if (!allreduce) return fail(h, HMX_ERR_KERNEL, "%s", why.c_str());
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  // every rank must enter the collective, also one that shares no slot (n = 0 is a zero-length all-reduce)
  if (n > 0) {
    hmx_halo_pack<<<grid_1d(n, 256, h->info[6]), 256, 0, h->stream>>>(n, (const long long*)slots, csr_vals, buf);
    HMX_CUDA(h, cudaGetLastError());
  }
  const int nccl_float64 = 8, nccl_sum = 0;  // ncclDataType_t / ncclRedOp_t of nccl.h (stable since NCCL 2.0)
  const int rc = allreduce(buf, buf, (size_t)n, nccl_float64, nccl_sum, nccl_comm, h->stream);
  if (rc != 0) return fail(h, HMX_ERR_CUDA, "ncclAllReduce failed with ncclResult_t %d", rc);
  if (n > 0) {
    hmx_halo_unpack<<<grid_1d(n, 256, h->info[6]), 256, 0, h->stream>>>(n, (const long long*)slots, csr_vals, buf);
    HMX_CUDA(h, cudaGetLastError());
  }
  return HMX_OK;
}

int hmx_macro_elements_dev(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, const double* node_xyz, int32_t nq,
                           const double* weights, const double* A_pts, double* S_loc) {
  if (!h) return HMX_ERR_ARG;
  if (n_cells < 0 || nq < 1 || !weights || (n_cells > 0 && (!cell_nodes || !node_xyz || !A_pts || !S_loc)))
    return fail(h, HMX_ERR_ARG, "hmx_macro_elements: null buffer or empty rule");
  if (n_cells == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  HMX_CUDA(h, h->l_qw.reserve((size_t)nq * sizeof(double)));
  HMX_CUDA(h, cudaMemcpyAsync(h->l_qw.p, weights, (size_t)nq * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const int g = grid_1d(n_cells, 256, h->info[6]);
  const double* wq = h->l_qw.as<double>();
  if (h->dim == 2 && h->kind == HMX_POISSON)
    hmx_macro_elements<2, 0><<<g, 256, 0, h->stream>>>(n_cells, cell_nodes, node_xyz, nq, wq, A_pts, S_loc);
  else if (h->dim == 3 && h->kind == HMX_POISSON)
    hmx_macro_elements<3, 0><<<g, 256, 0, h->stream>>>(n_cells, cell_nodes, node_xyz, nq, wq, A_pts, S_loc);
  else if (h->dim == 2)
    hmx_macro_elements<2, 1><<<g, 256, 0, h->stream>>>(n_cells, cell_nodes, node_xyz, nq, wq, A_pts, S_loc);
  else
    hmx_macro_elements<3, 1><<<g, 256, 0, h->stream>>>(n_cells, cell_nodes, node_xyz, nq, wq, A_pts, S_loc);
  HMX_CUDA(h, cudaGetLastError());
  return HMX_OK;
}

int hmx_macro_load_dev(hmx_t* h, const void* image, size_t image_size, int64_t n_cells, const int32_t* cell_nodes,
                       const double* node_xyz, int32_t nq, const double* qp, const double* qw, double* Fe) {
  if (!h) return HMX_ERR_ARG;
  if (!image || image_size == 0) return fail(h, HMX_ERR_KERNEL, "hmx_macro_load: no kernel image (the CUDA path has no fallback)");
  if (n_cells < 0 || nq < 1 || !qp || !qw || (n_cells > 0 && (!cell_nodes || !node_xyz || !Fe)))
    return fail(h, HMX_ERR_ARG, "hmx_macro_load: null buffer or empty quadrature rule");
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  Driver& drv = driver();
  unsigned long long hash = 1469598103934665603ULL;  // FNV-1a of the image: a new f loads a new module
  for (size_t i = 0; i < image_size; ++i) hash = (hash ^ static_cast<const unsigned char*>(image)[i]) * 1099511628211ULL;
  if (!h->load_fn || hash != h->load_hash) {
    if (h->load_module) {
      HMX_CUDA(h, cudaStreamSynchronize(h->stream));
      drv.ModuleUnload(h->load_module);
      h->load_module = nullptr;
      h->load_fn = nullptr;
    }
    HMX_CU(h, drv.ModuleLoadData(&h->load_module, image));
    HMX_CU(h, drv.ModuleGetFunction(&h->load_fn, h->load_module, "hmx_load"));
    CUdeviceptr gp = 0;
    size_t gs = 0;
    HMX_CU(h, drv.ModuleGetGlobal(&gp, &gs, h->load_module, "hmx_load_info"));
    HMX_CUDA(h, cudaMemcpy(h->load_info, (const void*)gp, sizeof h->load_info, cudaMemcpyDeviceToHost));
    h->load_hash = hash;
  }
  if (h->load_info[0] != h->dim) return fail(h, HMX_ERR_KERNEL, "load kernel was built for dim=%d, the solver has dim=%d", h->load_info[0], h->dim);
  if (n_cells == 0) return HMX_OK;
  HMX_CUDA(h, h->l_qp.reserve((size_t)nq * h->dim * sizeof(double)));
  HMX_CUDA(h, h->l_qw.reserve((size_t)nq * sizeof(double)));
  HMX_CUDA(h, cudaMemcpyAsync(h->l_qp.p, qp, (size_t)nq * h->dim * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  HMX_CUDA(h, cudaMemcpyAsync(h->l_qw.p, qw, (size_t)nq * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  struct {
    long long n_cells;
    const int* cell_nodes;
    const double* node_xyz;
    const double* qp;
    const double* qw;
    double* Fe;
    int nq;
  } P = {n_cells, cell_nodes, node_xyz, h->l_qp.as<double>(), h->l_qw.as<double>(), Fe, nq};
  void* args[] = {&P};
  const int threads = 128;
  HMX_CU(h, drv.LaunchKernel(h->load_fn, (unsigned)grid_1d(n_cells, threads, h->info[6]), 1, 1, threads, 1, 1, 0, (CUstream)h->stream, args,
                             nullptr));
  return HMX_OK;
}

int hmx_macro_lift_dev(hmx_t* h, int64_t n_dofs, const int64_t* indptr, const int32_t* indices, double* csr_vals,
                       const int8_t* bc_mask, const double* bc_values, double* b) {
  if (!h) return HMX_ERR_ARG;
  if (n_dofs < 0 || (n_dofs > 0 && (!indptr || !indices || !csr_vals || !bc_mask || !bc_values || !b)))
    return fail(h, HMX_ERR_ARG, "hmx_macro_lift: null buffer");
  if (n_dofs == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  hmx_lift<<<grid_1d(n_dofs, 256, h->info[6]), 256, 0, h->stream>>>(n_dofs, (const long long*)indptr, indices, csr_vals,
                                                                     (const signed char*)bc_mask, bc_values, b);
  HMX_CUDA(h, cudaGetLastError());
  return HMX_OK;
}

int hmx_macro_pcg_dev(hmx_t* h, int64_t n_dofs, const int64_t* indptr, const int32_t* indices, const double* csr_vals,
                      const double* b, double* x, double rtol, double atol, int32_t max_it, int32_t* iters, double* resid) {
  if (!h) return HMX_ERR_ARG;
  if (n_dofs < 0 || (n_dofs > 0 && (!indptr || !indices || !csr_vals || !b || !x)))
    return fail(h, HMX_ERR_ARG, "hmx_macro_pcg: null buffer");
  if (iters) *iters = 0;
  if (resid) *resid = 0.0;
  if (n_dofs == 0) return HMX_OK;
  DeviceGuard guard(h->device);
  HMX_CUDA(h, guard.status);
  const long long n = n_dofs;
  const int g = grid_1d(n, 256, h->info[6]);
  HMX_CUDA(h, h->m_work.reserve((4 * (size_t)n + 8 + 3 * (size_t)g) * sizeof(double)));
  double* r = h->m_work.as<double>();
  double *p = r + n, *y = p + n, *dinv = y + n, *sc = dinv + n;
  double *bp0 = sc + 8, *bp_pap = bp0 + g, *bp_rz = bp_pap + g;  // block partials of the three dot products
  const long long* ptr = (const long long*)indptr;
  HMX_CUDA(h, cudaMemsetAsync(sc, 0, 8 * sizeof(double), h->stream));
  hmx_pcg_init<<<g, 256, 0, h->stream>>>(n, ptr, indices, csr_vals, b, x, r, p, dinv, bp0);
  hmx_pcg_start<<<1, 256, 0, h->stream>>>(bp0, g, sc);
  double hs[4] = {0, 0, 0, 0};
  HMX_CUDA(h, cudaMemcpyAsync(hs, sc, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
  HMX_CUDA(h, cudaStreamSynchronize(h->stream));
  const double rz0 = hs[3];
  const double tol2 = std::max(rtol * rtol * rz0, atol * atol);
  double rz = rz0;
  int it = 0;
  if (max_it <= 0) max_it = 10000;
  while (rz > tol2 && it < max_it) {
    const int burst = std::min(16, max_it - it);  // the host looks at the residual every 16 iterations
    for (int k = 0; k < burst; ++k) {
      hmx_pcg_spmv<<<g, 256, 0, h->stream>>>(n, ptr, indices, csr_vals, p, y, bp_pap);
      hmx_pcg_update<<<g, 256, 0, h->stream>>>(n, x, r, p, y, dinv, sc, bp_pap, bp_rz);
      hmx_pcg_direction<<<g, 256, 0, h->stream>>>(n, r, p, dinv, sc, bp_rz);
      hmx_pcg_rotate<<<1, 256, 0, h->stream>>>(sc, bp_rz, g);
    }
    it += burst;
    HMX_CUDA(h, cudaMemcpyAsync(hs, sc, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    HMX_CUDA(h, cudaStreamSynchronize(h->stream));
    rz = hs[0];
    if (!(rz == rz)) return fail(h, HMX_ERR_CUDA, "hmx_macro_pcg: residual is not a number (matrix not positive definite?)");
  }
  HMX_CUDA(h, cudaGetLastError());
  if (iters) *iters = it;
  if (resid) *resid = rz0 > 0.0 ? std::sqrt(rz / rz0) : 0.0;
  return HMX_OK;
}

int hmx_measure_peaks(int32_t device, double* fp64_tflops, double* copy_gbs) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(nullptr, HMX_ERR_CUDA, "no CUDA device %d", device);
  }
  DeviceGuard guard(device);
  HMX_CUDA(nullptr, guard.status);
  cudaDeviceProp prop;
  HMX_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  cudaEvent_t e0, e1;
  HMX_CUDA(nullptr, cudaEventCreate(&e0));
  HMX_CUDA(nullptr, cudaEventCreate(&e1));
  if (fp64_tflops) {
    const int blocks = prop.multiProcessorCount * 4, threads = 256, iters = 4096;
    double* out = nullptr;
    HMX_CUDA(nullptr, cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      hmx_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      HMX_CUDA(nullptr, cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(out);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    *fp64_tflops = flops / (best * 1e-3) / 1e12;
  }
  if (copy_gbs) {
    const long long n = 1LL << 26;  // 64 Mi double2 = 1 GiB per buffer
    double2 *a = nullptr, *b = nullptr;
    HMX_CUDA(nullptr, cudaMalloc(&a, n * sizeof(double2)));
    HMX_CUDA(nullptr, cudaMalloc(&b, n * sizeof(double2)));
    cudaMemset(a, 0, n * sizeof(double2));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      hmx_copy<<<prop.multiProcessorCount * 16, 256>>>(a, b, n);
      cudaEventRecord(e1);
      HMX_CUDA(nullptr, cudaEventSynchronize(e1));
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(a);
    cudaFree(b);
    *copy_gbs = 2.0 * n * sizeof(double2) / (best * 1e-3) / 1e9;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return HMX_OK;
}

}  // extern "C"
