// Translation unit of one specialised micro cell kernel.  Compiled once per
// (coefficient program, micro mesh size, block size) by hommx_b200/native.py:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -cubin
//        -DHMX_COEFF_FILE="<generated struct HMX_COEFF>" -DHMX_KIND=k -DHMX_NM=n -DHMX_NT=threads
// The same file, compiled by g++ with -DHMX_EMULATE, is the CPU emulation used by the
// `not gpu` tests (tests/cpu_emu) -- test infrastructure, never shipped.
#ifndef HMX_VARIANT
#define HMX_VARIANT 0  // elasticity: 0 = matrix-free element kernel (PCG), 3 = dense Cholesky, 4 = cluster-resident stencil;
#endif                 // both kinds: 5 = element-list kernel for general periodic micro meshes (the mesh is data: HMX_NM = 0)
#ifndef HMX_CLUSTER
#define HMX_CLUSTER 1  // CTAs per thread-block cluster (variant 4: the cell is split into z-slabs over the cluster)
#endif
#ifndef HMX_TPN
#define HMX_TPN 2  // variant 4: threads per node
#endif
// (variants 1 and 2 -- the assembled operator streamed from L2, barrier- or TMA-staged -- were measured slower than
// variant 0 and live in experiments/assembled_operator/; they build only with -DHMX_EXPERIMENTAL_VARIANTS and that
// directory on the include path)
#if HMX_KIND == 0
#include "hmx_cell_poisson.cuh"
#else
#include "hmx_cell_elasticity.cuh"
#ifdef HMX_EXPERIMENTAL_VARIANTS
#include "hmx_cell_elasticity_asm.cuh"
#include "hmx_cell_elasticity_tma.cuh"
#endif
#include "hmx_cell_dense.cuh"
#if HMX_VARIANT == 4
#include "hmx_cell_cluster.cuh"
#endif
#endif
#if HMX_VARIANT == 5
#include "hmx_cell_generic.cuh"
#endif
#include HMX_COEFF_FILE

#ifndef HMX_MINB
#define HMX_MINB 1
#endif
#ifndef HMX_VGLOB
#define HMX_VGLOB 0  // large cells: 1 = elasticity p, K p / Poisson atoms live in the L2 scratch instead of shared memory
#endif
#ifndef HMX_COLL
#define HMX_COLL 0  // bit mask of collapsed micro axes (exact symmetry reduction, hmx_cell_common.cuh)
#endif

namespace {
#if HMX_VARIANT == 5
using Layout = hmx::GenericLayout<HMX_COEFF, HMX_NT>;
#elif HMX_KIND == 0
using Layout = hmx::PoissonLayout<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>;
#elif HMX_VARIANT == 1 && defined(HMX_EXPERIMENTAL_VARIANTS)
using Layout = hmx::ElasticityAsmLayout<HMX_COEFF, HMX_NM, HMX_NT>;
#elif HMX_VARIANT == 2 && defined(HMX_EXPERIMENTAL_VARIANTS)
using Layout = hmx::ElasticityTmaLayout<HMX_COEFF, HMX_NM, HMX_NT>;
#elif HMX_VARIANT == 3
using Layout = hmx::DenseLayout<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL>;
#elif HMX_VARIANT == 4
using Layout = hmx::ClusterLayout<HMX_COEFF, HMX_NM, HMX_NT, HMX_CLUSTER, HMX_TPN>;
#else
using Layout = hmx::ElasticityLayout<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>;
#endif
#ifndef HMX_EXPERIMENTAL_VARIANTS
static_assert(HMX_KIND == 0 || HMX_VARIANT == 0 || HMX_VARIANT == 3 || HMX_VARIANT == 4 || HMX_VARIANT == 5,
              "elasticity kernel variants: 0 (matrix-free PCG), 3 (dense Cholesky), 4 (cluster-resident stencil), 5 (element list)");
#endif
static_assert(HMX_VARIANT == 0 || HMX_VARIANT == 3 || HMX_COLL == 0, "the assembled / element-list variants have no collapsed form");
static_assert(HMX_VARIANT == 4 || HMX_CLUSTER == 1, "only variant 4 is launched as thread-block clusters");
static_assert(HMX_COEFF::KIND == HMX_KIND, "coefficient program / kernel kind mismatch");
constexpr int kSmemBytes = Layout::total * 8;
constexpr int kScratch = Layout::scratch_doubles;
}  // namespace

#ifndef HMX_EMULATE
#if HMX_VARIANT == 4
extern "C" HMX_GLOBAL_CLUSTER(HMX_NT, HMX_CLUSTER) hmx_cell(const hmx::CellParams P) {
  hmx::elasticity_cluster_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_CLUSTER, HMX_TPN>(P);
}
#else
extern "C" HMX_GLOBAL(HMX_NT, HMX_MINB) hmx_cell(const hmx::CellParams P) {
#if HMX_VARIANT == 5
  hmx::generic_cell_body<HMX_COEFF, HMX_NT>(P);
#elif HMX_KIND == 0
  hmx::poisson_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>(P);
#elif HMX_VARIANT == 1 && defined(HMX_EXPERIMENTAL_VARIANTS)
  hmx::elasticity_asm_cell_body<HMX_COEFF, HMX_NM, HMX_NT>(P);
#elif HMX_VARIANT == 2 && defined(HMX_EXPERIMENTAL_VARIANTS)
  hmx::elasticity_tma_cell_body<HMX_COEFF, HMX_NM, HMX_NT>(P);
#elif HMX_VARIANT == 3
  hmx::elasticity_dense_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL>(P);
#else
  hmx::elasticity_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>(P);
#endif
}
#endif
// 0 smem bytes, 1 threads, 2 nrhs, 3 dim, 4 kind, 5 n_micro, 6 scratch doubles per CTA, 7 quadrature degree,
// 8 CTAs per cluster (1: ordinary launch; the grid must be a multiple of it), 9 atoms of the coefficient program,
// 10-11 reserved
extern "C" __device__ const int hmx_info[12] = {kSmemBytes, HMX_NT, Layout::NRHS, HMX_COEFF::DIM, HMX_KIND, HMX_NM,
                                                kScratch, HMX_COEFF::QDEG, HMX_CLUSTER, HMX_COEFF::NATOMS, 0, 0};
#else
#include "emu_runtime.h"
static void emu_body(void* arg) {
  const hmx::CellParams& P = *static_cast<const hmx::CellParams*>(arg);
#if HMX_VARIANT == 5
  hmx::generic_cell_body<HMX_COEFF, HMX_NT>(P);
#elif HMX_KIND == 0
  hmx::poisson_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>(P);
#elif HMX_VARIANT == 1 && defined(HMX_EXPERIMENTAL_VARIANTS)
  hmx::elasticity_asm_cell_body<HMX_COEFF, HMX_NM, HMX_NT>(P);
#elif HMX_VARIANT == 2 && defined(HMX_EXPERIMENTAL_VARIANTS)
  hmx::elasticity_tma_cell_body<HMX_COEFF, HMX_NM, HMX_NT>(P);
#elif HMX_VARIANT == 3
  hmx::elasticity_dense_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL>(P);
#elif HMX_VARIANT == 4
  hmx::elasticity_cluster_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_CLUSTER, HMX_TPN>(P);
#else
  hmx::elasticity_cell_body<HMX_COEFF, HMX_NM, HMX_NT, HMX_COLL, HMX_VGLOB>(P);
#endif
}
extern "C" int hmx_emu_cluster() { return HMX_CLUSTER; }
extern "C" void hmx_emu_info(int* out) {
  const int v[8] = {kSmemBytes, HMX_NT, Layout::NRHS, HMX_COEFF::DIM, HMX_KIND, HMX_NM, kScratch, HMX_COEFF::QDEG};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
}
extern "C" void hmx_emu_launch(hmx::CellParams* P, int grid, int host_threads) {
  hmx::emu::run_grid(grid, HMX_NT, (size_t)Layout::total, emu_body, P, host_threads, HMX_CLUSTER);
}
#endif
