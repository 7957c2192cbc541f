// Macro load vector on the device: element vectors  Fe[cell][a * BS + k] = int_T f_k(x) phi_a(x) dx  for every macro cell.
//
// Replaces the FFCx kernel behind `_assemble_vector_array(b_local.array_w, self._L, ...)` of the reference
// (hmm.py:445-450, L = inner(f(x), v) dx, hmm.py:131-133).  f is the generated `struct HMX_LOAD` (hommx_b200/codegen.py
// build_load_program); the quadrature rule of the UFL-estimated degree arrives as a table.  The element vectors are
// then summed into b by the same deterministic gather as the stiffness values (hmx_gather_csr in hmx_lib.cu): no
// atomics, fixed order.  HBM-bound streaming kernel, one thread per macro cell.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -cubin -DHMX_LOAD_FILE="<generated struct HMX_LOAD>"
#include HMX_LOAD_FILE

struct LoadParams {
  long long n_cells;
  const int* cell_nodes;   // [n_cells][D+1]
  const double* node_xyz;  // [n_nodes][3]
  const double* qp;        // [nq][D] points on the reference simplex
  const double* qw;        // [nq] weights (sum 1/D!)
  double* Fe;              // [n_cells][(D+1)*BS]
  int nq;
};

extern "C" __global__ void __launch_bounds__(128) hmx_load(const LoadParams P) {
  constexpr int D = HMX_LOAD::DIM, BS = HMX_LOAD::BS, NV = D + 1;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < P.n_cells; c += (long long)gridDim.x * blockDim.x) {
    double v[NV][3];
#pragma unroll
    for (int a = 0; a < NV; ++a) {
      const long long node = P.cell_nodes[c * NV + a];
#pragma unroll
      for (int k = 0; k < 3; ++k) v[a][k] = P.node_xyz[node * 3 + k];
    }
    double J[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int cc = 0; cc < D; ++cc) J[r][cc] = v[cc + 1][r] - v[0][r];
    double det;
    if (D == 2)
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    else
      det = J[0][0] * (J[1][1] * J[2 % D][2 % D] - J[1][2 % D] * J[2 % D][1]) - J[0][1] * (J[1][0] * J[2 % D][2 % D] - J[1][2 % D] * J[2 % D][0]) +
            J[0][2 % D] * (J[1][0] * J[2 % D][1] - J[1][1] * J[2 % D][0]);
    det = fabs(det);
    double acc[NV][BS];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int k = 0; k < BS; ++k) acc[a][k] = 0.0;
    for (int q = 0; q < P.nq; ++q) {
      double phi[NV], x[3] = {0.0, 0.0, 0.0}, f[BS];
      phi[0] = 1.0;
#pragma unroll
      for (int a = 1; a < NV; ++a) {
        phi[a] = P.qp[q * D + a - 1];
        phi[0] -= phi[a];
      }
#pragma unroll
      for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) x[k] += phi[a] * v[a][k];
      HMX_LOAD::eval(x, f);
      const double w = P.qw[q] * det;
#pragma unroll
      for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int k = 0; k < BS; ++k) acc[a][k] += w * phi[a] * f[k];
    }
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int k = 0; k < BS; ++k) P.Fe[c * (NV * BS) + a * BS + k] = acc[a][k];
  }
}
// 0 dim, 1 block size, 2 quadrature degree
extern "C" __device__ const int hmx_load_info[4] = {HMX_LOAD::DIM, HMX_LOAD::BS, HMX_LOAD::QDEG, 0};
