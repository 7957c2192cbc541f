// TEST INFRASTRUCTURE ONLY: the block reductions of csrc/hmx_cell_common.cuh (transposing warp butterfly, warp partials,
// CTA totals) run on the fiber emulation and compared with plain sums.  Built and run by tests/test_block_reductions.py.
#include <cmath>
#include <cstdio>

#include "hmx_cell_common.cuh"
#include "emu_runtime.h"

namespace {
constexpr int NT = 128, NW = NT / 32;
double g_err = 0.0;
int g_bad = 0;

inline double value(int t, int k) { return std::sin(0.37 * t + 1.3 * k) + 0.01 * k; }

template <int NV>
void check(double* buf) {
  using namespace hmx;
  const int t = tid();
  double v[NV];
  for (int k = 0; k < NV; ++k) v[k] = value(t, k);
  block_sum<NV, NW>(v, buf);
  for (int k = 0; k < NV; ++k) {
    double ref = 0.0, mag = 0.0;
    for (int u = 0; u < NT; ++u) {
      ref += value(u, k);
      mag += std::fabs(value(u, k));
    }
    const double e = std::fabs(v[k] - ref) / mag;
    if (e > g_err) g_err = e;
    if (!(e < 1e-14)) ++g_bad;
  }
  sync();  // the buffer is reused by the next size
  // the warp totals alone (block_partials): partial of warp w and value k in buf[w * NV + k]
  double w[NV];
  for (int k = 0; k < NV; ++k) w[k] = value(t, k);
  block_partials<NV, NW>(w, buf);
  if (t < NW * NV) {
    const int wp = t / NV, k = t % NV;
    double ref = 0.0, mag = 0.0;
    for (int u = 32 * wp; u < 32 * wp + 32; ++u) {
      ref += value(u, k);
      mag += std::fabs(value(u, k));
    }
    const double e = std::fabs(buf[wp * NV + k] - ref) / mag;
    if (e > g_err) g_err = e;
    if (!(e < 1e-14)) ++g_bad;
  }
  sync();
}

void body(void*) {
  double* buf = hmx::dyn_smem();
  check<1>(buf);
  check<2>(buf);
  check<3>(buf);
  check<4>(buf);
  check<5>(buf);
  check<8>(buf);
  check<9>(buf);
  check<12>(buf);
  check<16>(buf);
  check<18>(buf);
  check<31>(buf);
  check<32>(buf);
}
}  // namespace

int main() {
  hmx::emu::run_grid(2, NT, NW * 32 + 8, body, nullptr, 1);
  std::printf("max relative error %.3e, failures %d\n", g_err, g_bad);
  return g_bad == 0 ? 0 : 1;
}
