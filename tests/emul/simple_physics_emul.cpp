// TEST-ONLY host emulation of k_simple_physics: the kernel's per-column code (climt_b200/csrc/simple_physics_core.cuh) compiled for
// the host and stepped column by column over a NaN-poisoned workspace.
#include <vector>

#include "../../climt_b200/csrc/simple_physics_core.cuh"

extern "C" int emul_simple_physics_run(int ncol, int nlev, int order, double dtime, const cb200_simple_physics_params* p,
                                       const cb200_simple_physics_inputs* in, const cb200_simple_physics_outputs* out) {
  const cb::sp::Geo G{ncol, nlev, order ? 1 : 0};
  const double nan = 0.0 / 0.0;
  std::vector<double> work((size_t)2 * nlev * ncol, nan);   // whatever the code reads from it must have been written by it
  for (int c = 0; c < ncol; ++c) cb::sp::simple_physics_column(G, dtime, *p, *in, *out, work.data(), c);
  return 0;
}
