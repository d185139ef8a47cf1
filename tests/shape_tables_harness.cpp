// Test harness (tests/ only): C wrappers around the product's host-side shape-table builder so that the conservative
// direction-cell tables can be checked on the CPU against the oracle's exact radius evaluation.
#include <cstring>
#include <string>
#include "../lammps-spherharm_b200/csrc/shape_tables.h"

extern "C" {
void *sth_build(int lmax, const double *a, const double *b, double density, int nt, int np, char *err, int errlen) {
  auto *t = new shgpu::ShapeTables();
  std::string e = shgpu::build_shape_tables(lmax, a, b, density, nt, np, *t);
  if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); delete t; return nullptr; }
  return t;
}
void sth_free(void *p) { delete static_cast<shgpu::ShapeTables *>(p); }
int sth_cube_n(void *p) { return static_cast<shgpu::ShapeTables *>(p)->cube_n; }
void sth_scalars(void *p, double *out) {   // rmax, rmin, cache_delta, volume
  auto *t = static_cast<shgpu::ShapeTables *>(p);
  out[0] = t->rmax; out[1] = t->rmin; out[2] = t->cache_delta; out[3] = t->volume;
}
void sth_cube(void *p, float *narrow, float *wide) {
  auto *t = static_cast<shgpu::ShapeTables *>(p);
  std::memcpy(narrow, t->cube_bound2.data(), t->cube_bound2.size() * sizeof(float));
  std::memcpy(wide, t->cube_wide2.data(), t->cube_wide2.size() * sizeof(float));
}
void sth_nodes(void *p, double *pts, double *nds) {
  auto *t = static_cast<shgpu::ShapeTables *>(p);
  for (int k = 0; k < t->nq; k++) for (int d = 0; d < 3; d++) { pts[3 * k + d] = t->node_p[d][k]; nds[3 * k + d] = t->node_n[d][k]; }
}
}
