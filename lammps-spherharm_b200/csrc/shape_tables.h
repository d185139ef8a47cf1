// shape_tables.h — host-side per-shape tables of `atom_style spherharm`.
//
// Built once per shape at sh_add_shape time (the AtomVec::process_args step of the reference's
// atom style; reference source NOT IN MOUNT, see include/shgpu.h).  Everything here is plain
// FP64 host arithmetic compiled with -ffp-contract=off: the values are inputs to the node
// inside/outside decision, so they must not depend on compiler FMA contraction.
#pragma once
#include <array>
#include <string>
#include <vector>

namespace shgpu {

struct ShapeTables {
  int lmax = 0;
  int nterms = 0;              // (lmax+1)(lmax+2)/2
  int n_theta = 0, n_phi = 0, nq = 0;
  std::vector<double> a_raw, b_raw;      // index l(l+1)/2+m
  // folded recurrence, m-major: entry off(m)+(l-m)
  std::vector<double> Ap;                // x-multiplier of the three-term recurrence
  std::vector<double> ah, bh;            // coefficients with alpha_lm * c_m folded in
  // node table, SoA: px,py,pz, nx,ny,nz (oriented area elements n dS)
  std::vector<double> node_p[3], node_n[3];
  // conservative squared upper bound of r over the cells of a cube map of directions (6 x cube_n x cube_n):
  // a node with rho^2 >= bound2(cell(direction)) is certainly outside, no SH evaluation needed
  int cube_n = 0;
  std::vector<float> cube_bound2;
  // candidate-cache table: (sqrt(max of cube_bound2 over all cells within the angle a node direction can drift
  // before the cache is rebuilt) + cache_delta)^2; cache_delta = node displacement margin of the cache
  std::vector<float> cube_wide2;
  double cache_delta = 0;
  std::vector<double> row_x;             // cos(theta) of the Gauss-Legendre rows (node k = row*n_phi+col)
  double density = 1, volume = 0, mass = 0;
  std::array<double, 3> com{}, inertia{};
  std::array<double, 4> quat_principal{};
  double Rp[3][3] = {};                  // principal frame -> shape frame
  double rmax = 0, rmin = 0;
};

// Returns "" on success, else an error message.
std::string build_shape_tables(int lmax, const double *a_lm, const double *b_lm, double density,
                               int n_theta, int n_phi, ShapeTables &out);

void gauss_legendre_nodes(int n, std::vector<double> &x, std::vector<double> &w);
void legendre_normalised(int lmax, double x, std::vector<double> &P);
void rotation_from_quat(const double q[4], double R[3][3]);

}  // namespace shgpu
