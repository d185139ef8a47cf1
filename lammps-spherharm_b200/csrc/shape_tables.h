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
  // Direction-cell tables over a cube map of directions (6 faces x cube_n x cube_n gnomonic cells), all RIGOROUS
  // (DESIGN §4.0): r is sampled on a gnomonic grid of spacing sample_step per face; between samples it is bounded with
  // the second-derivative bound h2_bound (Bernstein inequality per harmonic degree), so that
  //   cube_ub2[cell] >= r(d)^2 >= cube_lb2[cell]   for every direction d whose FP32 cell index is `cell`.
  // A node with rho^2 >= ub2 is outside and one with rho^2 <= lb2 is inside, without evaluating the series.
  int cube_n = 0;
  std::vector<float> cube_ub2, cube_lb2;
  // candidate-cache tables, one per margin level: (max of r over every direction a node direction can drift to before
  // the cache is rebuilt + cache_delta[level])^2; cache_delta = node displacement margin of that level
  std::vector<float> cube_wide2[4];
  double cache_delta[4] = {0, 0, 0, 0};
  double h1_bound = 0, h2_bound = 0;     // sup |dr/dt|, sup |d2r/dt2| along great circles (t = arc length), rigorous
  double sample_step = 0, sample_pad = 0;   // gnomonic sample spacing and the interpolation pad h2 step^2/4 (+ border term)
  double r_sup = 0, r_inf = 0;           // proven bounds of r over the sphere (max ub, min lb); rmax/rmin are checked against them
  std::vector<double> row_x;             // cos(theta) of the Gauss-Legendre rows (node k = row*n_phi+col)
  double density = 1, volume = 0, mass = 0;
  std::array<double, 3> com{}, inertia{};
  std::array<double, 4> quat_principal{};
  double Rp[3][3] = {};                  // principal frame -> shape frame
  double rmax = 0, rmin = 0;
};

// Returns "" on success, else an error message.
constexpr int SH_CACHE_LEVELS = 4;   // 0..2 chosen adaptively; 3 only for pairs that join a live cache (remap)
std::string build_shape_tables(int lmax, const double *a_lm, const double *b_lm, double density,
                               int n_theta, int n_phi, ShapeTables &out, int cube_n = 0);

void gauss_legendre_nodes(int n, std::vector<double> &x, std::vector<double> &w);
void legendre_normalised(int lmax, double x, std::vector<double> &P);
void rotation_from_quat(const double q[4], double R[3][3]);

}  // namespace shgpu
