/* sh_oracle.h — CPU FP64 oracle for the SPHERHARM contact hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lammps-spherharm_b200/)
 * includes, links or executes this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: /root/reference holds only README.md:1-3 (one heading,
 * "SPHERHARM Package to simulate complex shaped granular particles"); there is
 * no pair_spherharm source, test or golden vector to follow or to check against.
 * This file therefore restates the algorithm BASELINE.json:5 (north_star)
 * describes, made precise in SURVEY.md Appendix A (A.1-A.8) and DESIGN.md §3.
 * It is pinned only by analytic known-answer tests (tests/test_oracle_kat.py).
 */
#ifndef SH_ORACLE_H
#define SH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(void);
void orc_destroy(orc_ctx *c);
const char *orc_last_error(const orc_ctx *c);

int orc_set_box(orc_ctx *c, const double lo[3], const double hi[3], const int periodic[3]);
int orc_set_quadrature(orc_ctx *c, int n_theta, int n_phi);
/* a_lm / b_lm: (lmax+1)(lmax+2)/2 doubles, index l(l+1)/2+m, real orthonormal SH
 * without Condon-Shortley phase. */
int orc_add_shape(orc_ctx *c, int lmax, const double *a_lm, const double *b_lm, double density,
                  int *shape_id_out);
int orc_get_shape_props(const orc_ctx *c, int shape, double *volume, double com[3],
                        double inertia[3], double quat_principal[4], double *rmax, double *rmin);
int orc_get_nodes(const orc_ctx *c, int shape, double *p /*nq*3*/, double *nds /*nq*3*/);
int orc_set_atoms(orc_ctx *c, int64_t n, const int64_t *tag, const int *shape, const double *x,
                  const double *v, const double *quat, const double *angmom);
int orc_pair_coeff(orc_ctx *c, int shape_i, int shape_j, double k, double exponent);
int orc_add_wall(orc_ctx *c, const double point[3], const double normal[3], double k,
                 double exponent);
int orc_set_gravity(orc_ctx *c, const double g[3]);
int orc_set_neighbor(orc_ctx *c, double skin, int every, int check);
int orc_set_timestep(orc_ctx *c, double dt);
int orc_set_damping(orc_ctx *c, double gamma_lin, double gamma_rot);
/* dissipative contact terms (A.5b): viscous normal damping gamma_n, tangential gamma_t capped by Coulomb mu */
int orc_pair_dissipation(orc_ctx *c, int shape_i, int shape_j, double gamma_n, double gamma_t, double mu);
/* pressure-tensor sums: kinetic = sum m v v (owned atoms), virial = sum_pairs (x_i - x_j) F_i (row-major 3x3) */
int orc_get_stress(const orc_ctx *c, double virial[9], double kinetic[9]);
/* Lees-Edwards shear: flow along x, gradient along y, rate = dvx/dy; after orc_set_box */
int orc_set_shear(orc_ctx *c, double rate);
int orc_set_threads(orc_ctx *c, int nthreads);
int orc_compute_forces(orc_ctx *c);
int orc_run(orc_ctx *c, int64_t nsteps);
/* multi-rank test support (mirrors sh_set_ghost_count / sh_step_begin / sh_step_end / sh_pack_atoms /
 * sh_unpack_ghosts; pointers are HOST pointers) */
int orc_set_ghost_count(orc_ctx *c, int64_t nghost);
int orc_step_begin(orc_ctx *c, int *rebuild_wanted);
int orc_step_end(orc_ctx *c, int rebuild);
int orc_pack_atoms(const orc_ctx *c, int64_t m, const int *idx, const double *shift, double *out);
int orc_unpack_ghosts(orc_ctx *c, int64_t first, int64_t m, const double *in);
int orc_get_atoms(const orc_ctx *c, int64_t n, double *x, double *v, double *quat, double *angmom,
                  double *f, double *torque);
int orc_get_pairs(const orc_ctx *c, int64_t cap, int64_t *npairs, int64_t *tag_i, int64_t *tag_j,
                  double *V, double *F, double *tau_i, double *tau_j, double *centroid);
int orc_get_counters(const orc_ctx *c, int64_t *pair_evals, int64_t *nodes_transformed,
                     int64_t *nodes_evaluated, int64_t *nodes_inside);
int orc_get_energy(const orc_ctx *c, double *ke_trans, double *ke_rot, double *e_contact);

/* KAT hooks */
int orc_sh_radius(const orc_ctx *c, int shape, int64_t n, const double *dirs /*n*3*/, double *r);
int orc_legendre_norm(int lmax, double x, double *P /* (lmax+1)(lmax+2)/2, index l(l+1)/2+m */);
int orc_gauss_legendre(int n, double *x, double *w);
/* project r(theta,phi) of an axis-aligned ellipsoid (semi-axes a,b,c) onto lmax */
int orc_project_ellipsoid(int lmax, double a, double b, double c, int n_theta, int n_phi,
                          double *a_lm, double *b_lm);

#ifdef __cplusplus
}
#endif
#endif
