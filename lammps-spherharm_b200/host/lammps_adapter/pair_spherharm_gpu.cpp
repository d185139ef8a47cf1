/* UNTESTED adapter (see README.md in this directory).
 * Pair::compute offload: every step push x / quat of the nlocal+nghost atoms, run the device pair phase, read f / torque
 * back and add them to atom->f / atom->torque.  The neighbor list is built on the device (bounding-sphere binned build), so
 * no LAMMPS neighbor request is made; ghost atoms come from LAMMPS' own CommBrick (newton off). */
#include "pair_spherharm_gpu.h"

#include "atom.h"
#include "atom_vec_spherharm.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neighbor.h"

using namespace LAMMPS_NS;

PairSpherharmGPU::PairSpherharmGPU(LAMMPS *lmp) : Pair(lmp), h(nullptr), avec(nullptr), kcoef(nullptr), expo(nullptr),
                                                    fbuf(nullptr), tbuf(nullptr), nmax(0), uploaded(0) {
  single_enable = 0;
  restartinfo = 0;
}

PairSpherharmGPU::~PairSpherharmGPU() {
  if (h) sh_destroy(h);
  if (allocated) { memory->destroy(setflag); memory->destroy(cutsq); memory->destroy(kcoef); memory->destroy(expo); }
  memory->destroy(fbuf); memory->destroy(tbuf);
}

void PairSpherharmGPU::allocate() {
  allocated = 1;
  const int n = atom->ntypes;
  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
  memory->create(kcoef, n + 1, n + 1, "pair:kcoef");
  memory->create(expo, n + 1, n + 1, "pair:expo");
  for (int i = 1; i <= n; i++) for (int j = i; j <= n; j++) setflag[i][j] = 0;
}

void PairSpherharmGPU::settings(int narg, char ** /*arg*/) {
  if (narg != 0) error->all(FLERR, "Illegal pair_style command");
}

void PairSpherharmGPU::coeff(int narg, char **arg) {      // pair_coeff i j k exponent
  if (narg != 4) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  utils::bounds(FLERR, arg[0], 1, atom->ntypes, ilo, ihi, error);
  utils::bounds(FLERR, arg[1], 1, atom->ntypes, jlo, jhi, error);
  const double k = utils::numeric(FLERR, arg[2], false, lmp), e = utils::numeric(FLERR, arg[3], false, lmp);
  int count = 0;
  for (int i = ilo; i <= ihi; i++)
    for (int j = MAX(jlo, i); j <= jhi; j++) { kcoef[i][j] = k; expo[i][j] = e; setflag[i][j] = 1; count++; }
  if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
}

void PairSpherharmGPU::init_style() {
  avec = dynamic_cast<AtomVecSpherharm *>(atom->style_match("spherharm"));
  if (!avec) error->all(FLERR, "Pair spherharm/gpu requires atom style spherharm");
  if (force->newton_pair) error->all(FLERR, "Pair spherharm/gpu requires newton pair off");
  if (h) { sh_destroy(h); h = nullptr; }
  if (sh_create(&h, comm->me % /*GPUs per node*/ 8) != 0) error->all(FLERR, "No usable CUDA device for pair spherharm/gpu (no CPU fallback)");
  // this rank's sub-domain is non-periodic from the engine's point of view: periodic images arrive as LAMMPS ghosts
  int noper[3] = {0, 0, 0};
  sh_set_box(h, domain->sublo, domain->subhi, noper);
  sh_set_quadrature(h, avec->get_n_theta(), avec->get_n_phi());                 // placeholder accessors
  for (int s = 0; s < avec->get_nshapes(); s++)
    if (sh_add_shape(h, avec->get_lmax(), avec->get_alm(s), avec->get_blm(s), avec->get_density(s), nullptr) != 0)
      error->all(FLERR, sh_last_error(h));
  for (int i = 1; i <= atom->ntypes; i++)
    for (int j = i; j <= atom->ntypes; j++)
      if (setflag[i][j] && sh_pair_coeff(h, i - 1, j - 1, kcoef[i][j], expo[i][j]) != 0) error->all(FLERR, sh_last_error(h));
  sh_set_neighbor(h, neighbor->skin, 1, 1);
  uploaded = 0;
}

double PairSpherharmGPU::init_one(int i, int j) {
  if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
  kcoef[j][i] = kcoef[i][j]; expo[j][i] = expo[i][j];
  return avec->get_rmax(i - 1) + avec->get_rmax(j - 1);      // bounding-sphere cutoff: ghosts must cover it
}

void PairSpherharmGPU::upload_atoms() {      // after every re-neighboring (atoms migrated / ghosts re-created)
  const int nall = atom->nlocal + atom->nghost;
  std::vector<int> shape(nall);
  for (int i = 0; i < nall; i++) shape[i] = atom->type[i] - 1;
  if (sh_set_atoms(h, nall, atom->tag, shape.data(), &atom->x[0][0], &atom->v[0][0], &avec->quat[0][0], &atom->angmom[0][0]) != 0 ||
      sh_set_ghost_count(h, atom->nghost) != 0)
    error->one(FLERR, sh_last_error(h));
  if (nall > nmax) { nmax = nall; memory->grow(fbuf, 3 * nmax, "pair:fbuf"); memory->grow(tbuf, 3 * nmax, "pair:tbuf"); }
  uploaded = 1;
}

void PairSpherharmGPU::compute(int eflag, int vflag) {
  ev_init(eflag, vflag);
  const int nall = atom->nlocal + atom->nghost;
  if (!uploaded || neighbor->ago == 0) upload_atoms();
  else if (sh_put_state(h, nall, &atom->x[0][0], nullptr, &avec->quat[0][0], nullptr) != 0) error->one(FLERR, sh_last_error(h));
  if (sh_compute_forces(h) != 0) error->one(FLERR, sh_last_error(h));
  if (sh_get_forces(h, nall, fbuf, tbuf) != 0) error->one(FLERR, sh_last_error(h));
  double **f = atom->f, **torque = atom->torque;
  for (int i = 0; i < atom->nlocal; i++)
    for (int d = 0; d < 3; d++) { f[i][d] += fbuf[3 * i + d]; torque[i][d] += tbuf[3 * i + d]; }
  if (eflag_global) { double kt, kr, ec; sh_get_energy(h, &kt, &kr, &ec); eng_vdwl += ec; }
}
