/* UNTESTED adapter (see README.md in this directory): pair_style spherharm/gpu -> libshgpu.so */
#ifdef PAIR_CLASS
// clang-format off
PairStyle(spherharm/gpu,PairSpherharmGPU);
// clang-format on
#else
#ifndef LMP_PAIR_SPHERHARM_GPU_H
#define LMP_PAIR_SPHERHARM_GPU_H

#include "pair.h"

extern "C" {
#include "shgpu.h"
}

namespace LAMMPS_NS {

class PairSpherharmGPU : public Pair {
 public:
  PairSpherharmGPU(class LAMMPS *);
  ~PairSpherharmGPU() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;

 protected:
  sh_ctx *h;
  class AtomVecSpherharm *avec;
  double **kcoef, **expo;      // per type pair: stiffness, exponent
  double *fbuf, *tbuf;         // nall x 3 host buffers filled by sh_get_forces
  int nmax, uploaded;
  void allocate();
  void upload_atoms();
};

}    // namespace LAMMPS_NS
#endif
#endif
