/* -*- c++ -*- ----------------------------------------------------------
   fix B200 -- created by "package b200 ..." (or on demand by -sf b200);
   owns the device side of the B200 package: one context (one GPU per
   process) or an in-process group of sub-domains on several GPUs.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(B200,FixB200);
// clang-format on
#else

#ifndef LMP_FIX_B200_H
#define LMP_FIX_B200_H

#include "b200_lmp.h"
#include "fix.h"

#include <string>
#include <utility>
#include <vector>

namespace LAMMPS_NS {

class FixB200 : public Fix {
 public:
  FixB200(class LAMMPS *, int, char **);
  ~FixB200() override;
  int setmask() override;
  void init() override;
  double memory_usage() override;

  // sub-domain contexts: 1 without a group
  int nctx() const { return grp ? nsub : 1; }
  b200_ctx *context(int i = 0) { return grp ? b200_group_context(grp, i) : ctx; }
  b200_group *group() { return grp; }
  int precision() const { return prec; }
  int profile() const { return profile_flag; }
  // 1 while the device copy of the atoms is newer than atom->x/v/f (set by verlet/b200)
  int host_stale;

  // turn a negative b200_* status into error->one() with the library's message
  void check(int rc, const char *file, int line);

  // the calls that communicate between sub-domains, for one context or a group alike
  void dev_setup(int eflag, int vflag);
  void dev_step(int eflag, int vflag, int more, int *rebuilt);    // more: another step follows before any host read
  void dev_tallies(double *eng_vdwl, double *virial);
  void dev_peratom(double *eatom, double *vatom);    // Pair::ev_tally's eatom / vatom, download order
  void dev_ke(int groupbit, double *mv2, double *tensor);
  void dev_counts(int *nlocal, int *nghost);
  // the stages of a timestep one by one (a host fix integrates: fix nvt/b200)
  void dev_decide(int *rebuild);
  void dev_reneighbor();
  void dev_forward_comm();
  void dev_force_clear();
  void dev_pair_compute(int eflag, int vflag);
  void dev_reverse_comm();
  void dev_nve_v(double dtf, int groupbit);
  void dev_nve_x(double dtv, int groupbit);
  void dev_scale_v(double factor, int groupbit);
  void dev_scale_v3(const double *factor, int groupbit);
  void dev_remap(const double *oldlo, const double *oldhi, const double *newlo, const double *newhi, int groupbit);
  void dev_langevin(int ntypes, const double *gfactor1, const double *gfactor2_tsqrt, int groupbit, uint64_t seed,
                    int64_t step, const double *uniforms_by_tag, int64_t nuniform, double *fsum);
  void dev_add_force(const double *df, int groupbit);
  void dev_stats(b200_stats *st);
  // `package b200 langevin_rng host`: fix langevin/b200 draws its uniforms on the host (verification)
  int langevin_rng_host() const { return lang_rng_host; }

  // the package fix of this LAMMPS instance; issues "package b200" defaults if absent
  static FixB200 *instance(class LAMMPS *);

 private:
  b200_ctx *ctx;
  b200_group *grp;
  int nsub, device, prec, profile_flag, lang_rng_host;
};

}    // namespace LAMMPS_NS

#define B200_CHECK(fixptr, call) (fixptr)->check((call), FLERR)

#endif
#endif
