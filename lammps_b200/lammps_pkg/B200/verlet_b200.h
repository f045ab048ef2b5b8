/* -*- c++ -*- ----------------------------------------------------------
   run_style verlet/b200 -- the velocity-Verlet timestep loop with every
   per-step stage resident on the device (selected automatically by
   "-sf b200", like verlet/kk is by "-sf kk").
------------------------------------------------------------------------- */

#ifdef INTEGRATE_CLASS
// clang-format off
IntegrateStyle(verlet/b200,VerletB200);
// clang-format on
#else

#ifndef LMP_VERLET_B200_H
#define LMP_VERLET_B200_H

#include "b200_lmp.h"
#include "verlet.h"

#include <vector>

namespace LAMMPS_NS {

class VerletB200 : public Verlet {
 public:
  VerletB200(class LAMMPS *, int, char **);
  void init() override;
  void setup(int flag) override;
  void setup_minimal(int) override;
  void run(int) override;
  void cleanup() override;
  void reset_dt() override;

 protected:
  class FixB200 *pkg;
  b200_ctx *ctx;    // the context of this process (nullptr when the package drives a group)
  B200PairStyle *bpair;
  B200NVEFix *bnve;
  B200StagedFix *bstaged;    // fix nvt/b200: the host fix drives its per-atom loops itself
  class Fix *staged_fix;
  std::vector<class Fix *> post_fixes;    // B200PostForceFix instances (fix langevin/b200), in fix order
  int resident;    // 1 once atoms have been handed to the device in this run
  int joined;      // 1 once this rank joined the NCCL communicator of the package
  int thermo_on_device;    // every compute a thermo step evaluates reads device sums or scalars

  void upload();                  // host atom arrays + all parameters -> device
  void download(int with_ghosts); // device -> host atom arrays (x, v, f, type, tag, mask, image)
  void fetch_tallies();           // eng_vdwl / virial -> force->pair
  void device_setup(int flag, int output_flag);
  void publish_neighbor_stats();
  void refuse_per_atom_tallies();
  void fill_per_atom_tallies();   // Pair::eatom / vatom <- device on steps that ask for them
  void step_by_stage(int eflag, int vflag);    // `package b200 profile yes`: Timer breakdown
  void step_staged_fix(int eflag, int vflag);  // a B200StagedFix integrates (fix nvt/b200)
  void step_nve_post_force(int eflag, int vflag);    // fix nve/b200 + post-force fixes (fix langevin/b200)
};

}    // namespace LAMMPS_NS

#endif
#endif
