/* -*- c++ -*- ----------------------------------------------------------
   pair_style lj/cut/b200 -- lj/cut evaluated by the B200 device engine.
   Coefficients, mixing, restart and data-file handling are inherited from
   PairLJCut; only the force evaluation differs.
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(lj/cut/b200,PairLJCutB200);
// clang-format on
#else

#ifndef LMP_PAIR_LJ_CUT_B200_H
#define LMP_PAIR_LJ_CUT_B200_H

#include "b200_lmp.h"
#include "pair_lj_cut.h"

namespace LAMMPS_NS {

class PairLJCutB200 : public PairLJCut, public B200PairStyle {
 public:
  PairLJCutB200(class LAMMPS *);
  void compute(int, int) override;
  void init_style() override;
  int b200_upload(b200_ctx *ctx) override;
  void b200_ev_setup(int eflag, int vflag) override { ev_setup(eflag, vflag); }
};

}    // namespace LAMMPS_NS

#endif
#endif
