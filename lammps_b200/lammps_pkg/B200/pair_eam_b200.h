/* -*- c++ -*- ----------------------------------------------------------
   pair_style eam/b200 (+ eam/alloy/b200, eam/fs/b200) -- EAM evaluated by
   the B200 device engine: density pass, ghost-density reverse halo,
   embedding, F' forward halo, force pass.  File parsing and spline
   construction are inherited from PairEAM.
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(eam/b200,PairEAMB200);
PairStyle(eam/alloy/b200,PairEAMAlloyB200);
PairStyle(eam/fs/b200,PairEAMFSB200);
// clang-format on
#else

#ifndef LMP_PAIR_EAM_B200_H
#define LMP_PAIR_EAM_B200_H

#include "b200_lmp.h"
#include "pair_eam.h"

namespace LAMMPS_NS {

class PairEAMB200 : public PairEAM, public B200PairStyle {
 public:
  PairEAMB200(class LAMMPS *);
  void compute(int, int) override;
  void init_style() override;
  int b200_upload(b200_ctx *ctx) override;
  void b200_ev_setup(int eflag, int vflag) override { ev_setup(eflag, vflag); }
};

// setfl / Finnis-Sinclair files: same device kernels, other file reader (PairEAM::fileformat)
class PairEAMAlloyB200 : public PairEAMB200 {
 public:
  PairEAMAlloyB200(class LAMMPS *);
};

class PairEAMFSB200 : public PairEAMB200 {
 public:
  PairEAMFSB200(class LAMMPS *);
};

}    // namespace LAMMPS_NS

#endif
#endif
