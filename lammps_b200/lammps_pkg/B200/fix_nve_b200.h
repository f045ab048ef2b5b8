/* -*- c++ -*- ----------------------------------------------------------
   fix nve/b200 -- velocity-Verlet integration on the device.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(nve/b200,FixNVEB200);
// clang-format on
#else

#ifndef LMP_FIX_NVE_B200_H
#define LMP_FIX_NVE_B200_H

#include "b200_lmp.h"
#include "fix_nve.h"

namespace LAMMPS_NS {

class FixNVEB200 : public FixNVE, public B200NVEFix {
 public:
  FixNVEB200(class LAMMPS *, int, char **);
  void init() override;
  void initial_integrate(int) override;
  void final_integrate() override;
  void reset_dt() override;
  void b200_params(double &dtv_, double &dtf_, int &groupbit_) override;
};

}    // namespace LAMMPS_NS

#endif
#endif
