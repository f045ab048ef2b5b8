/* -*- c++ -*- ----------------------------------------------------------
   compute temp/b200 -- ComputeTemp whose sum over atoms runs on the device
   while the atoms are resident there (run_style verlet/b200): a thermo
   step then moves 56 bytes instead of every atom's velocity.  Selected by
   the suffix for the thermo_temp compute (output.cpp:74-76).
------------------------------------------------------------------------- */

#ifdef COMPUTE_CLASS
// clang-format off
ComputeStyle(temp/b200,ComputeTempB200);
// clang-format on
#else

#ifndef LMP_COMPUTE_TEMP_B200_H
#define LMP_COMPUTE_TEMP_B200_H

#include "compute_temp.h"

namespace LAMMPS_NS {

class ComputeTempB200 : public ComputeTemp {
 public:
  ComputeTempB200(class LAMMPS *, int, char **);
  double compute_scalar() override;
  void compute_vector() override;
};

}    // namespace LAMMPS_NS

#endif
#endif
