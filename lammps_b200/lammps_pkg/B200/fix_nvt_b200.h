/* -*- c++ -*- ----------------------------------------------------------
   fix nvt/b200 -- Nose-Hoover chain thermostat (fix nvt = FixNH without a
   barostat, fix_nh.cpp:916-1014) on device-resident atoms.  The chain is
   scalar arithmetic on the current temperature: it stays the reference's
   own FixNH code, inherited unchanged.  What FixNH does per atom -- nve_v,
   nve_x, nh_v_temp (fix_nh.cpp:2278-2352; virtual for exactly this kind of
   override) -- runs as device kernels through the C ABI (b200_nve_v,
   b200_nve_x, b200_scale_v), and the temperature the chain reads is
   compute temp/b200's device sum.  run_style verlet/b200 calls
   initial_integrate / final_integrate of this fix around the device stages
   of the timestep.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(nvt/b200,FixNVTB200);
// clang-format on
#else

#ifndef LMP_FIX_NVT_B200_H
#define LMP_FIX_NVT_B200_H

#include "b200_lmp.h"
#include "fix_nvt.h"

namespace LAMMPS_NS {

class FixNVTB200 : public FixNVT, public B200StagedFix {
 public:
  FixNVTB200(class LAMMPS *, int, char **);
  void init() override;
  void b200_params(double &dtv, double &dtf, int &groupbit) override;

 protected:
  void nve_v() override;
  void nve_x() override;
  void nh_v_temp() override;
};

}    // namespace LAMMPS_NS

#endif
#endif
