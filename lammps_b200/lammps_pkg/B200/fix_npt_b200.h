/* -*- c++ -*- ----------------------------------------------------------
   fix npt/b200, fix nph/b200 -- Nose-Hoover thermostat + barostat (FixNH
   with pstat_flag, fix_nh.cpp:916-1014) on device-resident atoms, for
   orthogonal boxes (iso / aniso / x y z; no tilt factors).  As for fix
   nvt/b200 the chain, the barostat equations (couple, nh_omega_dot,
   nhc_press_integrate) and the box update stay the reference's own FixNH
   code; its per-atom loops run on the device: nve_v, nve_x, nh_v_temp,
   nh_v_press (b200_scale_v3) and remap (b200_remap: the atoms are dilated
   with the box and the device adopts the new box).  Temperature and pressure
   come from compute temp/b200 and compute pressure/b200 (device sums,
   virial tallied on the device every step because FixNH asks for it).
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(npt/b200,FixNPTB200);
FixStyle(nph/b200,FixNPHB200);
// clang-format on
#else

#ifndef LMP_FIX_NPT_B200_H
#define LMP_FIX_NPT_B200_H

#include "b200_lmp.h"
#include "fix_nph.h"
#include "fix_npt.h"

namespace LAMMPS_NS {

// the device overrides shared by npt and nph (both are FixNH with different constructors)
template <class Base> class FixNHBaroB200 : public Base, public B200StagedFix {
 public:
  FixNHBaroB200(class LAMMPS *lmp, int narg, char **arg) : Base(lmp, narg, arg) {}
  void init() override;
  void b200_params(double &dtv, double &dtf, int &groupbit) override;
  bool b200_box_change() override { return true; }

 protected:
  void nve_v() override;
  void nve_x() override;
  void nh_v_temp() override;
  void nh_v_press() override;
  void remap() override;
};

class FixNPTB200 : public FixNHBaroB200<FixNPT> {
 public:
  FixNPTB200(class LAMMPS *lmp, int narg, char **arg) : FixNHBaroB200<FixNPT>(lmp, narg, arg) {}
};

class FixNPHB200 : public FixNHBaroB200<FixNPH> {
 public:
  FixNPHB200(class LAMMPS *lmp, int narg, char **arg) : FixNHBaroB200<FixNPH>(lmp, narg, arg) {}
};

}    // namespace LAMMPS_NS

#endif
#endif
