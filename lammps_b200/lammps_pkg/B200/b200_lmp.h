/* -*- c++ -*- ----------------------------------------------------------
   B200 package for LAMMPS: host-side glue between LAMMPS styles and the
   C ABI of libb200md (include/b200_md.h).  New code, not derived from the
   GPU or KOKKOS packages.
------------------------------------------------------------------------- */

#ifndef LMP_B200_LMP_H
#define LMP_B200_LMP_H

#include "b200_md.h"

namespace LAMMPS_NS {

class LAMMPS;

// bit for Pair::suffix_flag (next free bit after Suffix::KOKKOS, src/suffix.h)
enum { B200_SUFFIX_BIT = 1 << 5 };

// implemented by every pair style of the package: hand the finished coefficient tables
// (products of init_one() / array2spline()) to the device context
class B200PairStyle {
 public:
  virtual ~B200PairStyle() noexcept(false) {}
  virtual int b200_upload(b200_ctx *ctx) = 0;
  // Pair::ev_setup (protected): allocate and zero eatom / vatom for the atoms now on the host
  virtual void b200_ev_setup(int eflag, int vflag) = 0;
};

// marker for the time-integration fix verlet/b200 knows how to run on the device
class B200NVEFix {
 public:
  virtual ~B200NVEFix() noexcept(false) {}
  virtual void b200_params(double &dtv, double &dtf, int &groupbit) = 0;
};

// a time-integration fix whose own initial_integrate / final_integrate run on the host and call
// device kernels for the per-atom loops (fix nvt/b200): verlet/b200 then drives the timestep stage
// by stage around them
class B200StagedFix {
 public:
  virtual ~B200StagedFix() noexcept(false) {}
  virtual void b200_params(double &dtv, double &dtf, int &groupbit) = 0;
  virtual bool b200_box_change() { return false; }    // the fix moves the box itself (barostat)
};

// a fix that changes the forces after the pair stage (fix langevin/b200): its own post_force()
// launches device kernels on the stored forces; verlet/b200 then steps stage by stage and calls
// it where Verlet::run calls Modify::post_force (verlet.cpp:340-350)
class B200PostForceFix {
 public:
  virtual ~B200PostForceFix() noexcept(false) {}
};

}    // namespace LAMMPS_NS

#endif
