/* -*- c++ -*- ----------------------------------------------------------
   compute pe/b200 and compute pressure/b200 -- ComputePE (compute_pe.cpp:84)
   and ComputePressure (compute_pressure.cpp:240-302) under the names the
   `-sf b200` suffix looks for (output.cpp:74-76 creates thermo_pe and
   thermo_press with the suffix).  Both read only the pair style's scalars
   (eng_vdwl, virial[6]), which run_style verlet/b200 refreshes from the
   device tallies on every energy/virial step (VerletB200::fetch_tallies:
   7 doubles), and a temperature compute: with temp/b200 a thermo step
   needs no atom on the host.  The classes refuse the options that WOULD
   need per-atom host data the device run does not keep current.
------------------------------------------------------------------------- */

#ifdef COMPUTE_CLASS
// clang-format off
ComputeStyle(pe/b200,ComputePEB200);
ComputeStyle(pressure/b200,ComputePressureB200);
// clang-format on
#else

#ifndef LMP_COMPUTE_PE_B200_H
#define LMP_COMPUTE_PE_B200_H

#include "compute_pe.h"
#include "compute_pressure.h"

namespace LAMMPS_NS {

class ComputePEB200 : public ComputePE {
 public:
  ComputePEB200(class LAMMPS *, int, char **);
};

class ComputePressureB200 : public ComputePressure {
 public:
  ComputePressureB200(class LAMMPS *, int, char **);
  void init() override;
};

}    // namespace LAMMPS_NS

#endif
#endif
