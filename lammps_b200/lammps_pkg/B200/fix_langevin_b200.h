/* -*- c++ -*- ----------------------------------------------------------
   fix langevin/b200 -- Langevin thermostat (fix_langevin.cpp:383-507) on
   device-resident atoms.  Argument parsing, the per-type prefactors
   (FixLangevin::init, :268-280) and the target-temperature ramp
   (compute_target, :513-548) stay the reference's own FixLangevin code,
   inherited unchanged; the per-atom loop of post_force runs as a device
   kernel through the C ABI (b200_langevin) on the forces the pair stage of
   this step stored.  run_style verlet/b200 calls post_force between the
   reverse halo and final_integrate, as Verlet::run does (verlet.cpp:340-350).

   Random numbers: the reference draws three uniforms per atom from ONE
   sequential Marsaglia stream in host atom order.  The device stream is
   counter based instead (key = seed, counter = (atom tag, timestep)), so a
   trajectory does not depend on the atom order or on how many sub-domains
   or GPUs carry the box -- statistically equivalent, not the same sequence.
   `package b200 langevin_rng host` makes the host draw the uniforms from the
   reference's RanMars in tag order and hand them to the same kernel: with
   `atom_modify sort 0 0` (host order = tag order) a run then reproduces the
   CPU reference exactly; this is the verification mode.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(langevin/b200,FixLangevinB200);
// clang-format on
#else

#ifndef LMP_FIX_LANGEVIN_B200_H
#define LMP_FIX_LANGEVIN_B200_H

#include "b200_lmp.h"
#include "fix_langevin.h"

#include <vector>

namespace LAMMPS_NS {

class FixLangevinB200 : public FixLangevin, public B200PostForceFix {
 public:
  FixLangevinB200(class LAMMPS *, int, char **);
  void init() override;
  void setup(int) override;
  void post_force(int) override;

 private:
  std::vector<double> g2t, uni;
};

}    // namespace LAMMPS_NS

#endif
#endif
