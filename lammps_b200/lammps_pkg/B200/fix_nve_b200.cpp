/* ----------------------------------------------------------------------
   fix nve/b200: the two half-kicks and the drift run as device kernels
   inside run_style verlet/b200 (b200_step); this class carries dtv, dtf
   and the group to it and refuses to run anywhere else.
------------------------------------------------------------------------- */

#include "fix_nve_b200.h"

#include "atom.h"
#include "error.h"
#include "fix_b200.h"
#include "update.h"

#include <cstring>

using namespace LAMMPS_NS;

FixNVEB200::FixNVEB200(LAMMPS *lmp, int narg, char **arg) : FixNVE(lmp, narg, arg) {}

void FixNVEB200::init()
{
  FixNVE::init();    // dtv = dt, dtf = 0.5 * dt * ftm2v
  FixB200::instance(lmp);
  if (atom->rmass_flag) error->all(FLERR, "Fix nve/b200 requires per-type masses");
  if (strcmp(update->integrate_style, "verlet/b200") != 0)
    error->all(FLERR, "Fix nve/b200 requires run_style verlet/b200");
}

void FixNVEB200::reset_dt()
{
  FixNVE::reset_dt();
  b200_fix_nve(FixB200::instance(lmp)->context(), dtv, dtf, groupbit);
}

// time integration happens inside b200_step(); Modify never gets to call these because
// verlet/b200 drives the step, but a foreign integrator must not silently skip the update
void FixNVEB200::initial_integrate(int)
{
  error->all(FLERR, "Fix nve/b200 can only be driven by run_style verlet/b200");
}

void FixNVEB200::final_integrate()
{
  error->all(FLERR, "Fix nve/b200 can only be driven by run_style verlet/b200");
}

void FixNVEB200::b200_params(double &dtv_, double &dtf_, int &groupbit_)
{
  dtv_ = dtv;
  dtf_ = dtf;
  groupbit_ = groupbit;
}
