/* ----------------------------------------------------------------------
   fix nvt/b200: see fix_nvt_b200.h
------------------------------------------------------------------------- */

#include "fix_nvt_b200.h"

#include "atom.h"
#include "compute.h"
#include "error.h"
#include "fix_b200.h"
#include "update.h"

#include <cstring>

using namespace LAMMPS_NS;

FixNVTB200::FixNVTB200(LAMMPS *lmp, int narg, char **arg) : FixNVT(lmp, narg, arg) {}

void FixNVTB200::init()
{
  FixNVT::init();
  FixB200 *pkg = FixB200::instance(lmp);
  if (atom->rmass_flag) error->all(FLERR, "Fix nvt/b200 requires per-type masses");
  if (strcmp(update->integrate_style, "verlet/b200") != 0)
    error->all(FLERR, "Fix nvt/b200 requires run_style verlet/b200");
  if (which != 0)    // NOBIAS (fix_nh.cpp enum): a bias would need per-atom host work in nh_v_temp
    error->all(FLERR, "Fix nvt/b200 does not support temperature computes with a bias");
  // the chain reads the temperature every step: it must come from the device sum
  if (strcmp(temperature->style, "temp/b200") != 0)
    error->all(FLERR, "Fix nvt/b200 requires a compute temp/b200 (got {})", temperature->style);
}

void FixNVTB200::b200_params(double &dtv_, double &dtf_, int &groupbit_)
{
  dtv_ = dtv;
  dtf_ = dtf;
  groupbit_ = groupbit;
}

/* FixNH::nve_v, fix_nh.cpp:2300-2336 (per-type masses) */
void FixNVTB200::nve_v()
{
  FixB200::instance(lmp)->dev_nve_v(dtf, groupbit);
}

/* FixNH::nve_x, fix_nh.cpp:2278-2298 */
void FixNVTB200::nve_x()
{
  FixB200::instance(lmp)->dev_nve_x(dtv, groupbit);
}

/* FixNH::nh_v_temp, fix_nh.cpp:2338-2352 (no bias) */
void FixNVTB200::nh_v_temp()
{
  // (FixNH::setup only computes t_current on the host; the first scaling happens in
  // initial_integrate of the first step, when the atoms are on the device)
  FixB200::instance(lmp)->dev_scale_v(factor_eta, groupbit);
}
