/* ----------------------------------------------------------------------
   fix npt/b200, fix nph/b200: see fix_npt_b200.h
------------------------------------------------------------------------- */

#include "fix_npt_b200.h"

#include "atom.h"
#include "compute.h"
#include "domain.h"
#include "error.h"
#include "fix_b200.h"
#include "update.h"

#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;

template <class Base> void FixNHBaroB200<Base>::init()
{
  Base::init();
  FixB200 *pkg = FixB200::instance(this->lmp);
  if (this->atom->rmass_flag) this->error->all(FLERR, "Fix {} requires per-type masses", this->style);
  if (strcmp(this->update->integrate_style, "verlet/b200") != 0)
    this->error->all(FLERR, "Fix {} requires run_style verlet/b200", this->style);
  if (this->pstyle == 2)    // TRICLINIC (fix_nh.cpp enum{ISO,ANISO,TRICLINIC})
    this->error->all(FLERR, "Fix {} supports orthogonal boxes only (iso, aniso, x, y, z)", this->style);
  if (this->domain->triclinic) this->error->all(FLERR, "Fix {} requires an orthogonal box", this->style);
  if (!this->allremap) this->error->all(FLERR, "Fix {} dilates all atoms (no dilate keyword)", this->style);
  if (!this->rfix.empty()) this->error->all(FLERR, "Fix {} does not support rigid-body fixes", this->style);
  if (this->which != 0) this->error->all(FLERR, "Fix {} does not support temperature computes with a bias", this->style);
  if (this->tstat_flag && strcmp(this->temperature->style, "temp/b200") != 0)
    this->error->all(FLERR, "Fix {} requires a compute temp/b200 (got {})", this->style, this->temperature->style);
  if (pkg->precision() != 0) this->error->all(FLERR, "Fix {} requires package b200 prec double", this->style);
  if (pkg->group()) this->error->all(FLERR, "Fix {} runs on one sub-domain per process", this->style);
}

template <class Base> void FixNHBaroB200<Base>::b200_params(double &dtv_, double &dtf_, int &groupbit_)
{
  dtv_ = this->dtv;
  dtf_ = this->dtf;
  groupbit_ = this->groupbit;
}

template <class Base> void FixNHBaroB200<Base>::nve_v()
{
  FixB200::instance(this->lmp)->dev_nve_v(this->dtf, this->groupbit);
}

template <class Base> void FixNHBaroB200<Base>::nve_x()
{
  FixB200::instance(this->lmp)->dev_nve_x(this->dtv, this->groupbit);
}

template <class Base> void FixNHBaroB200<Base>::nh_v_temp()
{
  FixB200::instance(this->lmp)->dev_scale_v(this->factor_eta, this->groupbit);
}

/* FixNH::nh_v_press, fix_nh.cpp:2227-2252 (orthogonal box, no bias) */
template <class Base> void FixNHBaroB200<Base>::nh_v_press()
{
  double factor[3];
  factor[0] = exp(-this->dt4 * (this->omega_dot[0] + this->mtk_term2));
  factor[1] = exp(-this->dt4 * (this->omega_dot[1] + this->mtk_term2));
  factor[2] = exp(-this->dt4 * (this->omega_dot[2] + this->mtk_term2));
  FixB200::instance(this->lmp)->dev_scale_v3(factor, this->groupbit);
}

/* FixNH::remap, fix_nh.cpp:1156-1300.  The box arithmetic is the base class's; the atoms it
   would convert to lamda coordinates and back live on the device, so the base runs with zero
   host atoms and the device applies the same two conversions between the old and the new box. */
template <class Base> void FixNHBaroB200<Base>::remap()
{
  double oldlo[3], oldhi[3];
  for (int d = 0; d < 3; d++) {
    oldlo[d] = this->domain->boxlo[d];
    oldhi[d] = this->domain->boxhi[d];
  }
  const int nlocal = this->atom->nlocal;
  this->atom->nlocal = 0;
  Base::remap();
  this->atom->nlocal = nlocal;
  // (allremap, checked in init(): every atom is dilated -- group `all` is bit 0; dilate_group_bit
  // is only set for the `dilate` keyword, fix_nh.cpp init())
  FixB200::instance(this->lmp)->dev_remap(oldlo, oldhi, this->domain->boxlo, this->domain->boxhi, 1);
}

namespace LAMMPS_NS {
template class FixNHBaroB200<FixNPT>;
template class FixNHBaroB200<FixNPH>;
}    // namespace LAMMPS_NS
