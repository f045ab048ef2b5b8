/* ----------------------------------------------------------------------
   pair_style lj/cut/b200
------------------------------------------------------------------------- */

#include "pair_lj_cut_b200.h"

#include "atom.h"
#include "error.h"
#include "fix_b200.h"
#include "force.h"
#include "update.h"

#include <cstring>

using namespace LAMMPS_NS;

PairLJCutB200::PairLJCutB200(LAMMPS *lmp) : PairLJCut(lmp)
{
  respa_enable = 0;
  suffix_flag |= B200_SUFFIX_BIT;
}

void PairLJCutB200::init_style()
{
  PairLJCut::init_style();
  FixB200::instance(lmp);
  if (atom->molecular != Atom::ATOMIC)
    error->all(FLERR, "Pair style lj/cut/b200 requires an atomic system (no special bonds)");
}

// Forces live on the device and are computed inside run_style verlet/b200 (b200_step), which
// never calls Pair::compute.  Whoever does call it -- minimize, rerun, compute group/group, a
// foreign run_style -- would get stale forces and energies from a silent no-op, and a host-side
// evaluation would be the CPU fallback this package does not have: refuse.
void PairLJCutB200::compute(int, int)
{
  error->all(FLERR, "Pair style lj/cut/b200 computes forces only inside run_style verlet/b200 "
                    "(minimize, rerun and other host-side callers of Pair::compute are not supported)");
}

int PairLJCutB200::b200_upload(b200_ctx *ctx)
{
  // tables are (ntypes+1)^2 contiguous doubles (memory->create), row/column 0 unused
  return b200_pair_lj_cut(ctx, atom->ntypes, &cutsq[0][0], &lj1[0][0], &lj2[0][0], &lj3[0][0],
                          &lj4[0][0], &offset[0][0], force->special_lj);
}
