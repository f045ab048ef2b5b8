/* ----------------------------------------------------------------------
   pair_style eam/b200, eam/alloy/b200, eam/fs/b200
------------------------------------------------------------------------- */

#include "pair_eam_b200.h"

#include "atom.h"
#include "error.h"
#include "fix_b200.h"
#include "force.h"
#include "update.h"

#include <cstring>

using namespace LAMMPS_NS;

PairEAMB200::PairEAMB200(LAMMPS *lmp) : PairEAM(lmp)
{
  respa_enable = 0;
  suffix_flag |= B200_SUFFIX_BIT;
}

PairEAMAlloyB200::PairEAMAlloyB200(LAMMPS *lmp) : PairEAMB200(lmp)
{
  fileformat = SETFL;
  one_coeff = 1;
}

PairEAMFSB200::PairEAMFSB200(LAMMPS *lmp) : PairEAMB200(lmp)
{
  fileformat = FS;
  one_coeff = 1;
}

void PairEAMB200::init_style()
{
  PairEAM::init_style();    // file2array() + array2spline()
  FixB200::instance(lmp);
  if (he_flag) error->all(FLERR, "Pair style eam/b200 does not support eam/he tables");
}

// see PairLJCutB200::compute: a silent no-op would hand stale forces to host-side callers
void PairEAMB200::compute(int, int)
{
  error->all(FLERR, "Pair style eam/b200 computes forces only inside run_style verlet/b200 "
                    "(minimize, rerun and other host-side callers of Pair::compute are not supported)");
}

int PairEAMB200::b200_upload(b200_ctx *ctx)
{
  // splines are [n][nr+1 | nrho+1][7] contiguous (memory->create 3d); type maps are
  // (ntypes+1) and (ntypes+1)^2 ints; scale is (ntypes+1)^2 doubles
  return b200_pair_eam(ctx, atom->ntypes, nr, nrho, rdr, rdrho, rhomax, cutforcesq, type2frho,
                       &type2rhor[0][0], &type2z2r[0][0], &scale[0][0], nfrho,
                       &frho_spline[0][0][0], nrhor, &rhor_spline[0][0][0], nz2r,
                       &z2r_spline[0][0][0]);
}
