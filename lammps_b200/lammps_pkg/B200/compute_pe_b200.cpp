/* ----------------------------------------------------------------------
   compute pe/b200, compute pressure/b200: see compute_pe_b200.h
------------------------------------------------------------------------- */

#include "compute_pe_b200.h"

#include "error.h"
#include "force.h"

using namespace LAMMPS_NS;

ComputePEB200::ComputePEB200(LAMMPS *lmp, int narg, char **arg) : ComputePE(lmp, narg, arg) {}

ComputePressureB200::ComputePressureB200(LAMMPS *lmp, int narg, char **arg) :
    ComputePressure(lmp, narg, arg)
{
}

void ComputePressureB200::init()
{
  ComputePressure::init();
  // the pair virial is the only force contribution the device run keeps current
  if (force->bond || force->angle || force->dihedral || force->improper || force->kspace)
    error->all(FLERR, "compute pressure/b200 supports pairwise short-range forces only");
}
