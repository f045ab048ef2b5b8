/* ----------------------------------------------------------------------
   compute temp/b200: ComputeTemp::compute_scalar / compute_vector
   (compute_temp.cpp:73-140) with the sums over atoms taken on the device
   (b200_ke_group) whenever the device copy of the velocities is the current
   one; otherwise the host loops of the base class run on atom->v.
------------------------------------------------------------------------- */

#include "compute_temp_b200.h"

#include "atom.h"
#include "comm.h"
#include "error.h"
#include "fix_b200.h"
#include "force.h"
#include "update.h"

using namespace LAMMPS_NS;

ComputeTempB200::ComputeTempB200(LAMMPS *lmp, int narg, char **arg) : ComputeTemp(lmp, narg, arg) {}

double ComputeTempB200::compute_scalar()
{
  FixB200 *pkg = FixB200::instance(lmp);
  if (!pkg->host_stale || atom->rmass) return ComputeTemp::compute_scalar();

  invoked_scalar = update->ntimestep;
  double t = 0.0;
  pkg->dev_ke(groupbit, &t, nullptr);
  // one process per GPU under MPI: the device sum is this rank's (tallies local), reduce as usual
  if (comm->nprocs > 1) MPI_Allreduce(&t, &scalar, 1, MPI_DOUBLE, MPI_SUM, world);
  else scalar = t;
  if (dynamic) dof_compute();
  if (dof < 0.0 && natoms_temp > 0.0) error->all(FLERR, "Temperature compute degrees of freedom < 0");
  scalar *= tfactor;
  return scalar;
}

void ComputeTempB200::compute_vector()
{
  FixB200 *pkg = FixB200::instance(lmp);
  if (!pkg->host_stale || atom->rmass) {
    ComputeTemp::compute_vector();
    return;
  }
  invoked_vector = update->ntimestep;
  double mv2, t[6];
  pkg->dev_ke(groupbit, &mv2, t);
  if (comm->nprocs > 1) MPI_Allreduce(t, vector, 6, MPI_DOUBLE, MPI_SUM, world);
  else
    for (int i = 0; i < 6; i++) vector[i] = t[i];
  for (int i = 0; i < 6; i++) vector[i] *= force->mvv2e;
}
