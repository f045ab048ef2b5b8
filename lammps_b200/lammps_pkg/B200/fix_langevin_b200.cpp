/* ----------------------------------------------------------------------
   fix langevin/b200: see fix_langevin_b200.h
------------------------------------------------------------------------- */

#include "fix_langevin_b200.h"

#include "atom.h"
#include "comm.h"
#include "compute.h"
#include "error.h"
#include "fix_b200.h"
#include "group.h"
#include "random_mars.h"
#include "update.h"

#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;

enum { NOBIAS, BIAS };               // fix_langevin.cpp:42
enum { CONSTANT, EQUAL, ATOM };      // fix_langevin.cpp:43

FixLangevinB200::FixLangevinB200(LAMMPS *lmp, int narg, char **arg) : FixLangevin(lmp, narg, arg) {}

void FixLangevinB200::init()
{
  FixLangevin::init();
  if (strcmp(update->integrate_style, "verlet/b200") != 0)
    error->all(FLERR, "Fix langevin/b200 requires run_style verlet/b200");
  if (atom->rmass_flag) error->all(FLERR, "Fix langevin/b200 requires per-type masses");
  if (tstyle == ATOM) error->all(FLERR, "Fix langevin/b200 does not support an atom-style temperature variable");
  if (tallyflag) error->all(FLERR, "Fix langevin/b200 does not support tally yes");
  if (tbiasflag == BIAS) error->all(FLERR, "Fix langevin/b200 does not support a temperature bias");
  if (oflag || ascale != 0.0) error->all(FLERR, "Fix langevin/b200 thermostats point particles only");
}

/* FixLangevin::setup, fix_langevin.cpp:295-305 (verlet branch) */
void FixLangevinB200::setup(int vflag)
{
  post_force(vflag);
}

/* FixLangevin::post_force_templated<0,0,0,0,ZERO>, fix_langevin.cpp:383-507 */
void FixLangevinB200::post_force(int /*vflag*/)
{
  FixB200 *pkg = FixB200::instance(lmp);
  compute_target();

  const int ntypes = atom->ntypes;
  g2t.assign(ntypes + 1, 0.0);
  for (int t = 1; t <= ntypes; t++) g2t[t] = gfactor2[t] * tsqrt;

  bigint count = 0;
  if (zeroflag) {
    count = group->count(igroup);
    if (count == 0) error->all(FLERR, "Cannot zero Langevin force of 0 atoms");
  }

  const double *u = nullptr;
  bigint nu = 0;
  if (pkg->langevin_rng_host()) {
    // the reference's stream, consumed in tag order: three draws per atom of the group, atom
    // after atom (what its loop does when the host order is the tag order and the group is all)
    if (comm->nprocs > 1 || igroup != 0 || atom->natoms > MAXSMALLINT / 3 || !atom->tag_consecutive())
      error->all(FLERR, "package b200 langevin_rng host needs one process, group all, consecutive atom IDs "
                        "and < 7e8 atoms");
    nu = 3 * atom->natoms;
    uni.resize(nu);
    for (bigint k = 0; k < nu; k++) uni[k] = random->uniform();
    u = uni.data();
  }

  double fsum[3] = {0.0, 0.0, 0.0};
  pkg->dev_langevin(ntypes, gfactor1, g2t.data(), groupbit, (uint64_t) seed, (int64_t) update->ntimestep, u,
                    (int64_t) nu, zeroflag ? fsum : nullptr);
  if (zeroflag) {
    double fsumall[3];
    MPI_Allreduce(fsum, fsumall, 3, MPI_DOUBLE, MPI_SUM, world);
    double df[3];
    for (int d = 0; d < 3; d++) df[d] = -(fsumall[d] / count);
    pkg->dev_add_force(df, groupbit);
  }
  pkg->host_stale = 1;    // the forces on the device are newer than atom->f
}
