/* ----------------------------------------------------------------------
   fix B200: "package b200 [ngpu] keyword value ..."
     device D            CUDA device of this process (default: $LOCAL_RANK or 0)
     prec double|mixed   arithmetic of the pair kernels (default double)
     profile yes|no      per-phase device timing printed after each run (default no)
   The context is created here and destroyed with the fix, like the GPU
   package ties its device to fix GPU (precedent: src/GPU/fix_gpu.cpp).
------------------------------------------------------------------------- */

#include "fix_b200.h"

#include "comm.h"
#include "error.h"
#include "force.h"
#include "input.h"
#include "modify.h"
#include "utils.h"

#include <cstdlib>
#include <cstring>

using namespace LAMMPS_NS;

FixB200::FixB200(LAMMPS *lmp, int narg, char **arg) :
    Fix(lmp, narg, arg), ctx(nullptr), device(0), prec(B200_PREC_DOUBLE), profile_flag(0)
{
  if (const char *lr = getenv("LOCAL_RANK")) device = atoi(lr);

  int iarg = 3;
  // optional leading GPU count, accepted for symmetry with "package gpu N"; one GPU per process
  if (iarg < narg && utils::is_integer(arg[iarg])) {
    int ngpu = utils::inumeric(FLERR, arg[iarg], false, lmp);
    if (ngpu > 1)
      error->all(FLERR, "package b200: one GPU per process; run one process per GPU instead");
    iarg++;
  }
  while (iarg < narg) {
    if (iarg + 2 > narg) error->all(FLERR, "Illegal package b200 command: missing value");
    if (strcmp(arg[iarg], "device") == 0) {
      device = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
    } else if (strcmp(arg[iarg], "prec") == 0) {
      if (strcmp(arg[iarg + 1], "double") == 0) prec = B200_PREC_DOUBLE;
      else if (strcmp(arg[iarg + 1], "mixed") == 0) prec = B200_PREC_MIXED;
      else error->all(FLERR, "Illegal package b200 prec value: {}", arg[iarg + 1]);
    } else if (strcmp(arg[iarg], "profile") == 0) {
      profile_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp);
    } else
      error->all(FLERR, "Unknown package b200 keyword: {}", arg[iarg]);
    iarg += 2;
  }

  if (b200_device_count() <= 0)
    error->all(FLERR, "package b200: no CUDA device visible (the B200 package has no CPU fallback)");
  int rc = b200_create(&ctx, device, prec);
  if (rc != B200_OK) {
    std::string msg = ctx ? b200_last_error(ctx) : "cannot create device context";
    if (ctx) b200_destroy(ctx);
    ctx = nullptr;
    error->one(FLERR, "package b200: {}", msg);
  }
  if (comm->me == 0)
    utils::logmesg(lmp, "B200 package: device {} precision {}\n", device,
                   prec == B200_PREC_DOUBLE ? "double" : "mixed");
}

FixB200::~FixB200()
{
  if (ctx) b200_destroy(ctx);
}

int FixB200::setmask()
{
  return 0;
}

void FixB200::init()
{
  if (!force->newton_pair)
    error->all(FLERR, "The B200 package requires newton pair on (half neighbor lists)");
}

double FixB200::memory_usage()
{
  return 0.0;
}

void FixB200::check(int rc, const char *file, int line)
{
  if (rc == B200_OK) return;
  error->one(file, line, "B200 package: {} (code {})", ctx ? b200_last_error(ctx) : "no context", rc);
}

FixB200 *FixB200::instance(LAMMPS *lmp)
{
  Fix *f = lmp->modify->get_fix_by_id("package_b200");
  if (!f) f = lmp->modify->add_fix("package_b200 all B200");
  auto *me = dynamic_cast<FixB200 *>(f);
  if (!me) lmp->error->all(FLERR, "fix package_b200 is not of style B200");
  return me;
}
