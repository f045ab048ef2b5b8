/* ----------------------------------------------------------------------
   fix B200: "package b200 [ngpu] keyword value ..."
     gpus N              N GPUs driven by this process, one brick sub-domain each
                         (devices D, D+1, ...; default 1)
     subdomains M        M >= N sub-domains, dealt round-robin onto the N GPUs
                         (several per GPU: a test/debug configuration)
     device D            first CUDA device of this process (default: $LOCAL_RANK or 0)
     prec double|mixed   arithmetic of the pair kernels (default double)
     profile yes|no      step stage by stage with a device sync after each, so that
                         the Timer breakdown (Pair/Neigh/Comm/Modify) is filled, and
                         print per-phase device times after each run (default no)
     list tile|flat|auto neighbour list layout      tile tx ty tz   bins per tile
     overlap yes|no      interior tiles beside the halo          graph yes|no  CUDA graph
     tpa 1|2|4|8         lanes per atom of the flat kernels      mixed_fx yes|no
     fuse yes|no         fix nve inside the lj/cut pair kernel   eam2 yes|no|auto  eam tile kernels
     build2 yes|no       warp-per-bin list build (cross-check)
     lazy yes|no         queue the per-atom loops of host-driven integrators (nvt, npt, langevin)
                         and run them in one pass (default yes)
     langevin_rng device|host   fix langevin/b200: counter-based device stream (default) or
                         the reference's RanMars drawn on the host in tag order (verification)
   The device contexts are created here and destroyed with the fix, like the
   GPU package ties its devices to fix GPU (precedent: src/GPU/fix_gpu.cpp).
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
    Fix(lmp, narg, arg), host_stale(0), ctx(nullptr), grp(nullptr), nsub(1), device(0),
    prec(B200_PREC_DOUBLE), profile_flag(0), lang_rng_host(0)
{
  if (const char *lr = getenv("LOCAL_RANK")) device = atoi(lr);

  int ngpu = 1, nsubdom = 0;
  std::vector<std::pair<std::string, std::string>> options;
  int iarg = 3;
  // optional leading GPU count, like "package gpu N"
  if (iarg < narg && utils::is_integer(arg[iarg])) {
    ngpu = utils::inumeric(FLERR, arg[iarg], false, lmp);
    iarg++;
  }
  while (iarg < narg) {
    if (iarg + 2 > narg) error->all(FLERR, "Illegal package b200 command: missing value");
    const std::string key = arg[iarg];
    if (key == "device") {
      device = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
    } else if (key == "gpus") {
      ngpu = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
    } else if (key == "subdomains") {
      nsubdom = utils::inumeric(FLERR, arg[iarg + 1], false, lmp);
    } else if (key == "prec") {
      if (strcmp(arg[iarg + 1], "double") == 0) prec = B200_PREC_DOUBLE;
      else if (strcmp(arg[iarg + 1], "mixed") == 0) prec = B200_PREC_MIXED;
      else error->all(FLERR, "Illegal package b200 prec value: {}", arg[iarg + 1]);
    } else if (key == "profile") {
      profile_flag = utils::logical(FLERR, arg[iarg + 1], false, lmp);
    } else if (key == "langevin_rng") {
      if (strcmp(arg[iarg + 1], "host") == 0) lang_rng_host = 1;
      else if (strcmp(arg[iarg + 1], "device") == 0) lang_rng_host = 0;
      else error->all(FLERR, "Illegal package b200 langevin_rng value: {}", arg[iarg + 1]);
    } else if (key == "tile") {
      if (iarg + 4 > narg) error->all(FLERR, "Illegal package b200 command: tile needs three values");
      options.emplace_back("tile", std::string(arg[iarg + 1]) + "," + arg[iarg + 2] + "," + arg[iarg + 3]);
      iarg += 2;
    } else if (key == "list" || key == "overlap" || key == "graph" || key == "tpa" || key == "mixed_fx" ||
               key == "lazy" || key == "fuse" || key == "eam2" || key == "build2") {
      options.emplace_back(key, arg[iarg + 1]);
    } else
      error->all(FLERR, "Unknown package b200 keyword: {}", arg[iarg]);
    iarg += 2;
  }
  if (ngpu < 1) error->all(FLERR, "Illegal package b200 command: gpus must be >= 1");
  if (nsubdom == 0) nsubdom = ngpu;
  if (nsubdom < ngpu) error->all(FLERR, "Illegal package b200 command: fewer sub-domains than GPUs");

  const int have = b200_device_count();
  if (have <= 0)
    error->all(FLERR, "package b200: no CUDA device visible (the B200 package has no CPU fallback)");
  if (device + ngpu > have)
    error->all(FLERR, "package b200: devices {}..{} requested but {} visible", device, device + ngpu - 1, have);

  if (nsubdom > 1) {
    // one process, several sub-domains: LAMMPS sees one rank owning the whole box
    if (comm->nprocs > 1)
      error->all(FLERR, "package b200 gpus/subdomains > 1 is the single-process mode; with MPI run one "
                        "rank per GPU instead");
    nsub = nsubdom;
    std::vector<int> devs(nsub);
    for (int i = 0; i < nsub; i++) devs[i] = device + i % ngpu;
    int rc = b200_group_create(&grp, nsub, devs.data(), prec);
    if (rc != B200_OK) {
      std::string msg = grp ? b200_group_last_error(grp) : "cannot create the device group";
      if (grp) b200_group_destroy(grp);
      grp = nullptr;
      error->one(FLERR, "package b200: {}", msg);
    }
  } else {
    int rc = b200_create(&ctx, device, prec);
    if (rc != B200_OK) {
      std::string msg = ctx ? b200_last_error(ctx) : "cannot create device context";
      if (ctx) b200_destroy(ctx);
      ctx = nullptr;
      error->one(FLERR, "package b200: {}", msg);
    }
  }
  for (int i = 0; i < nctx(); i++)
    for (auto &kv : options) check(b200_set_option(context(i), kv.first.c_str(), kv.second.c_str()), FLERR);
  if (comm->me == 0)
    utils::logmesg(lmp, "B200 package: device {} precision {}{}\n", device,
                   prec == B200_PREC_DOUBLE ? "double" : "mixed",
                   nsub > 1 ? fmt::format(", {} sub-domains on {} GPU(s) in this process", nsub, ngpu) : "");
}

FixB200::~FixB200()
{
  if (grp) b200_group_destroy(grp);
  if (ctx) b200_destroy(ctx);
}

int FixB200::setmask()
{
  return 0;
}

void FixB200::init()
{
  // newton pair on or off: off runs on the tile rows that hold every ghost partner (b200_set_newton)
}

double FixB200::memory_usage()
{
  return 0.0;
}

void FixB200::check(int rc, const char *file, int line)
{
  if (rc == B200_OK) return;
  std::string msg = "no context";
  if (grp) {
    msg = b200_group_last_error(grp);
    if (msg.empty())
      for (int i = 0; i < nsub && msg.empty(); i++) msg = b200_last_error(b200_group_context(grp, i));
  } else if (ctx)
    msg = b200_last_error(ctx);
  error->one(file, line, "B200 package: {} (code {})", msg, rc);
}

/* ---------------------------------------------------------------------- */

void FixB200::dev_setup(int eflag, int vflag)
{
  check(grp ? b200_group_setup(grp, eflag, vflag) : b200_setup(ctx, eflag, vflag), FLERR);
}

void FixB200::dev_step(int eflag, int vflag, int more, int *rebuilt)
{
  check(grp ? b200_group_step_ahead(grp, eflag, vflag, more, rebuilt)
            : b200_step_ahead(ctx, eflag, vflag, more, rebuilt),
        FLERR);
}

void FixB200::dev_tallies(double *eng_vdwl, double *virial)
{
  check(grp ? b200_group_get_tallies(grp, eng_vdwl, virial) : b200_get_tallies(ctx, eng_vdwl, virial), FLERR);
}

void FixB200::dev_peratom(double *eatom, double *vatom)
{
  check(grp ? b200_group_pair_peratom(grp, eatom, vatom) : b200_pair_peratom(ctx, eatom, vatom), FLERR);
}

void FixB200::dev_ke(int groupbit, double *mv2, double *tensor)
{
  check(grp ? b200_group_ke_group(grp, groupbit, mv2, tensor) : b200_ke_group(ctx, groupbit, mv2, tensor),
        FLERR);
}

void FixB200::dev_decide(int *rebuild)
{
  check(grp ? b200_group_decide(grp, rebuild) : b200_decide(ctx, rebuild), FLERR);
}
void FixB200::dev_reneighbor() { check(grp ? b200_group_reneighbor(grp) : b200_reneighbor(ctx), FLERR); }
void FixB200::dev_forward_comm() { check(grp ? b200_group_forward_comm(grp) : b200_forward_comm(ctx), FLERR); }
void FixB200::dev_force_clear() { check(grp ? b200_group_force_clear(grp) : b200_force_clear(ctx), FLERR); }
void FixB200::dev_pair_compute(int eflag, int vflag)
{
  check(grp ? b200_group_pair_compute(grp, eflag, vflag) : b200_pair_compute(ctx, eflag, vflag), FLERR);
}
void FixB200::dev_reverse_comm() { check(grp ? b200_group_reverse_comm(grp) : b200_reverse_comm(ctx), FLERR); }
void FixB200::dev_nve_v(double dtf, int groupbit)
{
  check(grp ? b200_group_nve_v(grp, dtf, groupbit) : b200_nve_v(ctx, dtf, groupbit), FLERR);
}
void FixB200::dev_nve_x(double dtv, int groupbit)
{
  check(grp ? b200_group_nve_x(grp, dtv, groupbit) : b200_nve_x(ctx, dtv, groupbit), FLERR);
}
void FixB200::dev_scale_v(double factor, int groupbit)
{
  check(grp ? b200_group_scale_v(grp, factor, groupbit) : b200_scale_v(ctx, factor, groupbit), FLERR);
}

void FixB200::dev_scale_v3(const double *factor, int groupbit)
{
  check(grp ? b200_group_scale_v3(grp, factor, groupbit) : b200_scale_v3(ctx, factor, groupbit), FLERR);
}
void FixB200::dev_remap(const double *oldlo, const double *oldhi, const double *newlo, const double *newhi,
                        int groupbit)
{
  check(grp ? b200_group_remap(grp, oldlo, oldhi, newlo, newhi, groupbit)
            : b200_remap(ctx, oldlo, oldhi, newlo, newhi, groupbit),
        FLERR);
}

void FixB200::dev_langevin(int ntypes, const double *gfactor1, const double *gfactor2_tsqrt, int groupbit,
                           uint64_t seed, int64_t step, const double *uniforms_by_tag, int64_t nuniform,
                           double *fsum)
{
  check(grp ? b200_group_langevin(grp, ntypes, gfactor1, gfactor2_tsqrt, groupbit, seed, step, uniforms_by_tag,
                                  nuniform, fsum)
            : b200_langevin(ctx, ntypes, gfactor1, gfactor2_tsqrt, groupbit, seed, step, uniforms_by_tag, nuniform,
                            fsum),
        FLERR);
}
void FixB200::dev_add_force(const double *df, int groupbit)
{
  check(grp ? b200_group_add_force(grp, df, groupbit) : b200_add_force(ctx, df, groupbit), FLERR);
}

void FixB200::dev_counts(int *nlocal, int *nghost)
{
  check(grp ? b200_group_count(grp, nlocal, nghost) : b200_get_counts(ctx, nlocal, nghost), FLERR);
}

void FixB200::dev_stats(b200_stats *st)
{
  check(grp ? b200_group_get_stats(grp, st) : b200_get_stats(ctx, st), FLERR);
}

FixB200 *FixB200::instance(LAMMPS *lmp)
{
  Fix *f = lmp->modify->get_fix_by_id("package_b200");
  if (!f) f = lmp->modify->add_fix("package_b200 all B200");
  auto *me = dynamic_cast<FixB200 *>(f);
  if (!me) lmp->error->all(FLERR, "fix package_b200 is not of style B200");
  return me;
}
