/* ----------------------------------------------------------------------
   run_style verlet/b200

   Same sequence of stages as Verlet::setup()/run() (src/verlet.cpp), but
   every stage of a timestep -- fix nve half-kicks and drift, the rebuild
   decision, ghost halo or pbc/exchange/borders/list build, pair forces
   with energy/virial tallies -- executes on the device through the C ABI
   of libb200md.  Host arrays (atom->x/v/f) are refreshed only when host
   code needs them: when a dump, a restart or a compute without a /b200
   version is due, and at the end of a run; a thermo step whose computes
   are temp/b200, pe and pressure moves a few scalars.
------------------------------------------------------------------------- */

#include "verlet_b200.h"

#include "atom.h"
#include "atom_vec.h"
#include "comm.h"
#include "compute.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "fix_b200.h"
#include "force.h"
#include "memory.h"
#include "modify.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "timer.h"
#include "update.h"

#include <cstring>
#include <vector>

using namespace LAMMPS_NS;
using namespace FixConst;

VerletB200::VerletB200(LAMMPS *lmp, int narg, char **arg) :
    Verlet(lmp, narg, arg), pkg(nullptr), ctx(nullptr), bpair(nullptr), bnve(nullptr), bstaged(nullptr),
    staged_fix(nullptr), resident(0),
    joined(0), thermo_on_device(0)
{
}

/* ---------------------------------------------------------------------- */

void VerletB200::init()
{
  Verlet::init();

  pkg = FixB200::instance(lmp);
  ctx = pkg->group() ? nullptr : pkg->context();

  if (domain->dimension != 3) error->all(FLERR, "run_style verlet/b200 requires a 3d system");
  if (atom->molecular != Atom::ATOMIC || atom->rmass_flag)
    error->all(FLERR, "run_style verlet/b200 requires atom_style atomic with per-type masses");
  if (force->kspace || force->bond || force->angle || force->dihedral || force->improper)
    error->all(FLERR, "run_style verlet/b200 supports pairwise short-range forces only");

  bpair = dynamic_cast<B200PairStyle *>(force->pair);
  if (!bpair)
    error->all(FLERR, "run_style verlet/b200 requires a /b200 pair style (lj/cut/b200, eam/b200)");

  // the only fixes that may act during a timestep are nve/b200 instances (one device group)
  const int stepmask = INITIAL_INTEGRATE | POST_INTEGRATE | PRE_EXCHANGE | PRE_NEIGHBOR |
      POST_NEIGHBOR | PRE_FORCE | PRE_REVERSE | POST_FORCE | FINAL_INTEGRATE | END_OF_STEP;
  bnve = nullptr;
  bstaged = nullptr;
  staged_fix = nullptr;
  post_fixes.clear();
  for (auto &fix : modify->get_fix_list()) {
    if (dynamic_cast<B200PostForceFix *>(fix)) {
      post_fixes.push_back(fix);
      continue;
    }
    auto *nve = dynamic_cast<B200NVEFix *>(fix);
    auto *stg = dynamic_cast<B200StagedFix *>(fix);
    if (nve || stg) {
      if (bnve || bstaged)
        error->all(FLERR, "run_style verlet/b200 supports a single time-integration fix (nve/b200 or nvt/b200)");
      bnve = nve;
      bstaged = stg;
      if (stg) staged_fix = fix;
      continue;
    }
    if (modify->get_fix_mask(fix) & stepmask)
      error->all(FLERR, "Fix {} (style {}) acts during the timestep and has no /b200 version", fix->id,
                 fix->style);
  }
  if (!bnve && !bstaged)
    error->all(FLERR, "run_style verlet/b200 requires fix nve/b200, nvt/b200, npt/b200 or nph/b200");
  // the only thing that may change the box is a barostat whose device version moves the atoms too
  if (domain->box_change && !(bstaged && bstaged->b200_box_change()))
    error->all(FLERR, "run_style verlet/b200 requires a fixed box (or fix npt/b200, nph/b200)");
  resident = 0;
  pkg->host_stale = 0;

  // a thermo step needs no atoms on the host if every compute is one that reads device sums
  // (temp/b200) or the pair style's scalars (pe, pressure on a temp/b200)
  thermo_on_device = 1;
  for (auto &c : modify->get_compute_list()) {
    const std::string style = c->style;
    if (style == "temp/b200" || style == "pe" || style == "pressure" || style == "pe/b200" ||
        style == "pressure/b200")
      continue;
    thermo_on_device = 0;
  }
  for (auto &c : modify->get_compute_list())
    if (strcmp(c->style, "temp") == 0) thermo_on_device = 0;

  // one MPI rank = one GPU = one brick sub-domain: join the NCCL communicator once.  The
  // ncclUniqueId travels over MPI; afterwards all halo traffic is device-to-device.  The host
  // reduces energy, virial and kinetic energy over the ranks itself (compute pe / pressure /
  // temp call MPI_Allreduce), so the device tallies stay per rank.
  if (comm->nprocs > 1 && !joined) {
    char id[128];
    memset(id, 0, sizeof id);
    if (comm->me == 0) B200_CHECK(pkg, b200_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_CHAR, 0, world);
    B200_CHECK(pkg, b200_comm_init(ctx, comm->nprocs, comm->me, id));
    B200_CHECK(pkg, b200_set_option(ctx, "tallies", "local"));
    joined = 1;
  }
}

/* ----------------------------------------------------------------------
   host state -> device: box, decomposition, neighbor settings, atoms,
   pair coefficients, integrator constants
------------------------------------------------------------------------- */

void VerletB200::upload()
{
  double dtv, dtf;
  int groupbit;
  if (bnve) bnve->b200_params(dtv, dtf, groupbit);
  else bstaged->b200_params(dtv, dtf, groupbit);
  for (int i = 0; i < pkg->nctx(); i++) {
    b200_ctx *c = pkg->context(i);
    if (domain->triclinic)
      B200_CHECK(pkg,
                 b200_set_box_triclinic(c, domain->boxlo, domain->boxhi, domain->xy, domain->xz, domain->yz,
                                        domain->periodicity, force->angstrom));
    else
      B200_CHECK(pkg, b200_set_box(c, domain->boxlo, domain->boxhi, domain->periodicity));
    B200_CHECK(pkg, b200_set_newton(c, force->newton_pair));
    B200_CHECK(pkg, b200_neigh_modify_groups(c, neighbor->nex_group, neighbor->ex1_bit, neighbor->ex2_bit));
    B200_CHECK(pkg,
               b200_set_neighbor(c, neighbor->skin, neighbor->every, neighbor->delay,
                                 neighbor->dist_check, neighbor->oneatom));
    // neigh_modify once / exclude type: Neighbor::init has built the symmetric ex_type table
    B200_CHECK(pkg,
               b200_neigh_modify(c, neighbor->build_once, atom->ntypes,
                                 (neighbor->nex_type && neighbor->ex_type) ? &neighbor->ex_type[0][0] : nullptr));
  }
  const int nlocal = atom->nlocal;
  if (pkg->group()) {
    // single process, several sub-domains: this rank owns the whole box and hands all of it over
    int grid[3];
    B200_CHECK(pkg, b200_group_auto_grid(pkg->nctx(), domain->prd, grid));
    B200_CHECK(pkg, b200_group_set_grid(pkg->group(), grid));
    B200_CHECK(pkg,
               b200_group_set_atoms(pkg->group(), nlocal, atom->ntypes, atom->mass,
                                    nlocal ? &atom->x[0][0] : nullptr, nlocal ? &atom->v[0][0] : nullptr,
                                    atom->type, atom->tag, atom->mask, atom->image));
  } else {
    B200_CHECK(pkg, b200_set_decomposition(ctx, comm->procgrid, comm->myloc));
    if (comm->nprocs > 1)
      B200_CHECK(pkg,
                 b200_set_rank_grid(ctx, &comm->grid2proc[0][0][0],
                                    comm->procgrid[0] * comm->procgrid[1] * comm->procgrid[2]));
    B200_CHECK(pkg,
               b200_set_atoms(ctx, nlocal, atom->ntypes, atom->mass, nlocal ? &atom->x[0][0] : nullptr,
                              nlocal ? &atom->v[0][0] : nullptr, atom->type, atom->tag, atom->mask,
                              atom->image));
  }
  for (int i = 0; i < pkg->nctx(); i++) {
    b200_ctx *c = pkg->context(i);
    B200_CHECK(pkg, bpair->b200_upload(c));
    B200_CHECK(pkg, b200_fix_nve(c, dtv, dtf, groupbit));
  }
}

/* ----------------------------------------------------------------------
   device -> host arrays; atoms come back in device (bin-sorted) order,
   which is as legitimate a local order as any Atom::sort() leaves
------------------------------------------------------------------------- */

void VerletB200::download(int with_ghosts)
{
  int nlocal, nghost;
  pkg->dev_counts(&nlocal, &nghost);
  if (pkg->group()) with_ghosts = 0;    // ghosts live between the sub-domains, not on this host
  const int nall = nlocal + (with_ghosts ? nghost : 0);
  while (nall > atom->nmax) atom->avec->grow(0);
  atom->nlocal = nlocal;
  atom->nghost = with_ghosts ? nghost : 0;
  pkg->host_stale = 0;
  if (nall == 0) return;
  if (pkg->group())
    B200_CHECK(pkg,
               b200_group_get_atoms(pkg->group(), &atom->x[0][0], &atom->v[0][0], &atom->f[0][0],
                                    atom->type, atom->tag, atom->mask, atom->image));
  else
    B200_CHECK(pkg,
               b200_get_atoms(ctx, with_ghosts, &atom->x[0][0], &atom->v[0][0], &atom->f[0][0], atom->type,
                              atom->tag, atom->mask, atom->image));
}

void VerletB200::fetch_tallies()
{
  Pair *pair = force->pair;
  pkg->dev_tallies(&pair->eng_vdwl, pair->virial);
  pair->eng_coul = 0.0;
}

// compute centroid/stress/atom asks for Pair::cvatom (pair.cpp:1200-1350), which the device does
// not compute: handing back zeros would be silently wrong
void VerletB200::refuse_per_atom_tallies()
{
  if (vflag & VIRIAL_CENTROID)
    error->all(FLERR, "run_style verlet/b200 does not provide the per-atom centroid virial "
                      "(compute centroid/stress/atom needs the CPU pair styles)");
}

/* ----------------------------------------------------------------------
   compute pe/atom, stress/atom and friends read Pair::eatom / Pair::vatom
   (Pair::ev_tally, pair.cpp:1087-1182).  On a step that asks for them the
   device computes both for its owned atoms (b200_pair_peratom: each atom's
   half of every pair term, ghost shares already returned to their owners)
   and they are stored in the pair style's own arrays, in the order the atoms
   were just downloaded in.  The computes then run unchanged; their reverse
   communication (ComputePEAtom::compute_peratom) needs the host's swap
   lists to describe the current atoms, hence the borders() call: the host
   ghosts it creates carry zeros.
------------------------------------------------------------------------- */

void VerletB200::fill_per_atom_tallies()
{
  if (!(eflag & ENERGY_ATOM) && !(vflag & VIRIAL_ATOM)) return;
  Pair *pair = force->pair;
  if (pkg->host_stale) download(0);
  if (domain->triclinic) domain->x2lamda(atom->nlocal);
  comm->borders();
  if (domain->triclinic) domain->lamda2x(atom->nlocal + atom->nghost);
  bpair->b200_ev_setup(eflag, vflag);
  pkg->dev_peratom((eflag & ENERGY_ATOM) ? pair->eatom : nullptr,
                   (vflag & VIRIAL_ATOM) ? &pair->vatom[0][0] : nullptr);
}

/* ---------------------------------------------------------------------- */

void VerletB200::device_setup(int flag, int output_flag)
{
  update->setupflag = 1;

  // host side: what must happen before atoms are handed over (same calls as Verlet::setup
  // up to the point where ghosts would be created -- ghosts and lists are device business)
  if (flag) {
    atom->setup();
    modify->setup_pre_exchange();
    // triclinic: the reference wraps and migrates in lamda coordinates, once (Verlet::setup,
    // verlet.cpp:111-128).  One process: the device does exactly that in its own setup, so the
    // host leaves the atoms alone (a second x -> lamda -> x round trip would move them by an ulp).
    const int tri_on_device = domain->triclinic && comm->nprocs == 1;
    if (domain->triclinic && !tri_on_device) domain->x2lamda(atom->nlocal);
    if (!tri_on_device) domain->pbc();
    domain->reset_box();
    comm->setup();
    if (neighbor->style) neighbor->setup_bins();
    if (!tri_on_device) comm->exchange();
    if (domain->triclinic && !tri_on_device) domain->lamda2x(atom->nlocal);
  }
  force->setup();
  ev_set(update->ntimestep);
  refuse_per_atom_tallies();
  // list options of neigh_modify the device build does not implement (neighbor.cpp:2727-2940);
  // checked here because Neighbor::init() runs after Integrate::init()
  if (neighbor->nex_mol)
    error->all(FLERR, "run_style verlet/b200 supports neigh_modify exclude type and group "
                      "(not molecule: atomic systems carry no molecule ids)");
  if (neighbor->includegroup)
    error->all(FLERR, "run_style verlet/b200 does not support neigh_modify include");
  if (neighbor->style != Neighbor::BIN)
    error->all(FLERR, "run_style verlet/b200 requires neighbor style bin");

  // device side: ghosts, bins, half list, forces (+ tallies)
  upload();
  for (int i = 0; i < pkg->nctx(); i++)
    B200_CHECK(pkg, b200_set_profiling(pkg->context(i), pkg->profile()));
  pkg->dev_setup(eflag ? 1 : 0, vflag ? 1 : 0);
  resident = 1;
  download(0);
  fill_per_atom_tallies();
  if (eflag || vflag) fetch_tallies();
  neighbor->ncalls = 0;
  neighbor->ndanger = 0;

  modify->setup(vflag);
  // a post-force fix (fix langevin/b200) has just changed the forces on the device (Fix::setup)
  if (pkg->host_stale) download(0);
  if (output_flag >= 0) output->setup(output_flag);
  update->setupflag = 0;
}

void VerletB200::setup(int flag)
{
  if (comm->me == 0 && screen) {
    fputs("Setting up Verlet/B200 run ...\n", screen);
    if (flag) {
      utils::print(screen, "  Unit style    : {}\n  Current step  : {}\n  Time step     : {}\n",
                   update->unit_style, update->ntimestep, update->dt);
      timer->print_timeout(screen);
    }
  }
  device_setup(1, flag);
}

void VerletB200::setup_minimal(int flag)
{
  device_setup(flag, -1);
}

void VerletB200::reset_dt()
{
  // fix nve/b200::reset_dt() forwards the new dtv/dtf itself
}

/* ----------------------------------------------------------------------
   `package b200 profile yes`: one timestep stage by stage, each followed by
   a device sync and the Timer stamp Verlet::run gives it (verlet.cpp:257-355),
   so that Finish prints the usual Pair / Neigh / Comm / Modify breakdown
------------------------------------------------------------------------- */

void VerletB200::step_by_stage(int ef, int vf)
{
  int nflag = 0;
  B200_CHECK(pkg, b200_initial_integrate(ctx));
  B200_CHECK(pkg, b200_sync(ctx));
  timer->stamp(Timer::MODIFY);
  B200_CHECK(pkg, b200_decide(ctx, &nflag));
  if (nflag) {
    B200_CHECK(pkg, b200_reneighbor(ctx));
    B200_CHECK(pkg, b200_sync(ctx));
    timer->stamp(Timer::NEIGH);
  } else {
    B200_CHECK(pkg, b200_forward_comm(ctx));
    B200_CHECK(pkg, b200_sync(ctx));
    timer->stamp(Timer::COMM);
  }
  B200_CHECK(pkg, b200_force_clear(ctx));
  B200_CHECK(pkg, b200_pair_compute(ctx, ef, vf));
  B200_CHECK(pkg, b200_sync(ctx));
  timer->stamp(Timer::PAIR);
  B200_CHECK(pkg, b200_reverse_comm(ctx));
  B200_CHECK(pkg, b200_sync(ctx));
  timer->stamp(Timer::COMM);
  B200_CHECK(pkg, b200_final_integrate(ctx));
  B200_CHECK(pkg, b200_sync(ctx));
  timer->stamp(Timer::MODIFY);
}

/* ----------------------------------------------------------------------
   a timestep whose integrator is a host fix with device loops (fix nvt/b200):
   Verlet::run's sequence (verlet.cpp:257-355) with the fix's own
   initial_integrate / final_integrate around the device stages
------------------------------------------------------------------------- */

void VerletB200::step_staged_fix(int ef, int vf)
{
  int nflag = 0;
  staged_fix->initial_integrate(vflag);
  timer->stamp(Timer::MODIFY);
  pkg->dev_decide(&nflag);
  if (nflag) {
    pkg->dev_reneighbor();
    timer->stamp(Timer::NEIGH);
  } else {
    pkg->dev_forward_comm();
    timer->stamp(Timer::COMM);
  }
  pkg->dev_force_clear();
  pkg->dev_pair_compute(ef, vf);
  timer->stamp(Timer::PAIR);
  pkg->dev_reverse_comm();
  timer->stamp(Timer::COMM);
  for (auto &fix : post_fixes) fix->post_force(vflag);
  // a barostat reads the pressure of this step (pair virial) inside final_integrate
  if (ef || vf) fetch_tallies();
  staged_fix->final_integrate();
  timer->stamp(Timer::MODIFY);
}

/* ----------------------------------------------------------------------
   fix nve/b200 with fixes that act on the forces after the pair stage (fix
   langevin/b200): the integrator cannot live inside the pair kernel, so the
   timestep runs stage by stage -- FixNVE::initial_integrate as its two loops
   (v += dtf/m f, then x += dtv v: fix_nve.cpp:92-101), Modify::post_force
   where Verlet::run has it (verlet.cpp:340-350), FixNVE::final_integrate
------------------------------------------------------------------------- */

void VerletB200::step_nve_post_force(int ef, int vf)
{
  int nflag = 0, groupbit;
  double dtv, dtf;
  bnve->b200_params(dtv, dtf, groupbit);
  pkg->dev_nve_v(dtf, groupbit);
  pkg->dev_nve_x(dtv, groupbit);
  timer->stamp(Timer::MODIFY);
  pkg->dev_decide(&nflag);
  if (nflag) {
    pkg->dev_reneighbor();
    timer->stamp(Timer::NEIGH);
  } else {
    pkg->dev_forward_comm();
    timer->stamp(Timer::COMM);
  }
  pkg->dev_force_clear();
  pkg->dev_pair_compute(ef, vf);
  timer->stamp(Timer::PAIR);
  pkg->dev_reverse_comm();
  timer->stamp(Timer::COMM);
  for (auto &fix : post_fixes) fix->post_force(vflag);
  pkg->dev_nve_v(dtf, groupbit);
  timer->stamp(Timer::MODIFY);
}

/* ----------------------------------------------------------------------
   run for N steps
------------------------------------------------------------------------- */

void VerletB200::run(int n)
{
  bigint ntimestep;
  const int by_stage = pkg->profile() && !pkg->group();

  for (int i = 0; i < n; i++) {
    if (timer->check_timeout(i)) {
      update->nsteps = i;
      break;
    }

    ntimestep = ++update->ntimestep;
    ev_set(ntimestep);
    refuse_per_atom_tallies();

    // one whole timestep on the device; asynchronous unless the rebuild vote or a tally
    // needs a word back (verlet.cpp:229-360 is the sequence it implements)
    timer->stamp();
    int rebuilt = 0;
    // from here on the device copy of the atoms is the current one (compute temp/b200 inside a
    // staged fix must already read the device sums on the first step of a run)
    pkg->host_stale = 1;
    if (bstaged)
      step_staged_fix(eflag ? 1 : 0, vflag ? 1 : 0);
    else if (!post_fixes.empty())
      step_nve_post_force(eflag ? 1 : 0, vflag ? 1 : 0);
    else if (by_stage)
      step_by_stage(eflag ? 1 : 0, vflag ? 1 : 0);
    else {
      // the device may run ahead (next step's initial_integrate inside this step's pair kernel)
      // when nothing on the host looks at the atoms before the next step: not on the last step,
      // not on an output step, not when a timeout may end the run here
      const int more = (i < n - 1) && ntimestep != output->next && !timer->has_timeout();
      pkg->dev_step(eflag ? 1 : 0, vflag ? 1 : 0, more, &rebuilt);
    }
    resident = 1;
    pkg->host_stale = 1;

    if (ntimestep == output->next) {
      // atoms come to the host only for consumers that read them: a dump, a restart, a
      // compute without a /b200 version.  A thermo step of device-aware computes does not.
      const bool need_atoms = !thermo_on_device || ntimestep == output->next_dump_any ||
          ntimestep == output->next_restart;
      if (need_atoms) download(0);
      fill_per_atom_tallies();
      if (eflag || vflag) fetch_tallies();
      if (!by_stage && !bstaged && post_fixes.empty()) timer->stamp(Timer::PAIR);
      output->write(ntimestep);
      timer->stamp(Timer::OUTPUT);
    }
  }
}

/* ----------------------------------------------------------------------
   neighbor statistics for Finish (finish.cpp): builds, dangerous builds and
   the stored-pair count through the pair style's NeighList
------------------------------------------------------------------------- */

void VerletB200::publish_neighbor_stats()
{
  b200_stats st;
  pkg->dev_stats(&st);
  neighbor->ncalls = st.nbuilds;
  neighbor->ndanger = st.ndanger;
  neighbor->ago = (int) st.ago;
  NeighList *list = force->pair ? force->pair->list : nullptr;
  if (list) {
    const int nlocal = atom->nlocal;
    list->grow(nlocal, nlocal + atom->nghost);
    // numneigh in the order the atoms were downloaded in (sub-domain after sub-domain)
    int off = 0;
    for (int i = 0; i < pkg->nctx(); i++) {
      int nl = 0, ng = 0;
      int64_t npairs = 0;
      B200_CHECK(pkg, b200_get_counts(pkg->context(i), &nl, &ng));
      B200_CHECK(pkg, b200_get_neighbor_list(pkg->context(i), list->numneigh + off, nullptr, 0, &npairs));
      off += nl;
    }
    for (int i = 0; i < nlocal; i++) list->ilist[i] = i;
    list->inum = nlocal;
    list->gnum = 0;
  }
}

void VerletB200::cleanup()
{
  if (resident) {
    // the device timed everything up to here: one sync so "Loop time" covers all queued work
    download(1);
    publish_neighbor_stats();
    if (pkg->profile() && comm->me == 0) {
      double ms[B200_NPHASE];
      int64_t calls[B200_NPHASE];
      B200_CHECK(pkg, b200_get_phase_times(pkg->context(0), ms, calls));
      static const char *names[B200_NPHASE] = {"nve initial", "nve final", "halo forward",
                                               "halo reverse", "pair", "neigh (all)",
                                               "neigh list build", "force clear", "tallies"};
      std::string mesg = "B200 device time by phase (CUDA events):\n";
      for (int k = 0; k < B200_NPHASE; k++)
        if (calls[k])
          mesg += fmt::format("  {:18s} {:10.3f} ms  {:8d} calls\n", names[k], ms[k], calls[k]);
      utils::logmesg(lmp, mesg);
    }
  }
  Verlet::cleanup();
}
