/* b200_md.h -- C ABI of the B200-native short-range MD hot path (libb200md.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The host
 * side that binds it is (a) the LAMMPS package src/B200 (lammps_b200/lammps_pkg/B200:
 * pair_style lj/cut/b200, eam/b200, fix nve/b200, run_style verlet/b200, fix B200) and
 * (b) the ctypes mirror lammps_b200/engine.py used by tests and bench.py.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference's src/).  All functions return 0 on success or a negative B200_E* code;
 * b200_last_error() gives the message (the host turns it into error->one(FLERR,...),
 * precedent GPU/fix_gpu.cpp:251,358-361).  A context is single-threaded like a LAMMPS
 * instance.  All host arrays are caller-owned; device memory is owned by the
 * context and released by b200_destroy (precedent: lmp_clear_device, GPU/fix_gpu.cpp:256).
 *
 * Per-atom host layouts are the reference's: x,v,f = double[n][3] row-major (atom.h:72-75),
 * type/tag/mask/image = int32 (LAMMPS_SMALLBIG, lmptype.h:99-125).  Type tables are
 * (ntypes+1)x(ntypes+1) row-major with row/column 0 unused, as memory->create gives them.
 */
#ifndef B200_MD_H
#define B200_MD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

enum {
  B200_OK = 0,
  B200_ECUDA = -1,     /* CUDA runtime / NCCL failure */
  B200_EARG = -2,      /* invalid argument or call order */
  B200_ECAPACITY = -3, /* neighbor list overflow beyond `one` (npair_bin.cpp:248), ghost buffer */
  B200_ENONFINITE = -4,/* non-numeric coordinates (nbin.cpp:145, domain.cpp:787-790) */
  B200_ELOST = -5      /* atom left the (non-periodic) box or moved further than a sub-domain */
};

enum { B200_PREC_DOUBLE = 0, B200_PREC_MIXED = 1 };

/* ---- lifetime: replaces lmp_init_device / lmp_clear_device (GPU/fix_gpu.cpp:44-55) and the
 *      `package b200` fix constructor (input.cpp:1736-1777 Input::package) */
int b200_create(b200_ctx **out, int device, int precision);
void b200_destroy(b200_ctx *ctx);
const char *b200_last_error(const b200_ctx *ctx);
/* 1 if a CUDA device is usable from this process (never falls back to the CPU) */
int b200_device_count(void);

/* ---- domain: Domain::set_global_box / set_local_box (domain.cpp), Comm::set_proc_grid
 *      (comm.cpp:505); sub-box = boxlo + prd*myloc/procgrid, last one closed at boxhi */
int b200_set_box(b200_ctx *ctx, const double boxlo[3], const double boxhi[3],
                 const int periodicity[3]);
/*      triclinic box (Domain::set_global_box domain.cpp:263-290: tilt factors xy, xz, yz;
 *      Domain::x2lamda / lamda2x :2347-2390): periodic wrap, migration and ghost slabs are decided
 *      in lamda coordinates (verlet.cpp:293-313, comm_brick.cpp:177-237), bins cover the bounding
 *      box (nbin_standard.cpp:86-112), the stencil is full and the half list follows the tag rule
 *      of NPairBin<HALF,NEWTON,TRI> (npair_bin.cpp:133-155; angstrom = Force::angstrom for its
 *      0.01 angstrom tolerance).  Sub-domains are bricks in lamda coordinates. */
int b200_set_box_triclinic(b200_ctx *ctx, const double boxlo[3], const double boxhi[3], double xy,
                           double xz, double yz, const int periodicity[3], double angstrom);
/*      `newton off` (Force::newton_pair = 0, force.cpp; list rule NPairBin<HALF,!NEWTON>
 *      npair_bin.cpp:126-131): owned-ghost pairs are stored and evaluated by both owners, nothing
 *      is scattered onto ghosts or sent back.  Runs on the bin-tile rows that hold every ghost
 *      partner (lj/cut, single-element eam); default is on. */
int b200_set_newton(b200_ctx *ctx, int newton_pair);
int b200_set_decomposition(b200_ctx *ctx, const int procgrid[3], const int myloc[3]);
/*      optional: which rank owns grid location (ix,iy,iz), n = px*py*pz entries indexed
 *      (ix*py+iy)*pz+iz -- Comm::grid2proc (comm.h); default = that index itself (MPI_Cart order) */
int b200_set_rank_grid(b200_ctx *ctx, const int *grid2rank, int n);

/* ---- neighbor / neigh_modify (neighbor.cpp:2680-2940): skin, every, delay, check, one.
 *      cutneighsq = (sqrt(cutsq)+skin)^2 is derived from the pair style's cutsq
 *      (neighbor.cpp:337-383), triggersq = (skin/2)^2 */
int b200_set_neighbor(b200_ctx *ctx, double skin, int every, int delay, int dist_check, int one);
/* neigh_modify once yes|no (Neighbor::decide never asks for a rebuild, neighbor.cpp:2420) and
 * neigh_modify exclude type i j ... (NPair::exclusion, npair.cpp:244-248): ex_type is Neighbor's
 * symmetric table of excluded type pairs, [(ntypes+1)^2] flags, NULL = no exclusions.
 * neigh_modify exclude group g1 g2 (npair.cpp:249-254): n pairs of group bits (Neighbor::ex1_bit,
 * ex2_bit; n <= 8); the group masks of ghosts owned by other sub-domains then travel with the
 * border exchange.  Molecule exclusions need molecule ids, which atomic systems do not have. */
int b200_neigh_modify(b200_ctx *ctx, int build_once, int ntypes, const int *ex_type);
int b200_neigh_modify_groups(b200_ctx *ctx, int n, const int *bit1, const int *bit2);

/* ---- atoms: Atom arrays (atom.h:72-75) + per-type mass (atom.cpp set_mass).
 *      mask/image may be NULL (all atoms in group `all`, image flags 0). */
int b200_set_atoms(b200_ctx *ctx, int nlocal, int ntypes, const double *mass /*[ntypes+1]*/,
                   const double *x, const double *v, const int *type, const int *tag,
                   const int *mask, const int *image);
/* copies nlocal (+nghost if with_ghosts) atoms in current device order; any pointer may be NULL.
 * This is the sync the host needs on thermo/dump steps. */
int b200_get_atoms(b200_ctx *ctx, int with_ghosts, double *x, double *v, double *f, int *type,
                   int *tag, int *mask, int *image);
int b200_get_counts(const b200_ctx *ctx, int *nlocal, int *nghost);

/* ---- pair styles.  PairLJCut::init_one products (pair_lj_cut.cpp:503-524) and
 *      force->special_lj; replaces ljl_gpu_init (GPU/pair_lj_cut_gpu.cpp:35-53). */
int b200_pair_lj_cut(b200_ctx *ctx, int ntypes, const double *cutsq, const double *lj1,
                     const double *lj2, const double *lj3, const double *lj4,
                     const double *offset, const double special_lj[4]);
/*      PairEAM tables after file2array()+array2spline() (pair_eam.cpp:999-1205,1492-1545):
 *      splines are [n][nr+1 or nrho+1][7]; replaces eam_gpu_init (GPU/pair_eam_gpu.cpp:35-56). */
int b200_pair_eam(b200_ctx *ctx, int ntypes, int nr, int nrho, double rdr, double rdrho,
                  double rhomax, double cutforcesq, const int *type2frho /*[ntypes+1]*/,
                  const int *type2rhor, const int *type2z2r, const double *scale, int nfrho,
                  const double *frho_spline, int nrhor, const double *rhor_spline, int nz2r,
                  const double *z2r_spline);

/* ---- fix nve: FixNVE::init (fix_nve.cpp:55-59): dtv = dt, dtf = 0.5*dt*ftm2v */
int b200_fix_nve(b200_ctx *ctx, double dtv, double dtf, int groupbit);

/* ---- timestep pieces, one per call the reference's Verlet makes (verlet.cpp:93-162, 229-360).
 *      b200_setup      = Verlet::setup: pbc, comm->setup, setup_bins, exchange, borders,
 *                        neighbor->build, force_clear, pair->compute, reverse_comm
 *      b200_run        = Verlet::run for nsteps (all device-side; host only reads the 4-byte
 *                        rebuild vote when dist_check is on).  eflag/vflag tallies happen on
 *                        steps where (ntimestep % thermo_every == 0) and on the last step
 *                        (integrate.cpp:106-151 ev_set); thermo_out receives, per tallied step,
 *                        10 doubles {step, sum(m v^2), eng_vdwl, virial[0..5], 0}.
 */
int b200_setup(b200_ctx *ctx, int eflag, int vflag);
/* device time of the last b200_run: CUDA events recorded on the context's stream before the
 * first and after the last kernel of the run (the "Loop time" of finish.cpp:117-160) */
int b200_last_run_ms(b200_ctx *ctx, double *ms);
int b200_run(b200_ctx *ctx, int nsteps, int64_t first_step, int thermo_every,
             double *thermo_out, int max_thermo, int *n_thermo);

/*      one iteration of Verlet::run (verlet.cpp:246-355) minus output: what run_style verlet/b200
 *      calls every timestep.  Asynchronous unless dist_check needs the rebuild vote or
 *      eflag/vflag ask for tallies (then eng_vdwl/virial are current for b200_get_tallies). */
int b200_step(b200_ctx *ctx, int eflag, int vflag, int *rebuilt);
/* b200_step for a host that knows what comes next (Verlet::run does: i < n-1, output->next):
 * more != 0 promises that another step follows before anything reads atoms, velocities or forces;
 * the engine may then fuse the next step's FixNVE::initial_integrate into this step's pair kernel. */
int b200_step_ahead(b200_ctx *ctx, int eflag, int vflag, int more, int *rebuilt);
/*      stage-granular entry points (the same stages, one call each) */
int b200_initial_integrate(b200_ctx *ctx);            /* FixNVE::initial_integrate fix_nve.cpp:68 */
int b200_final_integrate(b200_ctx *ctx);              /* FixNVE::final_integrate  fix_nve.cpp:112 */
int b200_decide(b200_ctx *ctx, int *rebuild);         /* Neighbor::decide neighbor.cpp:2408 */
int b200_forward_comm(b200_ctx *ctx);                 /* CommBrick::forward_comm comm_brick.cpp:485 */
int b200_reverse_comm(b200_ctx *ctx);                 /* CommBrick::reverse_comm comm_brick.cpp:545 */
int b200_reneighbor(b200_ctx *ctx);                   /* pbc+exchange+borders+Neighbor::build */
int b200_force_clear(b200_ctx *ctx);                  /* Verlet::force_clear verlet.cpp:376 */
int b200_pair_compute(b200_ctx *ctx, int eflag, int vflag); /* Pair::compute pair.h:159 */
/* ---- the per-atom loops of FixNH (fix nvt; fix_nh.cpp:916-1014, 2278-2352), for a host that runs
 *      the Nose-Hoover chain itself (the package's fix nvt/b200 inherits the reference's FixNH):
 *      nve_v: v += dtf/m f; nve_x: x += dtv v (+ the displacement check decide() reads);
 *      nh_v_temp: v *= factor.  Atoms in groupbit only. */
int b200_nve_v(b200_ctx *ctx, double dtf, int groupbit);
int b200_nve_x(b200_ctx *ctx, double dtv, int groupbit);
int b200_scale_v(b200_ctx *ctx, double factor, int groupbit);
/* fix npt / nph (FixNH with a barostat, orthogonal box): nh_v_press (fix_nh.cpp:2227-2252: each
 * velocity component scaled twice by factor[d]) and remap (fix_nh.cpp:1156-1300): owned atoms of
 * groupbit are dilated from the old box into the new one (Domain::x2lamda / lamda2x) and the
 * device adopts the new box -- halo shifts at once, bins / stencil / sub-domain bounds at the next
 * rebuild, the displacement check's trigger shrinks with the box corners (neighbor.cpp:2443). */
int b200_scale_v3(b200_ctx *ctx, const double factor[3], int groupbit);
int b200_remap(b200_ctx *ctx, const double oldlo[3], const double oldhi[3], const double newlo[3],
               const double newhi[3], int groupbit);

/* fix langevin (FixLangevin::post_force, fix_langevin.cpp:383-507; constant or equal-style target,
 * per-type masses, no bias / tally): on the stored forces of this step, for atoms in groupbit,
 *   f += gfactor1[type] v + gfactor2_tsqrt[type] (u - 0.5)
 * with the per-type prefactors [ntypes+1] of FixLangevin::init (:268-280), the second already
 * multiplied by sqrt(t_target).  u: three uniforms per atom and step from the device's
 * counter-based stream (Philox-4x32-10, key = seed, counter = (tag, step): independent of atom
 * order and of the number of sub-domains) -- or, when uniforms_by_tag != NULL, u =
 * uniforms_by_tag[3 (tag-1) + d] drawn by the host (verification against the reference's
 * sequential RanMars stream).  fsum != NULL: returns the summed random force of the group
 * (zero yes, :481-497; the host subtracts fsum/count with b200_add_force). */
int b200_langevin(b200_ctx *ctx, int ntypes, const double *gfactor1, const double *gfactor2_tsqrt,
                  int groupbit, uint64_t seed, int64_t step, const double *uniforms_by_tag,
                  int64_t nuniform, double *fsum);
int b200_add_force(b200_ctx *ctx, const double df[3], int groupbit);

/* ---- tallies the host reads back: pair->eng_vdwl, pair->virial[6] (pair.h), and
 *      sum_i m_i v_i^2 (ComputeTemp::compute_scalar, compute_temp.cpp:73-97) */
int b200_get_tallies(b200_ctx *ctx, double *eng_vdwl, double virial[6]);
int b200_ke_sum(b200_ctx *ctx, double *mv2);
/*      the same for any group bit, with the kinetic tensor sums m*(vx vx, vy vy, vz vz, vx vy,
 *      vx vz, vy vz) of ComputeTemp::compute_vector (compute_temp.cpp:100-140): what
 *      compute temp/b200 needs, so that a thermo step moves 56 bytes instead of the atoms */
int b200_ke_group(b200_ctx *ctx, int groupbit, double *mv2, double tensor[6]);
/*      wait for everything enqueued on the context (timer sync, Timer::stamp with _sync) */
int b200_sync(b200_ctx *ctx);
/* ---- run-time knobs by name = the keywords of `package b200` (input.cpp Input::package):
 *      list tile|flat|auto, tile "tx,ty,tz", overlap yes|no, graph yes|no, tpa 1|2|4|8,
 *      mixed_fx yes|no, tallies local|global (local: eng_vdwl / virial / ke stay per sub-domain
 *      for a host that reduces them itself, e.g. LAMMPS over MPI) */
int b200_set_option(b200_ctx *ctx, const char *key, const char *value);

/* ---- statistics: neighbor->ncalls / ndanger / ago, pair counts, phase timings */
typedef struct {
  int64_t nbuilds;      /* Neighbor list builds            (finish.cpp) */
  int64_t ndanger;      /* Dangerous builds                (neighbor.cpp:2488) */
  int64_t ago;
  int64_t npairs;       /* stored pairs in the current list (Total # of neighbors) */
  int64_t maxneigh;     /* capacity per atom currently allocated */
  int64_t max_numneigh; /* largest numneigh[i] in the current list */
  int64_t nbins[3];     /* global bins (nbinx,nbiny,nbinz) */
  int64_t mbins;        /* local bins incl. ghost shell */
  int64_t nstencil;
  int64_t launches;     /* kernels launched since create */
  double  device_bytes; /* device memory currently allocated */
  int64_t halo_transport; /* per-step halo: 0 none (one rank), 1 NCCL send/recv, 2 peer-memory stores */
  int64_t lanes_per_atom; /* lanes sharing one atom in the flat pair kernels (B200_TPA) */
  int64_t list_kind;      /* 0 flat int32 half list (RED scatter), 1 bin-tile list (kernels_tile.cuh) */
  int64_t tile[3];        /* bin-tile size in bins (tx,ty,tz); 0 for the flat list */
  int64_t list_entries;   /* entries stored per build: the half-list pairs (npairs) plus, for the
                             tile list, the transposed copies of owned-owned pairs */
  int64_t tile_stage_max; /* most atoms one tile stages in shared memory */
  int64_t tiles_interior; /* tiles that touch owned atoms only ... */
  int64_t tiles_boundary; /* ... and tiles that stage ghosts or own atoms the reverse halo adds to */
  int64_t halo_overlap;   /* 1: interior tiles run on a second stream beside the forward halo, the
                             boundary tiles and the reverse halo (multi-GPU, tile list, B200_OVERLAP) */
} b200_stats;
int b200_get_stats(b200_ctx *ctx, b200_stats *out);

/* ---- list layout.  Default: the bin-tile list for lj/cut (16-bit entries into a shared-memory
 *      staging of a tile of bins; every owned-owned pair is stored in both atoms' rows so forces
 *      need no atomics; the entries flagged FWD are exactly the reference's half/Newton-on list)
 *      and the flat int32 half list for eam.  Environment B200_LIST=flat|tile overrides,
 *      B200_TILE=tx,ty,tz sets the tile size in bins. */

/* ---- test hooks: the half list as CSR over current owned order.  j indexes
 *      owned [0,nlocal) or ghost [nlocal, nlocal+nghost) atoms, as NeighList does
 *      (neigh_list.h:53-57).  Call with neigh == NULL to get the pair count only. */
int b200_get_neighbor_list(b200_ctx *ctx, int *numneigh /*[nlocal]*/, int *neigh, int64_t cap,
                           int64_t *npairs);
/*      EAM intermediates rho[], fp[] (pair_eam.h) for owned(+ghost) atoms */
int b200_get_eam_rho_fp(b200_ctx *ctx, int with_ghosts, double *rho, double *fp);

/* ---- per-atom energy / virial of the pair style: Pair::ev_tally's eatom[] and vatom[][6]
 *      (pair.cpp:1087-1182, order xx,yy,zz,xy,xz,yz; what compute pe/atom and stress/atom read,
 *      and what the reference's pair unit test requests, test_pair_style.cpp:143).  With Newton
 *      on each atom of a pair receives half of the pair term; ghost shares are returned to their
 *      owners (ComputePEAtom's reverse_comm).  nlocal values in current device order, like
 *      b200_get_atoms.  Call right after a setup/step that tallied (eflag or vflag set); either
 *      pointer may be NULL.  Collective over the sub-domains of a multi-GPU run. */
int b200_pair_peratom(b200_ctx *ctx, double *eatom /*[nlocal]*/, double *vatom /*[nlocal][6]*/);

/* ---- per-phase device timing (CUDA events recorded on the context's stream around each
 *      phase inside b200_run; replaces Timer::stamp of verlet.cpp:257-357).  Off by default. */
enum {
  B200_PH_INITIAL = 0,  /* fix nve initial_integrate (+ displacement check)   Timer::MODIFY */
  B200_PH_FINAL = 1,    /* fix nve final_integrate                            Timer::MODIFY */
  B200_PH_FORWARD = 2,  /* ghost position update                              Timer::COMM   */
  B200_PH_REVERSE = 3,  /* ghost force reduction                              Timer::COMM   */
  B200_PH_PAIR = 4,     /* pair->compute incl. EAM rho/fp halo                Timer::PAIR   */
  B200_PH_NEIGH = 5,    /* pbc + exchange + borders + bin + list build        Timer::NEIGH  */
  B200_PH_BUILD = 6,    /* the list-build kernel alone (subset of NEIGH)                   */
  B200_PH_CLEAR = 7,    /* force_clear                                                     */
  B200_PH_THERMO = 8,   /* energy/virial/ke reductions on tallied steps                    */
  B200_NPHASE = 9
};
int b200_set_profiling(b200_ctx *ctx, int on);
/* accumulated since the last call (which resets them): milliseconds and launch counts */
int b200_get_phase_times(b200_ctx *ctx, double ms[B200_NPHASE], int64_t calls[B200_NPHASE]);

/* ---- multi-GPU: one context per GPU/process = one brick sub-domain (CommBrick,
 *      comm_brick.cpp:172-430).  The ghost halo (forward x, reverse f, EAM rho/fp), border
 *      construction and atom migration run as grouped ncclSend/ncclRecv between the 26
 *      neighbour sub-domains, from device buffers.  b200_comm_unique_id fills a 128-byte
 *      ncclUniqueId on rank 0 (the host broadcasts it: MPI_Bcast in a LAMMPS+MPI build,
 *      torch.distributed in bench.py); b200_comm_init joins the communicator.  Call order:
 *      b200_create, b200_comm_init, b200_set_box, b200_set_decomposition, b200_set_atoms
 *      (atoms this rank owns: sublo <= x < subhi), ... */
int b200_comm_unique_id(void *id128);
/*      host-only helper (no GPU needed): rank of the neighbour sub-domain in each of the 27
 *      directions dir = (dz+1)*9 + (dy+1)*3 + (dx+1); ranks are numbered like MPI_Cart_create
 *      does for the reference (last dimension fastest, procmap.cpp:361-374); -1 = no
 *      neighbour beyond a non-periodic boundary.  Messages between two ranks are issued in
 *      ascending direction order on both sides. */
int b200_neighbor_ranks(const int procgrid[3], const int myloc[3], const int periodicity[3],
                        int nbr[27]);
int b200_comm_init(b200_ctx *ctx, int nranks, int rank, const void *id128);

/* ---- several sub-domains in ONE process: `package b200 gpus N` (precedent for one process
 *      owning several devices: GPU/fix_gpu.cpp:123-247, lib/gpu/lal_device.cpp; there is no MPI
 *      in this image).  A group is N contexts -- one brick sub-domain each, on N
 *      devices or sharing devices -- with one host thread per context; what CommBrick does with
 *      MPI (comm_brick.cpp:172-430, 599-899) happens between them through peer memory: counts
 *      through shared host memory, migration/border records with cudaMemcpyAsync between the
 *      contexts' buffers, the per-step halo with the same peer-store kernels as between
 *      processes.  Per-context settings (b200_set_box, b200_set_neighbor, b200_pair_*,
 *      b200_fix_nve) are made on every b200_group_context(); everything that communicates goes
 *      through the b200_group_* calls below.  Atoms are handed over for the whole box and come
 *      back concatenated in sub-domain order. */
typedef struct b200_group b200_group;
int b200_group_create(b200_group **out, int nsub, const int *devices /*[nsub]*/, int precision);
void b200_group_destroy(b200_group *g);
const char *b200_group_last_error(const b200_group *g);
int b200_group_size(const b200_group *g);
b200_ctx *b200_group_context(b200_group *g, int i);
/*      sub-domain grid: the factorisation of n with the smallest surface (ProcMap::onelevel_grid,
 *      procmap.cpp:48); sub-domain i sits at grid location (i/(py*pz), (i/pz)%py, i%pz) */
int b200_group_auto_grid(int n, const double prd[3], int grid[3]);
int b200_group_set_grid(b200_group *g, const int grid[3]);
int b200_group_set_atoms(b200_group *g, int n, int ntypes, const double *mass, const double *x,
                         const double *v, const int *type, const int *tag, const int *mask,
                         const int *image);
int b200_group_count(b200_group *g, int *nlocal_total, int *nghost_total);
int b200_group_pair_peratom(b200_group *g, double *eatom, double *vatom);
int b200_group_get_atoms(b200_group *g, double *x, double *v, double *f, int *type, int *tag,
                         int *mask, int *image);
int b200_group_setup(b200_group *g, int eflag, int vflag);
int b200_group_step(b200_group *g, int eflag, int vflag, int *rebuilt);
/* the stages of a timestep one by one (b200_decide ... b200_reverse_comm, b200_nve_v / _x /
 * b200_scale_v) for every sub-domain of the group, in step */
int b200_group_decide(b200_group *g, int *rebuild);
int b200_group_reneighbor(b200_group *g);
int b200_group_forward_comm(b200_group *g);
int b200_group_force_clear(b200_group *g);
int b200_group_pair_compute(b200_group *g, int eflag, int vflag);
int b200_group_reverse_comm(b200_group *g);
int b200_group_nve_v(b200_group *g, double dtf, int groupbit);
int b200_group_nve_x(b200_group *g, double dtv, int groupbit);
int b200_group_scale_v(b200_group *g, double factor, int groupbit);
int b200_group_scale_v3(b200_group *g, const double factor[3], int groupbit);
int b200_group_langevin(b200_group *g, int ntypes, const double *gfactor1, const double *gfactor2_tsqrt,
                        int groupbit, uint64_t seed, int64_t step, const double *uniforms_by_tag,
                        int64_t nuniform, double *fsum);
int b200_group_add_force(b200_group *g, const double df[3], int groupbit);
int b200_group_remap(b200_group *g, const double oldlo[3], const double oldhi[3], const double newlo[3],
                     const double newhi[3], int groupbit);
int b200_group_step_ahead(b200_group *g, int eflag, int vflag, int more, int *rebuilt);
int b200_group_run(b200_group *g, int nsteps, int64_t first_step, int thermo_every,
                   double *thermo_out, int max_thermo, int *n_thermo);
int b200_group_get_tallies(b200_group *g, double *eng_vdwl, double virial[6]);
int b200_group_ke_sum(b200_group *g, double *mv2);
int b200_group_ke_group(b200_group *g, int groupbit, double *mv2, double tensor[6]);
int b200_group_last_run_ms(b200_group *g, double *ms);
int b200_group_get_stats(b200_group *g, b200_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* B200_MD_H */
