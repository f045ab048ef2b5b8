/* md_oracle.c -- CPU restatement of the reference's short-range MD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA engine in
 * lammps_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it; the product path never does.
 *
 * Scope: one process, orthogonal or triclinic box, atom_style atomic, newton on or off, the
 * half/bin/atomonly lists (newton, newton/tri, newtoff), neigh_modify exclude type / group, pair
 * lj/cut or eam (funcfl / setfl / fs tables built by lammps_b200/eam.py), fix nve.  Each function cites the reference
 * file:line (relative to /root/reference/src) whose arithmetic it restates.  The code keeps
 * the reference's operation ORDER (atom order, swap order, neighbour order) so that results
 * can be compared bit-for-bit with oracle/_ref (the compiled reference).
 *
 * PINNED (tests/test_oracle_golden.py, tests/test_oracle_reference_yaml.py): against fixtures
 * generated from the compiled reference (pair sets, forces, thermo), against the reference's
 * published bench logs (bench/log.15Jul25.{lj,eam}.fixed.g++.1), against
 * unittest/cplusplus/test_neighbor_class.cpp:235-269, and against the reference's known-answer
 * vectors unittest/force-styles/tests/atomic-pair-eam{,_alloy,_fs}.yaml at their own epsilon.
 * Triclinic boxes, newton off and group exclusions are pinned against oracle/_ref live
 * (tests/test_oracle_tri_newton_live.py: pair sets, forces, energy, pxy, 40-step trajectories).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/libmd_oracle.so oracle/md_oracle.c -lm
 * (-ffp-contract=off: the reference is built for baseline x86-64, i.e. without FMA contraction)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BIG 1.0e20
#define SMALL 1.0e-6       /* nbin_standard.cpp:27 */
#define IMGMASK 1023       /* lmptype.h:122-125 (LAMMPS_SMALLBIG) */
#define IMGBITS 10
#define IMG2BITS 20
#define MAXSWAP 24 /* 2 * maxneed swaps per dimension, maxneed <= 4 */

#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

typedef struct {
  /* domain (domain.h) */
  double boxlo[3], boxhi[3], prd[3];
  int periodic[3];
  /* triclinic box (domain.cpp:263-290): tilt factors, h = {xprd,yprd,zprd,yz,xz,xy}, its inverse,
     the bounding box of the tilted cell; angstrom = Force::angstrom (npair_bin.cpp:59) */
  int triclinic;
  double xy, xz, yz, h[6], h_inv[6], boxlo_bound[3], boxhi_bound[3], angstrom;
  /* Force::newton_pair (force.cpp); neigh_modify exclude group (neighbor.h:70-72) */
  int newton_pair;
  int nex_group, ex1_bit[8], ex2_bit[8];
  /* atoms (atom.h:72-75) */
  int nlocal, nghost, nmax, ntypes;
  double *x, *v, *f; /* [nmax][3] row-major */
  int *type, *tag, *mask, *image;
  double *mass; /* [ntypes+1] */
  /* neighbor settings (neighbor.cpp:300-390) */
  double skin, cutneighmax, cutneighmaxsq, triggersq;
  double *cutneighsq; /* [(ntypes+1)^2] */
  int every, delay, dist_check;
  int build_once;  /* neigh_modify once yes (neighbor.cpp:2420) */
  int *ex_type;    /* neigh_modify exclude type: [(ntypes+1)^2] flags or NULL (neighbor.h:66-68) */
  int ago, ncalls, ndanger;
  double *xhold;
  /* bins (nbin_standard.cpp) */
  int nbinx, nbiny, nbinz, mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo, mbins;
  double binsizex, binsizey, binsizez, bininvx, bininvy, bininvz;
  int *binhead, *bins, *atom2bin;
  int maxbin, maxbinatom;
  /* stencil (nstencil_bin.cpp) */
  int nstencil, *stencil;
  /* comm (comm_brick.cpp) single proc: every swap is with self */
  double cutghost;
  double cutghost3[3]; /* CommBrick::cutghost[]: box units, or lamda units in a triclinic box */
  int nswap, sendnum[MAXSWAP], firstrecv[MAXSWAP], pbc_flag[MAXSWAP], pbc[MAXSWAP][6];
  double slablo[MAXSWAP], slabhi[MAXSWAP];
  int *sendlist[MAXSWAP], maxsendlist[MAXSWAP];
  /* neighbor list (neigh_list.h:53-57) in CSR form */
  int inum, *numneigh;
  int64_t *firstneigh, neighcap, nneigh;
  int *neigh;
  /* pair */
  int pair_style; /* 1 lj/cut, 2 eam */
  double *cutsq, *lj1, *lj2, *lj3, *lj4, *offset; /* [(ntypes+1)^2] */
  double special_lj[4];
  /* eam tables (pair_eam.h) */
  int nr, nrho, nfrho, nrhor, nz2r;
  double rdr, rdrho, rhomax, cutforcesq;
  int *type2frho, *type2rhor, *type2z2r; /* [ntypes+1], [(ntypes+1)^2] x2 */
  double *scale;                         /* [(ntypes+1)^2] */
  double *frho_spline, *rhor_spline, *z2r_spline; /* [n][nr+1][7] */
  double *rho, *fp;
  int *numforce;
  int exceeded_rhomax;
  /* tallies (pair.h) */
  double eng_vdwl, virial[6];
  /* fix nve (fix_nve.cpp:55-59) */
  double dtv, dtf;
  int groupbit;
} Orc;

static void die(const char *msg) {
  fprintf(stderr, "md_oracle: %s\n", msg);
  abort();
}

static void *xrealloc(void *p, size_t n) {
  void *q = realloc(p, n ? n : 1);
  if (!q) die("out of memory");
  return q;
}

static void grow_atoms(Orc *o, int n) {
  if (n <= o->nmax) return;
  int nmax = MAX(n, o->nmax + o->nmax / 2 + 1024);
  o->x = xrealloc(o->x, sizeof(double) * 3 * nmax);
  o->v = xrealloc(o->v, sizeof(double) * 3 * nmax);
  o->f = xrealloc(o->f, sizeof(double) * 3 * nmax);
  o->type = xrealloc(o->type, sizeof(int) * nmax);
  o->tag = xrealloc(o->tag, sizeof(int) * nmax);
  o->mask = xrealloc(o->mask, sizeof(int) * nmax);
  o->image = xrealloc(o->image, sizeof(int) * nmax);
  o->xhold = xrealloc(o->xhold, sizeof(double) * 3 * nmax);
  o->bins = xrealloc(o->bins, sizeof(int) * nmax);
  o->atom2bin = xrealloc(o->atom2bin, sizeof(int) * nmax);
  o->numneigh = xrealloc(o->numneigh, sizeof(int) * nmax);
  o->firstneigh = xrealloc(o->firstneigh, sizeof(int64_t) * nmax);
  o->rho = xrealloc(o->rho, sizeof(double) * nmax);
  o->fp = xrealloc(o->fp, sizeof(double) * nmax);
  o->numforce = xrealloc(o->numforce, sizeof(int) * nmax);
  o->nmax = nmax;
}

/* ------------------------------------------------------------------ API: lifetime */

Orc *orc_create(void) {
  Orc *o = (Orc *)calloc(1, sizeof(Orc));
  if (o) { o->newton_pair = 1; o->angstrom = 1.0; }
  return o;
}

void orc_destroy(Orc *o) {
  if (!o) return;
  free(o->x); free(o->v); free(o->f); free(o->type); free(o->tag); free(o->mask);
  free(o->image); free(o->mass); free(o->cutneighsq); free(o->xhold); free(o->binhead);
  free(o->bins); free(o->atom2bin); free(o->stencil); free(o->numneigh);
  free(o->firstneigh); free(o->neigh); free(o->cutsq); free(o->lj1); free(o->lj2);
  free(o->lj3); free(o->lj4); free(o->offset); free(o->type2frho); free(o->type2rhor);
  free(o->type2z2r); free(o->scale); free(o->frho_spline); free(o->rhor_spline);
  free(o->z2r_spline); free(o->rho); free(o->fp); free(o->numforce); free(o->ex_type);
  for (int i = 0; i < MAXSWAP; i++) free(o->sendlist[i]);
  free(o);
}

void orc_set_box(Orc *o, const double *lo, const double *hi, const int *periodic) {
  for (int d = 0; d < 3; d++) {
    o->boxlo[d] = lo[d];
    o->boxhi[d] = hi[d];
    o->prd[d] = hi[d] - lo[d]; /* domain.cpp set_global_box: xprd = boxhi - boxlo */
    o->periodic[d] = periodic[d];
  }
  o->triclinic = 0;
  o->xy = o->xz = o->yz = 0.0;
}

/* Domain::set_global_box for a triclinic cell, domain.cpp:263-290 */
void orc_set_box_triclinic(Orc *o, const double *lo, const double *hi, double xy, double xz, double yz,
                           const int *periodic, double angstrom) {
  orc_set_box(o, lo, hi, periodic);
  o->triclinic = 1;
  o->xy = xy; o->xz = xz; o->yz = yz;
  o->angstrom = angstrom;
  double *h = o->h, *h_inv = o->h_inv;
  h[0] = o->prd[0]; h[1] = o->prd[1]; h[2] = o->prd[2];
  h_inv[0] = 1.0 / h[0]; h_inv[1] = 1.0 / h[1]; h_inv[2] = 1.0 / h[2];
  h[3] = yz; h[4] = xz; h[5] = xy;
  h_inv[3] = -h[3] / (h[1] * h[2]);
  h_inv[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
  h_inv[5] = -h[5] / (h[0] * h[1]);
  o->boxlo_bound[0] = MIN(o->boxlo[0], o->boxlo[0] + xy);
  o->boxlo_bound[0] = MIN(o->boxlo_bound[0], o->boxlo_bound[0] + xz);
  o->boxlo_bound[1] = MIN(o->boxlo[1], o->boxlo[1] + yz);
  o->boxlo_bound[2] = o->boxlo[2];
  o->boxhi_bound[0] = MAX(o->boxhi[0], o->boxhi[0] + xy);
  o->boxhi_bound[0] = MAX(o->boxhi_bound[0], o->boxhi_bound[0] + xz);
  o->boxhi_bound[1] = MAX(o->boxhi[1], o->boxhi[1] + yz);
  o->boxhi_bound[2] = o->boxhi[2];
}

/* the `newton` command (force.cpp): newton_pair on (default) or off */
void orc_set_newton(Orc *o, int newton_pair) { o->newton_pair = newton_pair ? 1 : 0; }

/* neigh_modify exclude group: n pairs of group bits (Neighbor::ex1_bit / ex2_bit) */
void orc_neigh_modify_groups(Orc *o, int n, const int *bit1, const int *bit2) {
  if (n > 8) die("too many group exclusions");
  o->nex_group = n;
  for (int m = 0; m < n; m++) { o->ex1_bit[m] = bit1[m]; o->ex2_bit[m] = bit2[m]; }
}

/* Domain::x2lamda / lamda2x for the first n atoms, domain.cpp:2347-2390 */
static void x2lamda(Orc *o, int n) {
  const double *h_inv = o->h_inv, *boxlo = o->boxlo;
  for (int i = 0; i < n; i++) {
    double *x = &o->x[3 * i], delta[3];
    delta[0] = x[0] - boxlo[0];
    delta[1] = x[1] - boxlo[1];
    delta[2] = x[2] - boxlo[2];
    x[0] = h_inv[0] * delta[0] + h_inv[5] * delta[1] + h_inv[4] * delta[2];
    x[1] = h_inv[1] * delta[1] + h_inv[3] * delta[2];
    x[2] = h_inv[2] * delta[2];
  }
}
static void lamda2x_one(const Orc *o, const double *lamda, double *x) {
  const double *h = o->h, *boxlo = o->boxlo;
  double l0 = lamda[0], l1 = lamda[1], l2 = lamda[2];
  x[0] = h[0] * l0 + h[5] * l1 + h[4] * l2 + boxlo[0];
  x[1] = h[1] * l1 + h[3] * l2 + boxlo[1];
  x[2] = h[2] * l2 + boxlo[2];
}
static void lamda2x(Orc *o, int n) {
  for (int i = 0; i < n; i++) lamda2x_one(o, &o->x[3 * i], &o->x[3 * i]);
}

void orc_set_atoms(Orc *o, int n, int ntypes, const double *mass, const double *x,
                   const double *v, const int *type, const int *tag, const int *mask,
                   const int *image) {
  grow_atoms(o, n);
  o->nlocal = n;
  o->nghost = 0;
  o->ntypes = ntypes;
  o->mass = xrealloc(o->mass, sizeof(double) * (ntypes + 1));
  memcpy(o->mass, mass, sizeof(double) * (ntypes + 1));
  memcpy(o->x, x, sizeof(double) * 3 * n);
  memcpy(o->v, v, sizeof(double) * 3 * n);
  memset(o->f, 0, sizeof(double) * 3 * n);
  memcpy(o->type, type, sizeof(int) * n);
  memcpy(o->tag, tag, sizeof(int) * n);
  for (int i = 0; i < n; i++) o->mask[i] = mask ? mask[i] : 1;
  /* default image flags: (512,512,512) packed, atom_vec.cpp create_atom */
  for (int i = 0; i < n; i++)
    o->image[i] = image ? image[i] : ((512 << IMG2BITS) | (512 << IMGBITS) | 512);
}

void orc_set_neighbor(Orc *o, double skin, int every, int delay, int dist_check) {
  o->skin = skin;
  o->every = every;
  o->delay = delay;
  o->dist_check = dist_check;
}

/* neigh_modify once yes|no and exclude type i j ... (neighbor.cpp:2727-2790; Neighbor::init builds
 * the symmetric ex_type table, neighbor.cpp:560-575).  ex_type: [(ntypes+1)^2] flags or NULL. */
void orc_neigh_modify(Orc *o, int build_once, int ntypes, const int *ex_type) {
  o->build_once = build_once;
  free(o->ex_type);
  o->ex_type = NULL;
  if (ex_type) {
    int n2 = (ntypes + 1) * (ntypes + 1);
    o->ex_type = (int *)malloc(sizeof(int) * n2);
    memcpy(o->ex_type, ex_type, sizeof(int) * n2);
  }
}

void orc_fix_nve(Orc *o, double dt, double ftm2v, int groupbit) {
  o->dtv = dt;                /* fix_nve.cpp:57 */
  o->dtf = 0.5 * dt * ftm2v;  /* fix_nve.cpp:58 */
  o->groupbit = groupbit;
}

static double *dup_table(const double *src, int n) {
  double *p = malloc(sizeof(double) * n);
  memcpy(p, src, sizeof(double) * n);
  return p;
}

/* pair_lj_cut.cpp:503-559 init_one products are passed in ready-made */
void orc_pair_lj_cut(Orc *o, int ntypes, const double *cutsq, const double *lj1,
                     const double *lj2, const double *lj3, const double *lj4,
                     const double *offset, const double *special_lj) {
  int n2 = (ntypes + 1) * (ntypes + 1);
  o->pair_style = 1;
  o->ntypes = ntypes;
  free(o->cutsq); free(o->lj1); free(o->lj2); free(o->lj3); free(o->lj4); free(o->offset);
  o->cutsq = dup_table(cutsq, n2);
  o->lj1 = dup_table(lj1, n2);
  o->lj2 = dup_table(lj2, n2);
  o->lj3 = dup_table(lj3, n2);
  o->lj4 = dup_table(lj4, n2);
  o->offset = dup_table(offset, n2);
  for (int i = 0; i < 4; i++) o->special_lj[i] = special_lj ? special_lj[i] : 1.0;
}

/* pair_eam.cpp:1492-1545 array2spline output is passed in ready-made (see eam tables in
   lammps_b200/eam.py and oracle/eam_tables.py, both pinned against the reference). */
void orc_pair_eam(Orc *o, int ntypes, int nr, int nrho, double rdr, double rdrho,
                  double rhomax, double cutforcesq, const int *type2frho,
                  const int *type2rhor, const int *type2z2r, const double *scale, int nfrho,
                  const double *frho_spline, int nrhor, const double *rhor_spline, int nz2r,
                  const double *z2r_spline) {
  int n1 = ntypes + 1, n2 = n1 * n1;
  o->pair_style = 2;
  o->ntypes = ntypes;
  o->nr = nr; o->nrho = nrho; o->rdr = rdr; o->rdrho = rdrho;
  o->rhomax = rhomax; o->cutforcesq = cutforcesq;
  o->nfrho = nfrho; o->nrhor = nrhor; o->nz2r = nz2r;
  free(o->type2frho); free(o->type2rhor); free(o->type2z2r); free(o->scale);
  free(o->frho_spline); free(o->rhor_spline); free(o->z2r_spline); free(o->cutsq);
  o->type2frho = malloc(sizeof(int) * n1);
  memcpy(o->type2frho, type2frho, sizeof(int) * n1);
  o->type2rhor = malloc(sizeof(int) * n2);
  memcpy(o->type2rhor, type2rhor, sizeof(int) * n2);
  o->type2z2r = malloc(sizeof(int) * n2);
  memcpy(o->type2z2r, type2z2r, sizeof(int) * n2);
  o->scale = dup_table(scale, n2);
  o->frho_spline = dup_table(frho_spline, nfrho * (nrho + 1) * 7);
  o->rhor_spline = dup_table(rhor_spline, nrhor * (nr + 1) * 7);
  o->z2r_spline = dup_table(z2r_spline, nz2r * (nr + 1) * 7);
  /* pair.cpp:226-317 Pair::init: cutsq[i][j] = init_one()^2 = cutmax^2 for every pair */
  o->cutsq = malloc(sizeof(double) * n2);
  for (int i = 0; i < n2; i++) o->cutsq[i] = cutforcesq;
  o->exceeded_rhomax = 0;
}

/* ------------------------------------------------------------------ neighbor init */

/* neighbor.cpp:337-383 */
static void neighbor_init(Orc *o) {
  int n = o->ntypes, n1 = n + 1;
  o->triggersq = 0.25 * o->skin * o->skin;
  o->cutneighsq = xrealloc(o->cutneighsq, sizeof(double) * n1 * n1);
  o->cutneighmax = 0.0;
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      double cutoff = sqrt(o->cutsq[i * n1 + j]);
      double delta = cutoff > 0.0 ? o->skin : 0.0;
      double cut = cutoff + delta;
      o->cutneighsq[i * n1 + j] = cut * cut;
      o->cutneighmax = MAX(o->cutneighmax, cut);
    }
  o->cutneighmaxsq = o->cutneighmax * o->cutneighmax;
}

/* ------------------------------------------------------------------ domain */

/* domain.cpp:769-887 Domain::pbc (orthogonal, no deform) */
void orc_pbc(Orc *o) {
  double *lo = o->boxlo, *hi = o->boxhi, *period = o->prd;
  /* triclinic: atoms are in lamda coordinates here (verlet.cpp:293); boxlo_lamda = 0,
     boxhi_lamda = 1, prd_lamda = 1 (Domain::set_lamda_box) */
  static double lo_lamda[3] = {0.0, 0.0, 0.0}, hi_lamda[3] = {1.0, 1.0, 1.0}, prd_lamda[3] = {1.0, 1.0, 1.0};
  if (o->triclinic) { lo = lo_lamda; hi = hi_lamda; period = prd_lamda; }
  for (int i = 0; i < o->nlocal; i++) {
    double *x = &o->x[3 * i];
    int idim, otherdims;
    if (o->periodic[0]) {
      if (x[0] < lo[0]) {
        x[0] += period[0];
        idim = o->image[i] & IMGMASK;
        otherdims = o->image[i] ^ idim;
        idim--; idim &= IMGMASK;
        o->image[i] = otherdims | idim;
      }
      if (x[0] >= hi[0]) {
        x[0] -= period[0];
        x[0] = MAX(x[0], lo[0]);
        idim = o->image[i] & IMGMASK;
        otherdims = o->image[i] ^ idim;
        idim++; idim &= IMGMASK;
        o->image[i] = otherdims | idim;
      }
    }
    if (o->periodic[1]) {
      if (x[1] < lo[1]) {
        x[1] += period[1];
        idim = (o->image[i] >> IMGBITS) & IMGMASK;
        otherdims = o->image[i] ^ (idim << IMGBITS);
        idim--; idim &= IMGMASK;
        o->image[i] = otherdims | (idim << IMGBITS);
      }
      if (x[1] >= hi[1]) {
        x[1] -= period[1];
        x[1] = MAX(x[1], lo[1]);
        idim = (o->image[i] >> IMGBITS) & IMGMASK;
        otherdims = o->image[i] ^ (idim << IMGBITS);
        idim++; idim &= IMGMASK;
        o->image[i] = otherdims | (idim << IMGBITS);
      }
    }
    if (o->periodic[2]) {
      if (x[2] < lo[2]) {
        x[2] += period[2];
        idim = ((unsigned)o->image[i]) >> IMG2BITS;
        otherdims = o->image[i] ^ (idim << IMG2BITS);
        idim--; idim &= IMGMASK;
        o->image[i] = otherdims | (idim << IMG2BITS);
      }
      if (x[2] >= hi[2]) {
        x[2] -= period[2];
        x[2] = MAX(x[2], lo[2]);
        idim = ((unsigned)o->image[i]) >> IMG2BITS;
        otherdims = o->image[i] ^ (idim << IMG2BITS);
        idim++; idim &= IMGMASK;
        o->image[i] = otherdims | (idim << IMG2BITS);
      }
    }
  }
}

/* ------------------------------------------------------------------ comm */

/* comm_brick.cpp:172-430 CommBrick::setup, 1x1x1 processor grid, mode SINGLE.
   maxneed[d] = int(cutghost*1/prd)+1 (comm_brick.cpp:267-269): more than one layer of periodic
   images when the box edge is shorter than the ghost cutoff (the 7 A box of the reference's
   atomic-pair-eam.yaml); non-periodic dims get maxneed = min(maxneed, procgrid-1) = 0 swaps.
   Swaps beyond the first pair of a dimension forward only ghosts received by the previous
   pair, restricted to one half of the sub-box (slab bound at its middle, :385-411). */
static void comm_setup(Orc *o) {
  o->cutghost = o->cutneighmax; /* comm.cpp:683 with no user cutoff */
  /* comm_brick.cpp:220-237: box coordinates, or for a triclinic box lamda coordinates with the
     cutoff as a distance between lamda planes */
  double prd[3], lo3[3], hi3[3];
  for (int d = 0; d < 3; d++) {
    prd[d] = o->triclinic ? 1.0 : o->prd[d];
    lo3[d] = o->triclinic ? 0.0 : o->boxlo[d];
    hi3[d] = o->triclinic ? 1.0 : o->boxhi[d];
    o->cutghost3[d] = o->cutghost;
  }
  if (o->triclinic) {
    const double *h_inv = o->h_inv;
    double length0 = sqrt(h_inv[0] * h_inv[0] + h_inv[5] * h_inv[5] + h_inv[4] * h_inv[4]);
    o->cutghost3[0] = o->cutghost * length0;
    double length1 = sqrt(h_inv[1] * h_inv[1] + h_inv[3] * h_inv[3]);
    o->cutghost3[1] = o->cutghost * length1;
    double length2 = h_inv[2];
    o->cutghost3[2] = o->cutghost * length2;
  }
  int iswap = 0;
  for (int dim = 0; dim < 3; dim++) {
    int maxneed = (int)(o->cutghost3[dim] * 1 / prd[dim]) + 1;
    if (!o->periodic[dim]) maxneed = MIN(maxneed, 0);
    if (iswap + 2 * maxneed > MAXSWAP) die("box edge much shorter than the ghost cutoff: too many swaps");
    double sublo = lo3[dim], subhi = hi3[dim];
    for (int ineed = 0; ineed < 2 * maxneed; ineed++) {
      o->pbc_flag[iswap] = 0;
      for (int k = 0; k < 6; k++) o->pbc[iswap][k] = 0;
      int sign;
      if (ineed % 2 == 0) {
        o->slablo[iswap] = ineed < 2 ? -BIG : 0.5 * (sublo + subhi);
        o->slabhi[iswap] = sublo + o->cutghost3[dim];
        o->pbc_flag[iswap] = 1; /* myloc == 0 */
        sign = 1;
      } else {
        o->slablo[iswap] = subhi - o->cutghost3[dim];
        o->slabhi[iswap] = ineed < 2 ? BIG : 0.5 * (sublo + subhi);
        o->pbc_flag[iswap] = 1; /* myloc == procgrid-1 */
        sign = -1;
      }
      o->pbc[iswap][dim] = sign;
      if (o->triclinic) { /* comm_brick.cpp:396-420 */
        if (dim == 1) o->pbc[iswap][5] = sign;
        else if (dim == 2) o->pbc[iswap][4] = o->pbc[iswap][3] = sign;
      }
      iswap++;
    }
  }
  o->nswap = iswap;
}

/* swap -> dimension it belongs to (needed because non-periodic dims contribute no swaps) */
static int swap_dim(const Orc *o, int iswap) {
  for (int d = 0; d < 3; d++)
    if (o->pbc[iswap][d]) return d;
  return -1;
}

/* offsets a ghost gets from swap iswap: AtomVec::pack_comm (atom_vec.cpp:369-377, box units, with
   the tilt terms of a triclinic box) or pack_border (:813-821: lamda units if triclinic) */
static void swap_shift(const Orc *o, int iswap, int border, double *d) {
  const int *pbc = o->pbc[iswap];
  if (!o->triclinic) {
    d[0] = pbc[0] * o->prd[0];
    d[1] = pbc[1] * o->prd[1];
    d[2] = pbc[2] * o->prd[2];
  } else if (border) {
    d[0] = pbc[0];
    d[1] = pbc[1];
    d[2] = pbc[2];
  } else {
    d[0] = pbc[0] * o->prd[0] + pbc[5] * o->xy + pbc[4] * o->xz;
    d[1] = pbc[1] * o->prd[1] + pbc[3] * o->yz;
    d[2] = pbc[2] * o->prd[2];
  }
}

/* comm_brick.cpp:720-899 CommBrick::borders + atom_vec.cpp:796-830 pack_border /
   :1026-1042 unpack_border (x + pbc*prd, tag, type, mask); self swaps only */
void orc_borders(Orc *o) {
  o->nghost = 0;
  int iswap = 0, nfirst = 0, nlast = 0, lastdim = -1;
  for (iswap = 0; iswap < o->nswap; iswap++) {
    int dim = swap_dim(o, iswap);
    if (dim != lastdim) { nlast = 0; lastdim = dim; }
    int ineed = (o->pbc[iswap][dim] == 1) ? 0 : 1;
    if (ineed % 2 == 0) {
      nfirst = nlast;
      nlast = o->nlocal + o->nghost;
    }
    /* NOTE comm_brick.cpp:752-755: for ineed==0 nfirst = previous nlast which was reset to
       0 at the start of the dim, so owned + all earlier-dim ghosts are scanned. */
    double lo = o->slablo[iswap], hi = o->slabhi[iswap];
    int nsend = 0;
    for (int i = nfirst; i < nlast; i++)
      if (o->x[3 * i + dim] >= lo && o->x[3 * i + dim] <= hi) {
        if (nsend == o->maxsendlist[iswap]) {
          o->maxsendlist[iswap] = nsend + nsend / 2 + 1024;
          o->sendlist[iswap] = xrealloc(o->sendlist[iswap], sizeof(int) * o->maxsendlist[iswap]);
        }
        o->sendlist[iswap][nsend++] = i;
      }
    int first = o->nlocal + o->nghost;
    grow_atoms(o, first + nsend);
    double sh[3];
    swap_shift(o, iswap, 1, sh);
    const double dx = sh[0], dy = sh[1], dz = sh[2];
    for (int k = 0; k < nsend; k++) {
      int j = o->sendlist[iswap][k], g = first + k;
      o->x[3 * g + 0] = o->x[3 * j + 0] + dx;
      o->x[3 * g + 1] = o->x[3 * j + 1] + dy;
      o->x[3 * g + 2] = o->x[3 * j + 2] + dz;
      o->tag[g] = o->tag[j];
      o->type[g] = o->type[j];
      o->mask[g] = o->mask[j];
    }
    o->sendnum[iswap] = nsend;
    o->firstrecv[iswap] = first;
    o->nghost += nsend;
  }
}

/* comm_brick.cpp:485-538 forward_comm + atom_vec.cpp:354-440 pack_comm (self copy) */
void orc_forward_comm(Orc *o) {
  for (int iswap = 0; iswap < o->nswap; iswap++) {
    double sh[3];
    swap_shift(o, iswap, 0, sh);
    const double dx = sh[0], dy = sh[1], dz = sh[2];
    int first = o->firstrecv[iswap];
    for (int k = 0; k < o->sendnum[iswap]; k++) {
      int j = o->sendlist[iswap][k], g = first + k;
      o->x[3 * g + 0] = o->x[3 * j + 0] + dx;
      o->x[3 * g + 1] = o->x[3 * j + 1] + dy;
      o->x[3 * g + 2] = o->x[3 * j + 2] + dz;
    }
  }
}

/* comm_brick.cpp:545-586 reverse_comm + atom_vec.cpp:729 unpack_reverse */
void orc_reverse_comm(Orc *o) {
  for (int iswap = o->nswap - 1; iswap >= 0; iswap--) {
    int first = o->firstrecv[iswap];
    for (int k = 0; k < o->sendnum[iswap]; k++) {
      int j = o->sendlist[iswap][k], g = first + k;
      o->f[3 * j + 0] += o->f[3 * g + 0];
      o->f[3 * j + 1] += o->f[3 * g + 1];
      o->f[3 * j + 2] += o->f[3 * g + 2];
    }
  }
}

/* comm_brick.cpp:952-983 reverse_comm(Pair*) with pair_eam.cpp:1625-1646 (rho) */
static void reverse_comm_rho(Orc *o) {
  for (int iswap = o->nswap - 1; iswap >= 0; iswap--) {
    int first = o->firstrecv[iswap];
    for (int k = 0; k < o->sendnum[iswap]; k++) o->rho[o->sendlist[iswap][k]] += o->rho[first + k];
  }
}

/* comm_brick.cpp:910-941 forward_comm(Pair*) with pair_eam.cpp:1600-1621 (fp) */
static void forward_comm_fp(Orc *o) {
  for (int iswap = 0; iswap < o->nswap; iswap++) {
    int first = o->firstrecv[iswap];
    for (int k = 0; k < o->sendnum[iswap]; k++) o->fp[first + k] = o->fp[o->sendlist[iswap][k]];
  }
}

/* ------------------------------------------------------------------ binning */

/* nbin_standard.cpp:82-214 NBinStandard::setup_bins (style BIN, orthogonal, 3d) */
void orc_setup_bins(Orc *o) {
  double bbox[3], bsubboxlo[3], bsubboxhi[3];
  const double *bboxlo = o->triclinic ? o->boxlo_bound : o->boxlo; /* NBin::bboxlo/bboxhi */
  const double *bboxhi = o->triclinic ? o->boxhi_bound : o->boxhi;
  for (int d = 0; d < 3; d++) {
    bsubboxlo[d] = o->boxlo[d] - o->cutghost;
    bsubboxhi[d] = o->boxhi[d] + o->cutghost;
  }
  if (o->triclinic) {
    /* nbin_standard.cpp:103-111 + Domain::bbox (domain.cpp:2467-2530): bounding box of the lamda
       sub-box (0..1 here) extended by the lamda ghost cutoff */
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = 0.0 - o->cutghost3[d];
      hi[d] = 1.0 + o->cutghost3[d];
      bsubboxlo[d] = BIG;
      bsubboxhi[d] = -BIG;
    }
    for (int c = 0; c < 8; c++) {
      double l[3] = {(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]}, x[3];
      lamda2x_one(o, l, x);
      for (int d = 0; d < 3; d++) {
        bsubboxlo[d] = MIN(bsubboxlo[d], x[d]);
        bsubboxhi[d] = MAX(bsubboxhi[d], x[d]);
      }
    }
  }
  for (int d = 0; d < 3; d++) bbox[d] = bboxhi[d] - bboxlo[d];
  double binsize_optimal = 0.5 * o->cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  double binsizeinv = 1.0 / binsize_optimal;
  o->nbinx = (int)(bbox[0] * binsizeinv);
  o->nbiny = (int)(bbox[1] * binsizeinv);
  o->nbinz = (int)(bbox[2] * binsizeinv);
  if (o->nbinx == 0) o->nbinx = 1;
  if (o->nbiny == 0) o->nbiny = 1;
  if (o->nbinz == 0) o->nbinz = 1;
  o->binsizex = bbox[0] / o->nbinx;
  o->binsizey = bbox[1] / o->nbiny;
  o->binsizez = bbox[2] / o->nbinz;
  o->bininvx = 1.0 / o->binsizex;
  o->bininvy = 1.0 / o->binsizey;
  o->bininvz = 1.0 / o->binsizez;

  int mbinxhi, mbinyhi, mbinzhi;
  double coord;
  coord = bsubboxlo[0] - SMALL * bbox[0];
  o->mbinxlo = (int)((coord - bboxlo[0]) * o->bininvx);
  if (coord < bboxlo[0]) o->mbinxlo = o->mbinxlo - 1;
  coord = bsubboxhi[0] + SMALL * bbox[0];
  mbinxhi = (int)((coord - bboxlo[0]) * o->bininvx);

  coord = bsubboxlo[1] - SMALL * bbox[1];
  o->mbinylo = (int)((coord - bboxlo[1]) * o->bininvy);
  if (coord < bboxlo[1]) o->mbinylo = o->mbinylo - 1;
  coord = bsubboxhi[1] + SMALL * bbox[1];
  mbinyhi = (int)((coord - bboxlo[1]) * o->bininvy);

  coord = bsubboxlo[2] - SMALL * bbox[2];
  o->mbinzlo = (int)((coord - bboxlo[2]) * o->bininvz);
  if (coord < bboxlo[2]) o->mbinzlo = o->mbinzlo - 1;
  coord = bsubboxhi[2] + SMALL * bbox[2];
  mbinzhi = (int)((coord - bboxlo[2]) * o->bininvz);

  o->mbinxlo -= 1; mbinxhi += 1; o->mbinx = mbinxhi - o->mbinxlo + 1;
  o->mbinylo -= 1; mbinyhi += 1; o->mbiny = mbinyhi - o->mbinylo + 1;
  o->mbinzlo -= 1; mbinzhi += 1; o->mbinz = mbinzhi - o->mbinzlo + 1;
  int64_t bbin = (int64_t)o->mbinx * o->mbiny * o->mbinz + 1;
  if (bbin > 2147483647) die("too many neighbor bins");
  o->mbins = (int)bbin;
  if (o->mbins > o->maxbin) {
    o->maxbin = o->mbins;
    o->binhead = xrealloc(o->binhead, sizeof(int) * o->maxbin);
  }
}

/* nbin.cpp:141-173 NBin::coord2bin */
static int coord2bin(const Orc *o, const double *x) {
  int ix, iy, iz;
  const double *bboxlo = o->triclinic ? o->boxlo_bound : o->boxlo;
  const double *bboxhi = o->triclinic ? o->boxhi_bound : o->boxhi;
  if (!isfinite(x[0]) || !isfinite(x[1]) || !isfinite(x[2])) die("non-numeric positions");
  if (x[0] >= bboxhi[0])
    ix = (int)((x[0] - bboxhi[0]) * o->bininvx) + o->nbinx;
  else if (x[0] >= bboxlo[0]) {
    ix = (int)((x[0] - bboxlo[0]) * o->bininvx);
    ix = MIN(ix, o->nbinx - 1);
  } else
    ix = (int)((x[0] - bboxlo[0]) * o->bininvx) - 1;
  if (x[1] >= bboxhi[1])
    iy = (int)((x[1] - bboxhi[1]) * o->bininvy) + o->nbiny;
  else if (x[1] >= bboxlo[1]) {
    iy = (int)((x[1] - bboxlo[1]) * o->bininvy);
    iy = MIN(iy, o->nbiny - 1);
  } else
    iy = (int)((x[1] - bboxlo[1]) * o->bininvy) - 1;
  if (x[2] >= bboxhi[2])
    iz = (int)((x[2] - bboxhi[2]) * o->bininvz) + o->nbinz;
  else if (x[2] >= bboxlo[2]) {
    iz = (int)((x[2] - bboxlo[2]) * o->bininvz);
    iz = MIN(iz, o->nbinz - 1);
  } else
    iz = (int)((x[2] - bboxlo[2]) * o->bininvz) - 1;
  return (iz - o->mbinzlo) * o->mbiny * o->mbinx + (iy - o->mbinylo) * o->mbinx + (ix - o->mbinxlo);
}

/* nbin_standard.cpp:220-260 NBinStandard::bin_atoms */
static void bin_atoms(Orc *o) {
  for (int i = 0; i < o->mbins; i++) o->binhead[i] = -1;
  int nall = o->nlocal + o->nghost;
  for (int i = nall - 1; i >= 0; i--) {
    int ibin = coord2bin(o, &o->x[3 * i]);
    o->atom2bin[i] = ibin;
    o->bins[i] = o->binhead[ibin];
    o->binhead[ibin] = i;
  }
}

/* nstencil.cpp:374-391 NStencil::bin_distance */
static double bin_distance(const Orc *o, int i, int j, int k) {
  double delx, dely, delz;
  if (i > 0) delx = (i - 1) * o->binsizex;
  else if (i == 0) delx = 0.0;
  else delx = (i + 1) * o->binsizex;
  if (j > 0) dely = (j - 1) * o->binsizey;
  else if (j == 0) dely = 0.0;
  else dely = (j + 1) * o->binsizey;
  if (k > 0) delz = (k - 1) * o->binsizez;
  else if (k == 0) delz = 0.0;
  else delz = (k + 1) * o->binsizez;
  return delx * delx + dely * dely + delz * delz;
}

/* nstencil.cpp:203-237 create_setup + nstencil_bin.cpp:28-67 NStencilBin<HALF=1,3D=1,TRI=0>::create */
void orc_create_stencil(Orc *o) {
  int sx = (int)(o->cutneighmax * o->bininvx);
  if (sx * o->binsizex < o->cutneighmax) sx++;
  int sy = (int)(o->cutneighmax * o->bininvy);
  if (sy * o->binsizey < o->cutneighmax) sy++;
  int sz = (int)(o->cutneighmax * o->bininvz);
  if (sz * o->binsizez < o->cutneighmax) sz++;
  int smax = (2 * sx + 1) * (2 * sy + 1) * (2 * sz + 1);
  o->stencil = xrealloc(o->stencil, sizeof(int) * smax);
  int n = 0;
  /* half/newton/orthogonal: central bin first, then the upper half.  half/newton/triclinic
     (NStencilBin<1,1,1>) and half/newtoff (NStencilBin<0,1,0>, the full stencil): every bin, in
     loop order, no separate central bin (nstencil_bin.cpp:36-62) */
  const int full = o->triclinic || !o->newton_pair;
  if (!full) o->stencil[n++] = 0;
  for (int k = full ? -sz : 0; k <= sz; k++)
    for (int j = -sy; j <= sy; j++)
      for (int i = -sx; i <= sx; i++) {
        if (!full && k <= 0 && j <= 0 && (j != 0 || i <= 0)) continue;
        if (bin_distance(o, i, j, k) < o->cutneighmaxsq)
          o->stencil[n++] = k * o->mbiny * o->mbinx + j * o->mbinx + i;
      }
  o->nstencil = n;
}

/* ------------------------------------------------------------------ pair list */

/* npair_bin.cpp:52-253 NPairBin<HALF=1,NEWTON=1,TRI=0,SIZE=0,ATOMONLY=1>::build */
static void npair_build(Orc *o) {
  int nlocal = o->nlocal, n1 = o->ntypes + 1;
  const int newton = o->newton_pair, tri = o->triclinic;
  const double delta = 0.01 * o->angstrom; /* npair_bin.cpp:59 */
  int64_t total = 0;
  for (int i = 0; i < nlocal; i++) {
    int itype = o->type[i];
    double xtmp = o->x[3 * i], ytmp = o->x[3 * i + 1], ztmp = o->x[3 * i + 2];
    int ibin = o->atom2bin[i];
    o->firstneigh[i] = total;
    int n = 0;
    const int itag = o->tag[i];
    for (int k = 0; k < o->nstencil; k++) {
      int bin_start = o->binhead[ibin + o->stencil[k]];
      if (newton && !tri && k == 0) bin_start = o->bins[i];
      for (int j = bin_start; j >= 0; j = o->bins[j]) {
        if (!newton) {
          /* half list, newton off (npair_bin.cpp:126-131): own/own pairs once, own/ghost pairs
             on both procs */
          if (j <= i) continue;
        } else if (tri) {
          /* npair_bin.cpp:133-155 */
          if (j <= i) continue;
          if (j >= nlocal) {
            int jtag = o->tag[j];
            if (itag > jtag) {
              if ((itag + jtag) % 2 == 0) continue;
            } else if (itag < jtag) {
              if ((itag + jtag) % 2 == 1) continue;
            } else {
              if (fabs(o->x[3 * j + 2] - ztmp) > delta) {
                if (o->x[3 * j + 2] < ztmp) continue;
              } else if (fabs(o->x[3 * j + 1] - ytmp) > delta) {
                if (o->x[3 * j + 1] < ytmp) continue;
              } else {
                if (o->x[3 * j] < xtmp) continue;
              }
            }
          }
        } else if (k == 0) {
          if (j >= nlocal) {
            if (o->x[3 * j + 2] < ztmp) continue;
            if (o->x[3 * j + 2] == ztmp) {
              if (o->x[3 * j + 1] < ytmp) continue;
              if (o->x[3 * j + 1] == ytmp && o->x[3 * j] < xtmp) continue;
            }
          }
        }
        int jtype = o->type[j];
        /* NPair::exclusion, npair.cpp:244-254 (type pairs, group pairs; atomic systems have no
           molecules) */
        if (o->ex_type && o->ex_type[itype * n1 + jtype]) continue;
        if (o->nex_group) {
          int ex = 0;
          for (int m = 0; m < o->nex_group; m++) {
            if ((o->mask[i] & o->ex1_bit[m]) && (o->mask[j] & o->ex2_bit[m])) ex = 1;
            if ((o->mask[i] & o->ex2_bit[m]) && (o->mask[j] & o->ex1_bit[m])) ex = 1;
          }
          if (ex) continue;
        }
        double delx = xtmp - o->x[3 * j];
        double dely = ytmp - o->x[3 * j + 1];
        double delz = ztmp - o->x[3 * j + 2];
        double rsq = delx * delx + dely * dely + delz * delz;
        if (rsq <= o->cutneighsq[itype * n1 + jtype]) {
          if (total + n >= o->neighcap) {
            o->neighcap = o->neighcap + o->neighcap / 2 + 1000000;
            o->neigh = xrealloc(o->neigh, sizeof(int) * o->neighcap);
          }
          o->neigh[total + n++] = j;
        }
      }
    }
    o->numneigh[i] = n;
    total += n;
  }
  o->inum = nlocal;
  o->nneigh = total;
}

/* neighbor.cpp:2498-2551 Neighbor::build */
void orc_neighbor_build(Orc *o) {
  o->ago = 0;
  o->ncalls++;
  if (o->dist_check)
    for (int i = 0; i < 3 * o->nlocal; i++) o->xhold[i] = o->x[i];
  bin_atoms(o);
  npair_build(o);
}

/* neighbor.cpp:2408-2424 decide + :2438-2490 check_distance (no box change) */
int orc_decide(Orc *o) {
  o->ago++;
  if (o->ago >= o->delay && o->ago % o->every == 0) {
    if (o->build_once) return 0;
    if (o->dist_check == 0) return 1;
    int flag = 0;
    for (int i = 0; i < o->nlocal; i++) {
      double delx = o->x[3 * i] - o->xhold[3 * i];
      double dely = o->x[3 * i + 1] - o->xhold[3 * i + 1];
      double delz = o->x[3 * i + 2] - o->xhold[3 * i + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq > o->triggersq) { flag = 1; break; }
    }
    if (flag && o->ago == MAX(o->every, o->delay)) o->ndanger++;
    return flag;
  }
  return 0;
}

/* ------------------------------------------------------------------ force */

/* verlet.cpp:376-421 force_clear (newton on: owned + ghost) */
void orc_force_clear(Orc *o) {
  memset(o->f, 0, sizeof(double) * 3 * (o->nlocal + o->nghost));
}

/* pair.cpp:1809-1825 virial_fdotr_compute */
static void virial_fdotr(Orc *o) {
  int nall = o->nlocal + o->nghost;
  const double *x = o->x, *f = o->f;
  for (int i = 0; i < nall; i++) {
    o->virial[0] += f[3 * i + 0] * x[3 * i + 0];
    o->virial[1] += f[3 * i + 1] * x[3 * i + 1];
    o->virial[2] += f[3 * i + 2] * x[3 * i + 2];
    o->virial[3] += f[3 * i + 1] * x[3 * i + 0];
    o->virial[4] += f[3 * i + 2] * x[3 * i + 0];
    o->virial[5] += f[3 * i + 2] * x[3 * i + 1];
  }
}

/* Pair::ev_tally, pair.cpp:1087-1150, global tallies with newton_pair off: energy and the pairwise
   virial del (x) del * fpair go half to each LOCAL atom of the pair (with newton off the virial
   is tallied pair by pair, integrate.cpp:81-82 VIRIAL_PAIR, not by virial_fdotr_compute) */
static void ev_tally_newtoff(Orc *o, int i, int j, int eflag, int vflag, double evdwl, double fpair,
                             double delx, double dely, double delz) {
  const int nlocal = o->nlocal;
  if (eflag) {
    double evdwlhalf = 0.5 * evdwl;
    if (i < nlocal) o->eng_vdwl += evdwlhalf;
    if (j < nlocal) o->eng_vdwl += evdwlhalf;
  }
  if (vflag) {
    double v[6];
    v[0] = delx * delx * fpair;
    v[1] = dely * dely * fpair;
    v[2] = delz * delz * fpair;
    v[3] = delx * dely * fpair;
    v[4] = delx * delz * fpair;
    v[5] = dely * delz * fpair;
    if (i < nlocal)
      for (int k = 0; k < 6; k++) o->virial[k] += 0.5 * v[k];
    if (j < nlocal)
      for (int k = 0; k < 6; k++) o->virial[k] += 0.5 * v[k];
  }
}

/* pair_lj_cut.cpp:71-141 PairLJCut::compute; ev_tally (pair.cpp:1087-1182) reduced to the
   global-energy branch with newton_pair on: eng_vdwl += evdwl */
static void lj_compute(Orc *o, int eflag, int vflag) {
  int n1 = o->ntypes + 1;
  const int newton_pair = o->newton_pair, nlocal = o->nlocal;
  double *x = o->x, *f = o->f;
  for (int i = 0; i < o->inum; i++) {
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    int itype = o->type[i];
    const int *jlist = &o->neigh[o->firstneigh[i]];
    int jnum = o->numneigh[i];
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj];
      double factor_lj = o->special_lj[(j >> 30) & 3];
      j &= 0x1FFFFFFF;
      double delx = xtmp - x[3 * j];
      double dely = ytmp - x[3 * j + 1];
      double delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      int jtype = o->type[j];
      if (rsq < o->cutsq[itype * n1 + jtype]) {
        double r2inv = 1.0 / rsq;
        double r6inv = r2inv * r2inv * r2inv;
        double forcelj = r6inv * (o->lj1[itype * n1 + jtype] * r6inv - o->lj2[itype * n1 + jtype]);
        double fpair = factor_lj * forcelj * r2inv;
        f[3 * i + 0] += delx * fpair;
        f[3 * i + 1] += dely * fpair;
        f[3 * i + 2] += delz * fpair;
        if (newton_pair || j < nlocal) {
          f[3 * j + 0] -= delx * fpair;
          f[3 * j + 1] -= dely * fpair;
          f[3 * j + 2] -= delz * fpair;
        }
        double evdwl = 0.0;
        if (eflag) {
          evdwl = r6inv * (o->lj3[itype * n1 + jtype] * r6inv - o->lj4[itype * n1 + jtype]) -
                  o->offset[itype * n1 + jtype];
          evdwl *= factor_lj;
          if (newton_pair) o->eng_vdwl += evdwl;
        }
        if (!newton_pair && (eflag || vflag))
          ev_tally_newtoff(o, i, j, eflag, vflag, evdwl, fpair, delx, dely, delz);
      }
    }
  }
  if (vflag && newton_pair) virial_fdotr(o);
}

/* pair_eam.cpp:124-327 PairEAM::compute + :338-366 compute_embedding<0> + pair_eam.h:146-169 */
static void eam_compute(Orc *o, int eflag, int vflag) {
  int n1 = o->ntypes + 1, nlocal = o->nlocal, nall = o->nlocal + o->nghost;
  int nr = o->nr, nrho = o->nrho;
  double *x = o->x, *f = o->f, *rho = o->rho, *fp = o->fp;
  const double rdr = o->rdr, rdrho = o->rdrho, cutforcesq = o->cutforcesq;
#define RHOR(t, m) (&o->rhor_spline[((size_t)(t) * (nr + 1) + (m)) * 7])
#define Z2R(t, m) (&o->z2r_spline[((size_t)(t) * (nr + 1) + (m)) * 7])
#define FRHO(t, m) (&o->frho_spline[((size_t)(t) * (nrho + 1) + (m)) * 7])
  const int newton_pair = o->newton_pair;
  for (int i = 0; i < (newton_pair ? nall : nlocal); i++) rho[i] = 0.0; /* pair_eam.cpp:149-153 */
  for (int i = 0; i < o->inum; i++) {
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    int itype = o->type[i];
    const int *jlist = &o->neigh[o->firstneigh[i]];
    int jnum = o->numneigh[i];
    double rhotmp = rho[i];
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj] & 0x1FFFFFFF;
      double delx = xtmp - x[3 * j];
      double dely = ytmp - x[3 * j + 1];
      double delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq < cutforcesq) {
        int jtype = o->type[j];
        double p = sqrt(rsq) * rdr + 1.0;
        int m = (int)p;
        m = MIN(m, nr - 1);
        p -= m;
        p = MIN(p, 1.0);
        const double *coeff = RHOR(o->type2rhor[jtype * n1 + itype], m);
        rhotmp += ((coeff[3] * p + coeff[4]) * p + coeff[5]) * p + coeff[6];
        if (newton_pair || j < nlocal) {
          coeff = RHOR(o->type2rhor[itype * n1 + jtype], m);
          rho[j] += ((coeff[3] * p + coeff[4]) * p + coeff[5]) * p + coeff[6];
        }
      }
    }
    rho[i] = rhotmp;
  }
  if (newton_pair) reverse_comm_rho(o); /* pair_eam.cpp:215 */

  int beyond_rhomax = 0;
  for (int i = 0; i < o->inum; i++) {
    double p = rho[i] * rdrho + 1.0;
    int m = (int)p;
    m = MAX(1, MIN(m, nrho - 1));
    p -= m;
    p = MIN(p, 1.0);
    const double *coeff = FRHO(o->type2frho[o->type[i]], m);
    fp[i] = (coeff[0] * p + coeff[1]) * p + coeff[2];
    if (eflag) {
      double phi = ((coeff[3] * p + coeff[4]) * p + coeff[5]) * p + coeff[6];
      if (rho[i] > o->rhomax) {
        phi += fp[i] * (rho[i] - o->rhomax);
        beyond_rhomax = 1;
      }
      phi *= o->scale[o->type[i] * n1 + o->type[i]];
      o->eng_vdwl += phi;
    }
  }
  forward_comm_fp(o);

  for (int i = 0; i < o->inum; i++) {
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    int itype = o->type[i];
    const int *jlist = &o->neigh[o->firstneigh[i]];
    int jnum = o->numneigh[i];
    int nforce = 0;
    double fxtmp = f[3 * i], fytmp = f[3 * i + 1], fztmp = f[3 * i + 2], fptmp = fp[i];
    for (int jj = 0; jj < jnum; jj++) {
      int j = jlist[jj] & 0x1FFFFFFF;
      double delx = xtmp - x[3 * j];
      double dely = ytmp - x[3 * j + 1];
      double delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq < cutforcesq) {
        ++nforce;
        int jtype = o->type[j];
        double r = sqrt(rsq);
        double p = r * rdr + 1.0;
        int m = (int)p;
        m = MIN(m, nr - 1);
        p -= m;
        p = MIN(p, 1.0);
        const double *coeff = RHOR(o->type2rhor[itype * n1 + jtype], m);
        double rhoip = (coeff[0] * p + coeff[1]) * p + coeff[2];
        coeff = RHOR(o->type2rhor[jtype * n1 + itype], m);
        double rhojp = (coeff[0] * p + coeff[1]) * p + coeff[2];
        coeff = Z2R(o->type2z2r[itype * n1 + jtype], m);
        double z2p = (coeff[0] * p + coeff[1]) * p + coeff[2];
        double z2 = ((coeff[3] * p + coeff[4]) * p + coeff[5]) * p + coeff[6];
        double recip = 1.0 / r;
        double phi = z2 * recip;
        double phip = z2p * recip - phi * recip;
        double psip = fptmp * rhojp + fp[j] * rhoip + phip;
        double fpair = -o->scale[itype * n1 + jtype] * psip * recip;
        fxtmp += delx * fpair;
        fytmp += dely * fpair;
        fztmp += delz * fpair;
        if (newton_pair || j < nlocal) {
          f[3 * j + 0] -= delx * fpair;
          f[3 * j + 1] -= dely * fpair;
          f[3 * j + 2] -= delz * fpair;
        }
        if (newton_pair) {
          if (eflag) o->eng_vdwl += o->scale[itype * n1 + jtype] * phi;
        } else if (eflag || vflag)
          ev_tally_newtoff(o, i, j, eflag, vflag, eflag ? o->scale[itype * n1 + jtype] * phi : 0.0, fpair, delx,
                           dely, delz);
      }
    }
    o->numforce[i] = nforce;
    f[3 * i] = fxtmp;
    f[3 * i + 1] = fytmp;
    f[3 * i + 2] = fztmp;
  }
  if (eflag && beyond_rhomax) o->exceeded_rhomax = 1;
  if (vflag && newton_pair) virial_fdotr(o);
#undef RHOR
#undef Z2R
#undef FRHO
}

/* pair.cpp:966-1058 ev_setup (global tallies only) then the style's compute() */
void orc_pair_compute(Orc *o, int eflag, int vflag) {
  o->eng_vdwl = 0.0;
  for (int k = 0; k < 6; k++) o->virial[k] = 0.0;
  if (o->pair_style == 1) lj_compute(o, eflag, vflag);
  else if (o->pair_style == 2) eam_compute(o, eflag, vflag);
  else die("no pair style");
}

/* Per-atom energy and virial of the pair style for the positions, list and (eam) rho/fp in force:
 * Pair::ev_tally with eflag_atom / vflag_atom and newton_pair on (pair.cpp:1087-1182: half of the
 * pair energy and of del (x) del * fpair to each atom of the pair, ghosts included), the embedding
 * energy of PairEAM::compute (pair_eam.cpp:219-231: eatom[i] += phi), then the reverse
 * communication that ComputePEAtom / ComputeStressAtom apply (compute_pe_atom.cpp:141-145,
 * compute_stress_atom.cpp:349-356: ghost shares are summed into their owners, last swap first).
 * eatom[nlocal], vatom[nlocal][6] in the order xx,yy,zz,xy,xz,yz; call after a compute with eflag. */
void orc_pair_peratom(Orc *o, double *eatom_out, double *vatom_out) {
  if (!o->newton_pair) die("per-atom tallies are restated for newton_pair on only");
  int n1 = o->ntypes + 1, nall = o->nlocal + o->nghost;
  double *x = o->x;
  double *t = (double *)calloc((size_t)MAX(nall, 1) * 7, sizeof(double)); /* [nall][7]: e, v[6] */
  for (int i = 0; i < o->inum; i++) {
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    int itype = o->type[i];
    const int *jlist = &o->neigh[o->firstneigh[i]];
    for (int jj = 0; jj < o->numneigh[i]; jj++) {
      int j = jlist[jj] & 0x1FFFFFFF;
      double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      int jtype = o->type[j];
      double evdwl, fpair;
      if (o->pair_style == 1) {
        int tij = itype * n1 + jtype;
        if (!(rsq < o->cutsq[tij])) continue;
        double r2inv = 1.0 / rsq;
        double r6inv = r2inv * r2inv * r2inv;
        double forcelj = r6inv * (o->lj1[tij] * r6inv - o->lj2[tij]);
        fpair = forcelj * r2inv;
        evdwl = r6inv * (o->lj3[tij] * r6inv - o->lj4[tij]) - o->offset[tij];
      } else {
        if (!(rsq < o->cutforcesq)) continue;
        int nr = o->nr;
        double r = sqrt(rsq);
        double p = r * o->rdr + 1.0;
        int m = (int)p;
        m = MIN(m, nr - 1);
        p -= m;
        p = MIN(p, 1.0);
        const double *c = &o->rhor_spline[((size_t)o->type2rhor[itype * n1 + jtype] * (nr + 1) + m) * 7];
        double rhoip = (c[0] * p + c[1]) * p + c[2];
        c = &o->rhor_spline[((size_t)o->type2rhor[jtype * n1 + itype] * (nr + 1) + m) * 7];
        double rhojp = (c[0] * p + c[1]) * p + c[2];
        c = &o->z2r_spline[((size_t)o->type2z2r[itype * n1 + jtype] * (nr + 1) + m) * 7];
        double z2p = (c[0] * p + c[1]) * p + c[2];
        double z2 = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
        double recip = 1.0 / r;
        double phi = z2 * recip;
        double phip = z2p * recip - phi * recip;
        double psip = o->fp[i] * rhojp + o->fp[j] * rhoip + phip;
        fpair = -o->scale[itype * n1 + jtype] * psip * recip;
        evdwl = o->scale[itype * n1 + jtype] * phi;
      }
      double v[6] = {delx * delx * fpair, dely * dely * fpair, delz * delz * fpair,
                     delx * dely * fpair, delx * delz * fpair, dely * delz * fpair};
      t[7 * i] += 0.5 * evdwl;
      t[7 * j] += 0.5 * evdwl;
      for (int k = 0; k < 6; k++) {
        t[7 * i + 1 + k] += 0.5 * v[k];
        t[7 * j + 1 + k] += 0.5 * v[k];
      }
    }
  }
  if (o->pair_style == 2)
    for (int i = 0; i < o->inum; i++) {
      int nrho = o->nrho;
      double p = o->rho[i] * o->rdrho + 1.0;
      int m = (int)p;
      m = MAX(1, MIN(m, nrho - 1));
      p -= m;
      p = MIN(p, 1.0);
      const double *c = &o->frho_spline[((size_t)o->type2frho[o->type[i]] * (nrho + 1) + m) * 7];
      double phi = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
      if (o->rho[i] > o->rhomax) phi += o->fp[i] * (o->rho[i] - o->rhomax);
      t[7 * i] += phi * o->scale[o->type[i] * n1 + o->type[i]];
    }
  for (int iswap = o->nswap - 1; iswap >= 0; iswap--) {
    int first = o->firstrecv[iswap];
    for (int k = 0; k < o->sendnum[iswap]; k++)
      for (int q = 0; q < 7; q++) t[7 * o->sendlist[iswap][k] + q] += t[7 * (first + k) + q];
  }
  for (int i = 0; i < o->nlocal; i++) {
    if (eatom_out) eatom_out[i] = t[7 * i];
    if (vatom_out)
      for (int k = 0; k < 6; k++) vatom_out[6 * i + k] = t[7 * i + 1 + k];
  }
  free(t);
}

/* ------------------------------------------------------------------ fix nve */

/* fix_nve.cpp:68-108 */
void orc_initial_integrate(Orc *o) {
  for (int i = 0; i < o->nlocal; i++)
    if (o->mask[i] & o->groupbit) {
      double dtfm = o->dtf / o->mass[o->type[i]];
      o->v[3 * i + 0] += dtfm * o->f[3 * i + 0];
      o->v[3 * i + 1] += dtfm * o->f[3 * i + 1];
      o->v[3 * i + 2] += dtfm * o->f[3 * i + 2];
      o->x[3 * i + 0] += o->dtv * o->v[3 * i + 0];
      o->x[3 * i + 1] += o->dtv * o->v[3 * i + 1];
      o->x[3 * i + 2] += o->dtv * o->v[3 * i + 2];
    }
}

/* fix_nve.cpp:112-145 */
void orc_final_integrate(Orc *o) {
  for (int i = 0; i < o->nlocal; i++)
    if (o->mask[i] & o->groupbit) {
      double dtfm = o->dtf / o->mass[o->type[i]];
      o->v[3 * i + 0] += dtfm * o->f[3 * i + 0];
      o->v[3 * i + 1] += dtfm * o->f[3 * i + 1];
      o->v[3 * i + 2] += dtfm * o->f[3 * i + 2];
    }
}

/* ------------------------------------------------------------------ timestep driver */

/* verlet.cpp:93-162 Verlet::setup (atom->sort() is skipped: it only permutes atoms) */
void orc_setup(Orc *o, int eflag, int vflag) {
  neighbor_init(o);
  if (o->triclinic) x2lamda(o, o->nlocal); /* verlet.cpp:111 */
  orc_pbc(o);
  comm_setup(o);
  orc_setup_bins(o);
  orc_create_stencil(o);
  /* comm->exchange(): single proc, nothing migrates after pbc() */
  orc_borders(o);
  if (o->triclinic) lamda2x(o, o->nlocal + o->nghost); /* verlet.cpp:121 */
  orc_neighbor_build(o);
  o->ncalls = 0;
  o->ndanger = 0;
  orc_force_clear(o);
  orc_pair_compute(o, eflag, vflag);
  if (o->newton_pair) orc_reverse_comm(o); /* verlet.cpp:147: if (force->newton) */
}

/* one iteration of verlet.cpp:229-360 Verlet::run; returns 1 if the list was rebuilt */
int orc_step(Orc *o, int eflag, int vflag) {
  orc_initial_integrate(o);
  int nflag = orc_decide(o);
  if (nflag == 0) {
    orc_forward_comm(o);
  } else {
    if (o->triclinic) x2lamda(o, o->nlocal); /* verlet.cpp:293 */
    orc_pbc(o);
    orc_borders(o);
    if (o->triclinic) lamda2x(o, o->nlocal + o->nghost); /* verlet.cpp:313 */
    orc_neighbor_build(o);
  }
  orc_force_clear(o);
  orc_pair_compute(o, eflag, vflag);
  if (o->newton_pair) orc_reverse_comm(o); /* verlet.cpp:345: if (force->newton) */
  orc_final_integrate(o);
  return nflag;
}

/* thermo_every > 0: energy/virial tallied on steps that are multiples of it and on the
   last step (integrate.cpp:106-151 ev_set follows output->next).  thermo_out gets 8 doubles
   per thermo step: step, ke_sum(=sum m v^2), eng_vdwl, virial[0..2] trace parts, ... */
void orc_ke_sum(const Orc *o, double *out);

int orc_run(Orc *o, int nsteps, int step0, int thermo_every, double *thermo_out, int max_out) {
  int nout = 0;
  for (int s = 1; s <= nsteps; s++) {
    int step = step0 + s;
    int ev = (thermo_every > 0 && (step % thermo_every == 0)) || s == nsteps;
    orc_step(o, ev, ev);
    if (ev && thermo_out && nout < max_out) {
      double *t = &thermo_out[10 * nout++];
      t[0] = step;
      orc_ke_sum(o, &t[1]);
      t[2] = o->eng_vdwl;
      for (int k = 0; k < 6; k++) t[3 + k] = o->virial[k];
      t[9] = 0.0;
    }
  }
  return nout;
}

/* compute_temp.cpp:73-97: t = sum (vx^2+vy^2+vz^2)*mass[type] over owned atoms in group all */
void orc_ke_sum(const Orc *o, double *out) {
  double t = 0.0;
  for (int i = 0; i < o->nlocal; i++) {
    const double *v = &o->v[3 * i];
    t += (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * o->mass[o->type[i]];
  }
  *out = t;
}

/* ------------------------------------------------------------------ getters */

int orc_nlocal(const Orc *o) { return o->nlocal; }
int orc_nghost(const Orc *o) { return o->nghost; }
int64_t orc_nneigh(const Orc *o) { return o->nneigh; }
int orc_ncalls(const Orc *o) { return o->ncalls; }
int orc_ndanger(const Orc *o) { return o->ndanger; }
int orc_ago(const Orc *o) { return o->ago; }
double orc_eng_vdwl(const Orc *o) { return o->eng_vdwl; }
void orc_get_virial(const Orc *o, double *v) { memcpy(v, o->virial, sizeof(double) * 6); }
void orc_get_bins(const Orc *o, int *out) {
  out[0] = o->nbinx; out[1] = o->nbiny; out[2] = o->nbinz;
  out[3] = o->mbinx; out[4] = o->mbiny; out[5] = o->mbinz;
  out[6] = o->mbinxlo; out[7] = o->mbinylo; out[8] = o->mbinzlo;
  out[9] = o->nstencil;
}
void orc_get_stencil(const Orc *o, int *out) { memcpy(out, o->stencil, sizeof(int) * o->nstencil); }

/* which: 0 x, 1 v, 2 f ; n = number of atoms to copy (owned first, then ghosts) */
void orc_get_vec(const Orc *o, int which, int n, double *out) {
  const double *src = which == 0 ? o->x : which == 1 ? o->v : o->f;
  memcpy(out, src, sizeof(double) * 3 * n);
}
/* which: 0 type, 1 tag, 2 mask, 3 image, 4 numneigh */
void orc_get_ivec(const Orc *o, int which, int n, int *out) {
  const int *src = which == 0 ? o->type : which == 1 ? o->tag : which == 2 ? o->mask
                 : which == 3 ? o->image : o->numneigh;
  memcpy(out, src, sizeof(int) * n);
}
void orc_get_rho_fp(const Orc *o, int n, double *rho, double *fp) {
  memcpy(rho, o->rho, sizeof(double) * n);
  memcpy(fp, o->fp, sizeof(double) * n);
}
/* neighbour pairs as local indices (i owned, j owned-or-ghost), list order */
void orc_get_pairs(const Orc *o, int *pi, int *pj) {
  int64_t k = 0;
  for (int i = 0; i < o->inum; i++)
    for (int jj = 0; jj < o->numneigh[i]; jj++) {
      pi[k] = i;
      pj[k++] = o->neigh[o->firstneigh[i] + jj];
    }
}

/* ------------------------------------------------------------------ velocity create */

/* random_park.cpp:41-48 uniform, :96-130 reset(ibase, coord): Jenkins one-at-a-time hash
   over the bytes of the seed and the 3 coordinates (char is signed on the reference
   platform), 5 warm-up draws.  Used by `velocity ... loop geom` (velocity.cpp:327-352). */
#define IA 16807
#define IM 2147483647
#define AM (1.0 / IM)
#define IQ 127773
#define IR 2836

static double park_uniform(int *seed) {
  int k = *seed / IQ;
  *seed = IA * (*seed - k * IQ) - IR * k;
  if (*seed < 0) *seed += IM;
  return AM * *seed;
}

void orc_velocity_loop_geom(int n, int seed, const double *x, const double *mass_per_atom,
                            double *v) {
  for (int i = 0; i < n; i++) {
    int ibase = seed;
    const signed char *str = (const signed char *)&ibase;
    unsigned int hash = 0;
    for (int b = 0; b < (int)sizeof(int); b++) {
      hash += str[b];
      hash += (hash << 10);
      hash ^= (hash >> 6);
    }
    str = (const signed char *)&x[3 * i];
    for (int b = 0; b < (int)(3 * sizeof(double)); b++) {
      hash += str[b];
      hash += (hash << 10);
      hash ^= (hash >> 6);
    }
    hash += (hash << 3);
    hash ^= (hash >> 11);
    hash += (hash << 15);
    int s = hash & 0x7ffffff;
    if (!s) s = 1;
    for (int w = 0; w < 5; w++) park_uniform(&s);
    double vx = park_uniform(&s) - 0.5;
    double vy = park_uniform(&s) - 0.5;
    double vz = park_uniform(&s) - 0.5;
    double factor = 1.0 / sqrt(mass_per_atom[i]);
    v[3 * i] = vx * factor;
    v[3 * i + 1] = vy * factor;
    v[3 * i + 2] = vz * factor;
  }
}
