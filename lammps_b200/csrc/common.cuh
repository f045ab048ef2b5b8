// common.cuh -- shared types and device helpers of the B200 MD engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_md.h"

#define NEIGHMASK 0x1FFFFFFF /* lmptype.h:63 */
#define IMGMASK 1023         /* lmptype.h:122-125, LAMMPS_SMALLBIG */
#define IMGBITS 10
#define IMG2BITS 20
#define B200_SMALL 1.0e-6 /* nbin_standard.cpp:27 */
#define MAXROWS 64        /* stencil rows (dz,dy): 13 of the half stencil for sx=sy=sz=2, 25..49 of the
                             full stencil a triclinic box uses (nstencil_bin.cpp:36-62) */
#define NDIR 27

// Geometry of the box, the sub-domain, the bins and the ghost slabs; passed to kernels by value.
// Everything that decides where an atom belongs -- periodic wrap, sub-domain ownership, ghost
// slabs, the shift a ghost gets when it is created -- lives in the "comm frame": box coordinates
// for an orthogonal box, lamda (0-1) coordinates for a triclinic one (comm_brick.cpp:177-237,
// verlet.cpp:293-313: x2lamda before pbc/exchange/borders, lamda2x after).  Bins are always in box
// coordinates, over the box (orthogonal) or its bounding box (triclinic, nbin_standard.cpp:86-112).
struct Geom {
  double boxlo[3], boxhi[3], prd[3];  // comm frame: the box, or (0,1,1) in lamda coordinates
  double sublo[3], subhi[3];          // comm frame
  int periodic[3];
  // bins: nbin_standard.cpp:82-214
  double binlo[3], binhi[3];  // NBin::bboxlo / bboxhi: boxlo/boxhi, or boxlo_bound/boxhi_bound
  int nbin[3];
  double bininv[3];
  int mbinlo[3], mbin[3];
  int mbins;
  // ghost slabs: comm_brick.cpp:389-420 (slabhi of the "send left" swap, slablo of "send right")
  double slab_left_hi[3], slab_right_lo[3];  // comm frame
  int send_left[3], send_right[3];
  // per direction (dz+1)*9+(dy+1)*3+(dx+1): periodic shift added to a ghost's position when it is
  // created at a rebuild (comm frame: pbc * prd, or pbc in lamda units, atom_vec.cpp:796-830)
  double shift[NDIR][3];
  // ... and on every forward halo (box coordinates, AtomVec::pack_comm atom_vec.cpp:354-440): the
  // reference moves a ghost through up to three swaps, each adding its own offset with its own
  // rounding: x swap (pbc[0] xprd), y swap (pbc[5] xy, pbc[1] yprd), z swap (pbc[4] xz, pbc[3] yz,
  // pbc[2] zprd).  fshift = the diagonal terms, tilt = {xy, xz, yz} terms; zero in an orthogonal box.
  double fshift[NDIR][3];
  double tilt[NDIR][3];
  // triclinic box (domain.cpp:263-290): h = {xprd, yprd, zprd, yz, xz, xy}, h_inv, lamda origin
  int tri;
  double h[6], h_inv[6], origin[3];
};

// neigh_modify exclude group g1 g2 (NPair::exclusion, npair.cpp:249-254): a pair is never stored when
// one atom is in the first group and the other in the second
#define MAXEXGROUP 8
struct ExGroups {
  int n;
  int bit1[MAXEXGROUP], bit2[MAXEXGROUP];
};
__device__ __forceinline__ bool ex_group(const ExGroups &ex, int mi, int mj) {
  for (int m = 0; m < ex.n; m++) {
    if ((mi & ex.bit1[m]) && (mj & ex.bit2[m])) return true;
    if ((mi & ex.bit2[m]) && (mj & ex.bit1[m])) return true;
  }
  return false;
}

// ghost position on a forward halo = owner + the offsets of the swaps it travels through, each
// added and rounded separately in the reference's order (x, then y, then z swap)
__device__ __forceinline__ void halo_shift(const Geom &g, int dir, double qx, double qy, double qz,
                                           double *o) {
  o[0] = __dadd_rn(__dadd_rn(__dadd_rn(qx, g.fshift[dir][0]), g.tilt[dir][0]), g.tilt[dir][1]);
  o[1] = __dadd_rn(__dadd_rn(qy, g.fshift[dir][1]), g.tilt[dir][2]);
  o[2] = __dadd_rn(qz, g.fshift[dir][2]);
}

// Domain::x2lamda / lamda2x (domain.cpp:2347-2390), every operation rounded separately
__device__ __forceinline__ void x2lamda(const Geom &g, double &x, double &y, double &z) {
  const double d0 = __dadd_rn(x, -g.origin[0]), d1 = __dadd_rn(y, -g.origin[1]), d2 = __dadd_rn(z, -g.origin[2]);
  x = __dadd_rn(__dadd_rn(__dmul_rn(g.h_inv[0], d0), __dmul_rn(g.h_inv[5], d1)), __dmul_rn(g.h_inv[4], d2));
  y = __dadd_rn(__dmul_rn(g.h_inv[1], d1), __dmul_rn(g.h_inv[3], d2));
  z = __dmul_rn(g.h_inv[2], d2);
}
__device__ __forceinline__ void lamda2x(const Geom &g, double &x, double &y, double &z) {
  const double l0 = x, l1 = y, l2 = z;
  x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.h[0], l0), __dmul_rn(g.h[5], l1)), __dmul_rn(g.h[4], l2)), g.origin[0]);
  y = __dadd_rn(__dadd_rn(__dmul_rn(g.h[1], l1), __dmul_rn(g.h[3], l2)), g.origin[1]);
  z = __dadd_rn(__dmul_rn(g.h[2], l2), g.origin[2]);
}

// Half stencil as rows of x-contiguous bins: nstencil_bin.cpp:28-67 regrouped by (dz,dy).
struct Stencil {
  int nrows;
  int rowoff[MAXROWS];  // dz*mbiny*mbinx + dy*mbinx
  int dxlo[MAXROWS], dxhi[MAXROWS];
};

__device__ __forceinline__ int d2type(double w) { return (int)__double_as_longlong(w); }
__device__ __forceinline__ double type2d(int t) { return __longlong_as_double((long long)t); }

// One atom record {x,y,z,type} = 32 bytes = one sector.  sm_100a has a 256-bit global load
// (LDG.E.ENL2.256): the gather of a neighbour costs ONE load instruction / one L1 wavefront set
// instead of the two LDG.128 a plain double4 read compiles to.  Read-only (.nc) path: positions
// are never written by a kernel that gathers them.
__device__ __forceinline__ double4 ld_xt(const double4 *p) {
  double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}

// NBin::coord2bin, nbin.cpp:141-173 (three branches per dimension, truncating casts)
__device__ __forceinline__ int coord2bin_dim(double x, double lo, double hi, double bininv,
                                             int nbin) {
  int ix;
  if (x >= hi)
    ix = (int)((x - hi) * bininv) + nbin;
  else if (x >= lo) {
    ix = (int)((x - lo) * bininv);
    ix = min(ix, nbin - 1);
  } else
    ix = (int)((x - lo) * bininv) - 1;
  return ix;
}

__device__ __forceinline__ int coord2bin(const Geom &g, double x, double y, double z) {
  int ix = coord2bin_dim(x, g.binlo[0], g.binhi[0], g.bininv[0], g.nbin[0]);
  int iy = coord2bin_dim(y, g.binlo[1], g.binhi[1], g.bininv[1], g.nbin[1]);
  int iz = coord2bin_dim(z, g.binlo[2], g.binhi[2], g.bininv[2], g.nbin[2]);
  ix -= g.mbinlo[0];
  iy -= g.mbinlo[1];
  iz -= g.mbinlo[2];
  // a bin outside the local grid means the atom moved further than the ghost shell allows
  if (ix < 0 || ix >= g.mbin[0] || iy < 0 || iy >= g.mbin[1] || iz < 0 || iz >= g.mbin[2]) return -1;
  return (iz * g.mbin[1] + iy) * g.mbin[0] + ix;
}

// rsq exactly as the reference's x86-64 build evaluates delx*delx + dely*dely + delz*delz
// (separate multiplies and adds, no FMA contraction): the cutoff decisions `rsq <= cutneighsq`
// and `rsq < cutsq` then agree bit-for-bit with the CPU path.
__device__ __forceinline__ double rsq_ref(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Reciprocal and square root without the IEEE special-case paths the compiler emits for `1.0/x`
// and `sqrt(x)` (a slow-path call per use that also stops it interleaving neighbouring pairs):
// MUFU seed + Newton steps, <= 1 ulp for the normal, positive arguments a pair distance can be.
__device__ __forceinline__ double rcp_nr(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  e = fma(e, e, e);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  return fma(x, e, x);
}
__device__ __forceinline__ double sqrt_nr(double a, double &rinv) {  // returns sqrt(a), rinv ~ 1/sqrt(a)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, a);
  g = fma(d, h, g);
  rinv = h + h;
  return g;
}
__device__ __forceinline__ float rcp_f(float a) {
  float x;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));
  return x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  blockDim.x <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* [NV*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; k++) smem[k * 32 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double t = lane < nwarp ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
}
