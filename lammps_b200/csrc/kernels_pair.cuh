// kernels_pair.cuh -- pair force/energy kernels over the half (Newton on) Verlet list.
//   lj/cut : PairLJCut::compute                pair_lj_cut.cpp:71-141
//   eam    : PairEAM::compute (3 phases)       pair_eam.cpp:124-327, 338-366, pair_eam.h:146-169
//
// Layout: positions+type are one 32-byte record double4{x,y,z,type} per atom (one sector, one
// 256-bit load per gather); forces are SoA fx/fy/fz.  T lanes share one atom ("threads per
// atom"): lane t walks neighbours t, t+T, t+2T, ... and the T partial sums of f_i are combined
// with warp shuffles; only the Newton scatter onto atom j uses atomics (RED.ADD.F64).
// What bounds these kernels (ncu, profiles/r01c_*): the L1TEX data pipe.  Every lane of a
// gather or scatter touches its own 32-byte sector (one atom record = one sector; three
// RED.F64 per pair), so a warp-wide access costs ~32 sector wavefronts however the lanes are
// arranged: l1tex__data_pipe_lsu_wavefronts sits at 77-78 % of peak, DRAM at 12 %.  Measured
// T sweep on 4 M LJ atoms: T=1 1265 us, T=2 1178 us, T=4 1220 us, T=8 1365 us -> default T=2.
// The list is stored so that a warp reads 32 consecutive ints per slot group:
//   entry n of atom i lives at neigh[((n / T) * nstride + i) * T + n % T].
// No tensor cores: nothing here is a dense contraction.
#pragma once
#include "common.cuh"

__host__ __device__ __forceinline__ size_t list_index(int n, int i, int nstride, int T) {
  return ((size_t)(n / T) * nstride + i) * T + (n % T);
}

// sum over the T lanes that share an atom (T a power of two <= 32, lanes contiguous)
template <int T>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct LJOne {  // single-type fast path: coefficients travel as kernel arguments
  double cutsq, lj1, lj2, lj3, lj4, offset;
};

// ev[0] += eng_vdwl (only when EV)
template <bool EV, bool ONETYPE, int T>
__global__ void __launch_bounds__(128) k_pair_lj(int nlocal, int nstride,
                                                 const double4 *__restrict__ xt,
                                                 const int *__restrict__ numneigh,
                                                 const int *__restrict__ neigh,
                                                 double *__restrict__ fx, double *__restrict__ fy,
                                                 double *__restrict__ fz, LJOne one,
                                                 const double *__restrict__ tab /*6 tables*/,
                                                 int ntypes, double *__restrict__ ev) {
  extern __shared__ double stab[];  // !ONETYPE: cutsq, lj1, lj2, lj3, lj4, offset  [(ntypes+1)^2 each]
  const int n1 = ntypes + 1, n2 = n1 * n1;
  if (!ONETYPE) {
    for (int k = threadIdx.x; k < 6 * n2; k += blockDim.x) stab[k] = tab[k];
    __syncthreads();
  }
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double evdwl = 0.0, fxi = 0.0, fyi = 0.0, fzi = 0.0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w);
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
#pragma unroll 4
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      double cutsq, lj1, lj2;
      int tij = 0;
      if (ONETYPE) {
        cutsq = one.cutsq; lj1 = one.lj1; lj2 = one.lj2;
      } else {
        tij = itype * n1 + d2type(pj.w);
        cutsq = stab[tij]; lj1 = stab[n2 + tij]; lj2 = stab[2 * n2 + tij];
      }
      if (rsq < cutsq) {
        const double r2inv = 1.0 / rsq;
        const double r6inv = r2inv * r2inv * r2inv;
        const double forcelj = r6inv * (lj1 * r6inv - lj2);
        const double fpair = forcelj * r2inv;
        fxi += delx * fpair;
        fyi += dely * fpair;
        fzi += delz * fpair;
        atomicAdd(&fx[j], -(delx * fpair));
        atomicAdd(&fy[j], -(dely * fpair));
        atomicAdd(&fz[j], -(delz * fpair));
        if (EV) {
          const double lj3 = ONETYPE ? one.lj3 : stab[3 * n2 + tij];
          const double lj4 = ONETYPE ? one.lj4 : stab[4 * n2 + tij];
          const double off = ONETYPE ? one.offset : stab[5 * n2 + tij];
          evdwl += r6inv * (lj3 * r6inv - lj4) - off;
        }
      }
    }
  }
  if (T > 1) {
    fxi = group_sum<T>(fxi);
    fyi = group_sum<T>(fyi);
    fzi = group_sum<T>(fzi);
  }
  if (i < nlocal && t == 0) {
    atomicAdd(&fx[i], fxi);
    atomicAdd(&fy[i], fyi);
    atomicAdd(&fz[i], fzi);
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {evdwl};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}

// Pair::virial_fdotr_compute (pair.cpp:1809-1825): sum over owned+ghost of f (x) x, taken
// AFTER the pair kernel and BEFORE reverse comm.  ev[1..6] += virial.
__global__ void __launch_bounds__(256) k_virial_fdotr(int nall, const double4 *__restrict__ xt,
                                                      const double *__restrict__ fx,
                                                      const double *__restrict__ fy,
                                                      const double *__restrict__ fz,
                                                      double *__restrict__ ev) {
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
    const double4 p = xt[i];
    const double a = fx[i], b = fy[i], c = fz[i];
    v[0] += a * p.x; v[1] += b * p.y; v[2] += c * p.z;
    v[3] += b * p.x; v[4] += c * p.x; v[5] += c * p.y;
  }
  __shared__ double red[6 * 32];
  block_sum<6>(v, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 6; k++) atomicAdd(&ev[1 + k], v[k]);
}

// ------------------------------------------------------------------------------- EAM
struct EAMParams {
  int nr, nrho, ntypes;
  double rdr, rdrho, rhomax, cutforcesq;
  const int *type2frho, *type2rhor, *type2z2r;  // device, [(ntypes+1)] and [(ntypes+1)^2]
  const double *scale;                          // device [(ntypes+1)^2]
  const double *frho, *rhor, *z2r;              // device splines [n][nr+1|nrho+1][7]
};

// phase 1, pair_eam.cpp:163-211: rho_i += rho_j(r), rho_j += rho_i(r) for rsq < cutforcesq
template <int T>
__global__ void __launch_bounds__(128) k_eam_rho(int nlocal, int nstride,
                                                 const double4 *__restrict__ xt,
                                                 const int *__restrict__ numneigh,
                                                 const int *__restrict__ neigh, EAMParams P,
                                                 double *__restrict__ rho) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double rhoi = 0.0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w), n1 = P.ntypes + 1;
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
#pragma unroll 4
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      if (rsq < P.cutforcesq) {
        const int jtype = d2type(pj.w);
        double rinv;
        double p = sqrt_nr(rsq, rinv) * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        const int tji = P.type2rhor[jtype * n1 + itype], tij = P.type2rhor[itype * n1 + jtype];
        const double *c = P.rhor + ((size_t)tji * (P.nr + 1) + m) * 7;
        const double rj = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
        rhoi += rj;
        double ri = rj;
        if (tij != tji) {
          c = P.rhor + ((size_t)tij * (P.nr + 1) + m) * 7;
          ri = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
        }
        atomicAdd(&rho[j], ri);
      }
    }
  }
  if (T > 1) rhoi = group_sum<T>(rhoi);
  if (i < nlocal && t == 0) atomicAdd(&rho[i], rhoi);
}

// phase 2, compute_embedding<0> (pair_eam.cpp:338-366) + embedding_index<0> (pair_eam.h:146-169)
// err |= 4 when rho > rhomax on an energy step (the reference warns once).
template <bool EV>
__global__ void __launch_bounds__(256) k_eam_embed(int nlocal, const double4 *__restrict__ xt,
                                                   EAMParams P, const double *__restrict__ rho,
                                                   double *__restrict__ fp, double *__restrict__ ev,
                                                   int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double phi = 0.0;
  if (i < nlocal) {
    const int itype = d2type(xt[i].w), n1 = P.ntypes + 1;
    const double r = rho[i];
    double p = r * P.rdrho + 1.0;
    int m = (int)p;
    m = max(1, min(m, P.nrho - 1));
    p -= m;
    p = fmin(p, 1.0);
    const double *c = P.frho + ((size_t)P.type2frho[itype] * (P.nrho + 1) + m) * 7;
    const double fpi = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
    fp[i] = fpi;
    if (EV) {
      phi = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
      if (r > P.rhomax) {
        phi += fpi * (r - P.rhomax);
        atomicOr(err, 4);
      }
      phi *= P.scale[itype * n1 + itype];
    }
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {phi};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}

// phase 3, pair_eam.cpp:233-314: force from fp_i, fp_j, rho', z2r
template <bool EV, int T>
__global__ void __launch_bounds__(128) k_eam_force(int nlocal, int nstride,
                                                   const double4 *__restrict__ xt,
                                                   const int *__restrict__ numneigh,
                                                   const int *__restrict__ neigh, EAMParams P,
                                                   const double *__restrict__ fp,
                                                   double *__restrict__ fx, double *__restrict__ fy,
                                                   double *__restrict__ fz, double *__restrict__ ev) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double evdwl = 0.0, fxi = 0.0, fyi = 0.0, fzi = 0.0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w), n1 = P.ntypes + 1;
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
    const double fpi = fp[i];
#pragma unroll 2
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      if (rsq < P.cutforcesq) {
        const int jtype = d2type(pj.w);
        double recip;  // sqrt and 1/r without the IEEE slow paths (<= 1 ulp; rsq > 0 here)
        const double r = sqrt_nr(rsq, recip);
        recip = fma(recip, fma(-r, recip, 1.0), recip);
        double p = r * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        const int tij = P.type2rhor[itype * n1 + jtype], tji = P.type2rhor[jtype * n1 + itype];
        const double *c = P.rhor + ((size_t)tij * (P.nr + 1) + m) * 7;
        const double rhoip = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        double rhojp = rhoip;
        if (tji != tij) {
          c = P.rhor + ((size_t)tji * (P.nr + 1) + m) * 7;
          rhojp = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        }
        c = P.z2r + ((size_t)P.type2z2r[itype * n1 + jtype] * (P.nr + 1) + m) * 7;
        const double z2p = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        const double z2 = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
        const double phi = z2 * recip;
        const double phip = z2p * recip - phi * recip;
        const double psip = fpi * rhojp + fp[j] * rhoip + phip;
        const double sc = P.scale[itype * n1 + jtype];
        const double fpair = -sc * psip * recip;
        fxi += delx * fpair;
        fyi += dely * fpair;
        fzi += delz * fpair;
        atomicAdd(&fx[j], -(delx * fpair));
        atomicAdd(&fy[j], -(dely * fpair));
        atomicAdd(&fz[j], -(delz * fpair));
        if (EV) evdwl += sc * phi;
      }
    }
  }
  if (T > 1) {
    fxi = group_sum<T>(fxi);
    fyi = group_sum<T>(fyi);
    fzi = group_sum<T>(fzi);
  }
  if (i < nlocal && t == 0) {
    atomicAdd(&fx[i], fxi);
    atomicAdd(&fy[i], fyi);
    atomicAdd(&fz[i], fzi);
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {evdwl};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}

// ------------------------------------------------------------------- EAM, one atom type
// The flat eam kernels are bound by the L1 data pipe, and most of its load is the spline
// coefficients: 10 scattered 8-byte reads per in-cutoff pair (profiles/r01d_ncu_full_k_eam_force.txt:
// 484 M load sectors for 45 M in-cutoff pairs).  For a single-element potential (funcfl, the
// bench input) the two r-space tables are indexed by the same knot, so the host packs them:
//   rho4[m] = {c3,c4,c5,c6} of rhor                     (32 B, two 16-byte loads)
//   frc8[m] = {c3,c4,c5 of rhor, -, c3,c4,c5,c6 of z2r} (64 B, four 16-byte loads)
// The derivative quadratic is not stored: PairEAM::array2spline builds it from the same cubic
// (c0 = 3 c3/dr, c1 = 2 c4/dr, c2 = c5/dr, pair_eam.cpp:1517-1545), so
// f'(p) = rdr * ((3 c3 p + 2 c4) p + c5) to rounding (<= 2 ulp, far inside the 1e-12 bound).
struct EAMFast {
  const double2 *rho4;  // [(nr+1)][2]
  const double2 *frc8;  // [(nr+1)][4]
  double scale;
};

template <int T>
__global__ void __launch_bounds__(128) k_eam_rho_one(int nlocal, int nstride,
                                                     const double4 *__restrict__ xt,
                                                     const int *__restrict__ numneigh,
                                                     const int *__restrict__ neigh, EAMParams P,
                                                     EAMFast F, double *__restrict__ rho) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double rhoi = 0.0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
#pragma unroll 4
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double rsq = rsq_ref(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      if (rsq < P.cutforcesq) {
        double rinv;
        double p = sqrt_nr(rsq, rinv) * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        // one 32-byte record per knot, one LDG.256: a scattered load costs the L1 data pipe a
        // wavefront per lane whatever its width (profiles/r02z_ncu_eam.txt)
        const double4 a = ld_xt(reinterpret_cast<const double4 *>(F.rho4 + 2 * m));
        const double rj = ((a.x * p + a.y) * p + a.z) * p + a.w;
        rhoi += rj;
        atomicAdd(&rho[j], rj);
      }
    }
  }
  if (T > 1) rhoi = group_sum<T>(rhoi);
  if (i < nlocal && t == 0) atomicAdd(&rho[i], rhoi);
}

template <bool EV, int T>
__global__ void __launch_bounds__(128) k_eam_force_one(
    int nlocal, int nstride, const double4 *__restrict__ xt, const int *__restrict__ numneigh,
    const int *__restrict__ neigh, EAMParams P, EAMFast F, const double *__restrict__ fp,
    double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
    double *__restrict__ ev) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double evdwl = 0.0, fxi = 0.0, fyi = 0.0, fzi = 0.0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
    const double fpi = fp[i];
#pragma unroll 2
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      if (rsq < P.cutforcesq) {
        double recip;  // sqrt and 1/r without the IEEE slow paths (<= 1 ulp; rsq > 0 here)
        const double r = sqrt_nr(rsq, recip);
        recip = fma(recip, fma(-r, recip, 1.0), recip);
        double p = r * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        const double4 *c = reinterpret_cast<const double4 *>(F.frc8 + 4 * m);  // two LDG.256
        const double4 qa = ld_xt(c), qb = ld_xt(c + 1);
        const double rhop = P.rdr * ((3.0 * qa.x * p + 2.0 * qa.y) * p + qa.z);   // rhoip == rhojp
        const double z2p = P.rdr * ((3.0 * qb.x * p + 2.0 * qb.y) * p + qb.z);
        const double z2 = ((qb.x * p + qb.y) * p + qb.z) * p + qb.w;
        const double phi = z2 * recip;
        const double phip = z2p * recip - phi * recip;
        const double psip = fpi * rhop + fp[j] * rhop + phip;
        const double fpair = -F.scale * psip * recip;
        fxi += delx * fpair;
        fyi += dely * fpair;
        fzi += delz * fpair;
        atomicAdd(&fx[j], -(delx * fpair));
        atomicAdd(&fy[j], -(dely * fpair));
        atomicAdd(&fz[j], -(delz * fpair));
        if (EV) evdwl += F.scale * phi;
      }
    }
  }
  if (T > 1) {
    fxi = group_sum<T>(fxi);
    fyi = group_sum<T>(fyi);
    fzi = group_sum<T>(fzi);
  }
  if (i < nlocal && t == 0) {
    atomicAdd(&fx[i], fxi);
    atomicAdd(&fy[i], fyi);
    atomicAdd(&fz[i], fzi);
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {evdwl};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}
