// kernels_pair_mixed.cuh -- mixed-precision pair kernels (package b200 prec mixed).
//
// What stays FP64: positions, the separation vector del = x_i - x_j, rsq and therefore every
// cutoff decision (the set of interacting pairs is identical to double mode), the per-atom
// accumulators of rho_i / energy, the force arrays the integrator reads, and all integration.  What runs in FP32: the pair
// function itself (1/rsq, r^-6, spline evaluation, force components).  The Newton scatter onto
// atom j is ONE 16-byte vector reduction (RED.ADD.F32x4, sm_90+) into a float4 force array
// instead of three 8-byte FP64 reductions: a third of the L2 atomic sector traffic, which is
// what bounds the FP64 kernel (profiles/r01a_ncu_full_k_pair_lj_lj4m.txt).  k_merge_ff then folds
// the float4 array into the FP64 SoA forces that the rest of the step consumes.
// Tolerances (BASELINE.json): forces <= 1e-5 relative, energy/pressure <= 1e-6.
#pragma once
#include "common.cuh"
#include "kernels_pair.cuh"

template <int T>
__device__ __forceinline__ float group_sumf(float v) {
#pragma unroll
  for (int o = T / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct LJOneF {
  float lj1, lj2, lj3, lj4, offset;
};

// PairLJCut::compute (pair_lj_cut.cpp:71-141), FP32 pair math.
// tabf: 5 float tables lj1, lj2, lj3, lj4, offset [(ntypes+1)^2 each]; cutsq stays double (tabd).
template <bool EV, bool ONETYPE, int T>
__global__ void __launch_bounds__(128) k_pair_lj_mixed(
    int nlocal, int nstride, const double4 *__restrict__ xt, const int *__restrict__ numneigh,
    const int *__restrict__ neigh, double *__restrict__ fx, double *__restrict__ fy,
    double *__restrict__ fz, float4 *__restrict__ ff, double cutsq_one, LJOneF one,
    const double *__restrict__ cutsq_tab, const float *__restrict__ tabf, int ntypes,
    double *__restrict__ ev) {
  extern __shared__ double smem_mixed[];
  const int n1 = ntypes + 1, n2 = n1 * n1;
  double *scut = smem_mixed;                                   // [n2]
  float *stab = reinterpret_cast<float *>(smem_mixed + n2);    // [5*n2]
  if (!ONETYPE) {
    for (int k = threadIdx.x; k < n2; k += blockDim.x) scut[k] = cutsq_tab[k];
    for (int k = threadIdx.x; k < 5 * n2; k += blockDim.x) stab[k] = tabf[k];
    __syncthreads();
  }
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double evdwl = 0.0;
  float fxi = 0.0f, fyi = 0.0f, fzi = 0.0f;  // < 64 terms: FP32 is ample for the 1e-5 bound
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
      int tij = 0;
      double cutsq = cutsq_one;
      if (!ONETYPE) {
        tij = itype * n1 + d2type(pj.w);
        cutsq = scut[tij];
      }
      if (rsq < cutsq) {
        const float lj1 = ONETYPE ? one.lj1 : stab[tij];
        const float lj2 = ONETYPE ? one.lj2 : stab[n2 + tij];
        const float r2inv = 1.0f / (float)rsq;
        const float r6inv = r2inv * r2inv * r2inv;
        const float fpair = r6inv * (lj1 * r6inv - lj2) * r2inv;
        const float gx = (float)delx * fpair, gy = (float)dely * fpair, gz = (float)delz * fpair;
        fxi += gx;
        fyi += gy;
        fzi += gz;
        atomicAdd(&ff[j], make_float4(-gx, -gy, -gz, 0.0f));
        if (EV) {
          const float lj3 = ONETYPE ? one.lj3 : stab[2 * n2 + tij];
          const float lj4 = ONETYPE ? one.lj4 : stab[3 * n2 + tij];
          const float off = ONETYPE ? one.offset : stab[4 * n2 + tij];
          evdwl += (double)(r6inv * (lj3 * r6inv - lj4) - off);
        }
      }
    }
  }
  if (T > 1) {
    fxi = group_sumf<T>(fxi);
    fyi = group_sumf<T>(fyi);
    fzi = group_sumf<T>(fzi);
  }
  if (i < nlocal && t == 0) {
    fx[i] = (double)fxi;  // the scatter half of f_i arrives through ff (k_merge_ff)
    fy[i] = (double)fyi;
    fz[i] = (double)fzi;
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {evdwl};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}

// owned: f (written by the pair kernel) += ff ; ghosts: f = ff.  16 B read + 24 B write (+24 B
// read for owned atoms) per atom, one coalesced pass; replaces force_clear of f in mixed mode.
__global__ void __launch_bounds__(256) k_merge_ff(int nall, int nlocal,
                                                  const float4 *__restrict__ ff,
                                                  double *__restrict__ fx, double *__restrict__ fy,
                                                  double *__restrict__ fz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nall) return;
  const float4 g = ff[i];
  if (i < nlocal) {
    fx[i] += (double)g.x;
    fy[i] += (double)g.y;
    fz[i] += (double)g.z;
  } else {
    fx[i] = (double)g.x;
    fy[i] = (double)g.y;
    fz[i] = (double)g.z;
  }
}

// ------------------------------------------------------------------------------- EAM
// float spline tables, 8 floats per knot: {c0..c6, pad} -> one 32-byte sector per lookup
struct EAMParamsF {
  const float *rhor, *z2r;  // device [n][nr+1][8]
  float rdr;
};

// phase 1 (pair_eam.cpp:163-211): FP32 spline evaluation, FP64 rho accumulators / reductions
template <int T>
__global__ void __launch_bounds__(128) k_eam_rho_mixed(int nlocal, int nstride,
                                                       const double4 *__restrict__ xt,
                                                       const int *__restrict__ numneigh,
                                                       const int *__restrict__ neigh, EAMParams P,
                                                       EAMParamsF F, double *__restrict__ rho) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double rhoi = 0.0;
  const bool active = i < nlocal;
  const double4 pi = active ? xt[i] : make_double4(0, 0, 0, 0);
  const int itype = d2type(pi.w), n1 = P.ntypes + 1;
  const int jnum = active ? numneigh[i] : 0;
  const int *jl = neigh + (size_t)i * T + t;
#pragma unroll 4
  for (int n = t, kk = 0; n < jnum; n += T, kk++) {
    const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
    const double4 pj = ld_xt(xt + j);
    const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
    const double rsq = rsq_ref(delx, dely, delz);
    if (rsq < P.cutforcesq) {
      const int jtype = d2type(pj.w);
      const float rsqf = (float)rsq;
      float p = rsqf * rsqrtf(rsqf) * F.rdr + 1.0f;  // MUFU.RSQ, no IEEE slow path
      int m = (int)p;
      m = min(m, P.nr - 1);
      p -= (float)m;
      p = fminf(p, 1.0f);
      const int tji = P.type2rhor[jtype * n1 + itype], tij = P.type2rhor[itype * n1 + jtype];
      // knot = {c0,c1,c2,c3 | c4,c5,c6,pad}: the value cubic is c3..c6 (one 32-byte sector)
      const float4 *kc = reinterpret_cast<const float4 *>(F.rhor + ((size_t)tji * (P.nr + 1) + m) * 8);
      const float4 c0 = __ldg(kc), c1 = __ldg(kc + 1);
      const float rj = ((c0.w * p + c1.x) * p + c1.y) * p + c1.z;
      rhoi += (double)rj;
      float ri = rj;
      if (tij != tji) {
        const float *d = F.rhor + ((size_t)tij * (P.nr + 1) + m) * 8;
        ri = ((__ldg(d + 3) * p + __ldg(d + 4)) * p + __ldg(d + 5)) * p + __ldg(d + 6);
      }
      atomicAdd(&rho[j], (double)ri);
    }
  }
  if (T > 1) rhoi = group_sum<T>(rhoi);
  if (active && t == 0) atomicAdd(&rho[i], rhoi);
}

// phase 3 (pair_eam.cpp:233-314): FP32 force evaluation, one RED.ADD.F32x4 per pair
template <bool EV, int T>
__global__ void __launch_bounds__(128) k_eam_force_mixed(
    int nlocal, int nstride, const double4 *__restrict__ xt, const int *__restrict__ numneigh,
    const int *__restrict__ neigh, EAMParams P, EAMParamsF F, const double *__restrict__ fp,
    double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
    float4 *__restrict__ ff, double *__restrict__ ev) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / T, t = tid % T;
  double evdwl = 0.0;
  float fxi = 0.0f, fyi = 0.0f, fzi = 0.0f;  // < 64 terms: FP32 is ample for the 1e-5 bound
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w), n1 = P.ntypes + 1;
    const int jnum = numneigh[i];
    const int *jl = neigh + (size_t)i * T + t;
    const float fpi = (float)fp[i];
#pragma unroll 2
    for (int n = t, kk = 0; n < jnum; n += T, kk++) {
      const int j = jl[(size_t)kk * nstride * T] & NEIGHMASK;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      if (rsq < P.cutforcesq) {
        const int jtype = d2type(pj.w);
        const float rsqf = (float)rsq;
        const float recip = rsqrtf(rsqf);  // MUFU.RSQ, no IEEE slow paths
        const float r = rsqf * recip;
        float p = r * F.rdr + 1.0f;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= (float)m;
        p = fminf(p, 1.0f);
        const int tij = P.type2rhor[itype * n1 + jtype], tji = P.type2rhor[jtype * n1 + itype];
        const float4 a = __ldg(reinterpret_cast<const float4 *>(F.rhor + ((size_t)tij * (P.nr + 1) + m) * 8));
        const float rhoip = (a.x * p + a.y) * p + a.z;
        float rhojp = rhoip;
        if (tji != tij) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(F.rhor + ((size_t)tji * (P.nr + 1) + m) * 8));
          rhojp = (b.x * p + b.y) * p + b.z;
        }
        const float4 *zc = reinterpret_cast<const float4 *>(
            F.z2r + ((size_t)P.type2z2r[itype * n1 + jtype] * (P.nr + 1) + m) * 8);
        const float4 z0 = __ldg(zc), z1 = __ldg(zc + 1);
        const float z2p = (z0.x * p + z0.y) * p + z0.z;
        const float z2 = ((z0.w * p + z1.x) * p + z1.y) * p + z1.z;
        const float phi = z2 * recip;
        const float phip = z2p * recip - phi * recip;
        const float psip = fpi * rhojp + (float)fp[j] * rhoip + phip;
        const float sc = (float)P.scale[itype * n1 + jtype];
        const float fpair = -sc * psip * recip;
        const float gx = (float)delx * fpair, gy = (float)dely * fpair, gz = (float)delz * fpair;
        fxi += gx;
        fyi += gy;
        fzi += gz;
        atomicAdd(&ff[j], make_float4(-gx, -gy, -gz, 0.0f));
        if (EV) evdwl += (double)(sc * phi);
      }
    }
  }
  if (T > 1) {
    fxi = group_sumf<T>(fxi);
    fyi = group_sumf<T>(fyi);
    fzi = group_sumf<T>(fzi);
  }
  if (i < nlocal && t == 0) {
    fx[i] = (double)fxi;
    fy[i] = (double)fyi;
    fz[i] = (double)fzi;
  }
  if (EV) {
    __shared__ double red[32];
    double v[1] = {evdwl};
    block_sum<1>(v, red);
    if (threadIdx.x == 0) atomicAdd(&ev[0], v[0]);
  }
}
