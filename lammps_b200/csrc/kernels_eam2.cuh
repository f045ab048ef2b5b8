// kernels_eam2.cuh -- second-generation eam kernels over the bin-tile list (single-element
// potentials: funcfl, or setfl/fs files with one element; FP64).
//
// PairEAM::compute (pair_eam.cpp:124-327) as two tile kernels and ONE halo:
//   k_tile_eam2_rho    rho_i = sum_j rho(r_ij) over the FULL row of atom i (every partner, owned
//                      or ghost, stored by k_tile_build<.,FULLGHOST,SPLIT>), then -- the density
//                      of an owned atom being complete in the thread that owns it -- the embedding
//                      step of pair_eam.cpp:219-231 in the epilogue: fp_i = F'(rho_i), and
//                      F(rho_i) on energy steps.  No RED.F64 onto partners, no density reverse
//                      halo (pair_eam.cpp:215), no k_eam_embed launch, no rho clear.
//   forward halo of fp (pair_eam.cpp:233, 1600-1621): the ghost-density exchange that remains.
//   k_tile_eam2_force  f_i = sum_j fpair(r_ij) del over the same row, fp_j staged next to the
//                      positions; f_i is STORED (no atomics, no force clear, no reverse halo), or
//                      -- NVE -- not stored at all: fix nve's final_integrate of this step and
//                      initial_integrate of the next are applied in the epilogue (NveFuse,
//                      kernels_tile2.cuh).  Energy and virial are tallied on the FWD entries (the
//                      reference's half list), so every pair counts once across sub-domains.
// What it buys and what it costs, measured on one B200 (profiles/r02z_ncu_eam.txt, r02ab_*):
//   + three launches and one halo per step instead of eight launches and three halos: 32 k atoms
//     (bench/in.eam) 249 vs 215 M atom-steps/s; forces are sums in a fixed order (no atomics), so
//     lmp_b200 reproduces the golden in.eam neighbour count exactly (the flat path's RED order
//     noise moved up to 200 of 1.2 M skin-shell pairs);
//   - every owned-owned pair is evaluated from both sides, and what loads the L1 data pipe in
//     BOTH formulations is the scattered spline-table reads (one wavefront per lane and read),
//     not the Newton scatter: ncu shows l1tex__data_pipe_lsu_wavefronts at 91 % with 68 % of the
//     wavefronts from the table loads.  At 2 M atoms the pair phase takes 1.70 ms against 1.33 ms
//     for the flat kernels and the FULLGHOST list build 2.0 against 0.7 ms, so large sub-domains
//     stay on the flat half list (engine.cu: eam2_usable).
// The build-time NEAR/FAR split keeps the warp from running the pair function for partners that
// sit in the skin: NEAR = stored within force cutoff + margin at build time (about 54 of 78
// entries per atom for Cu at 800 K) is evaluated unconditionally, FAR behind a test that rarely
// fires.  A per-lane `if (rsq < cutsq)` without the split saves nothing: some lane of the 32 is
// inside the cutoff for almost every entry.
// Cutoff decisions use the reference's rsq (rsq_ref): the set of interacting pairs is the CPU
// path's.  sqrt and 1/r: MUFU.RSQ64H seed (20 bits) + one Goldschmidt step + one correction
// (error ~2^-78 before rounding, <= 1 ulp); spline knots as packed by b200_pair_eam
// (EAMFast: value cubics only, derivatives derived from them, pair_eam.cpp:1517-1545).
// No tensor cores: nothing here is a dense contraction.
#pragma once
#include "kernels_tile2.cuh"

// sqrt(a) and ~1/sqrt(a) for a normal positive a (see header)
__device__ __forceinline__ double sqrt_fast(double a, double &rinv) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double g = a * y, h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, a);
  g = fma(d, h, g);
  rinv = h + h;
  return g;
}

__host__ __device__ __forceinline__ size_t eam2_smem_bytes(int scap, bool with_fp) {
  return (TILE_HDR_BYTES + (size_t)scap * (with_fp ? 32 : 24) + 127) / 128 * 128;
}

// spline knot of distance r: m = knot, p = offset in [0,1] (pair_eam.cpp:196-200)
__device__ __forceinline__ void eam2_knot(double r, double rdr, int nr, int &m, double &p) {
  p = fma(r, rdr, 1.0);
  m = min(__double2int_rz(p), nr - 1);
  p = fmin(p - (double)m, 1.0);
}

// --------------------------------------------------------------------------- density + embedding
template <bool EV, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_tile_eam2_rho(
    int nlocal, const double4 *__restrict__ xt, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const unsigned short *__restrict__ tfar, const int *__restrict__ tgi, const uint4 *__restrict__ list,
    EAMParams P, EAMFast F, double *__restrict__ rho, double *__restrict__ fp, double *__restrict__ ev,
    int *__restrict__ err, int scap, int *__restrict__ tflags, const int *__restrict__ tile_ids,
    const unsigned char *__restrict__ hdrs) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  double *pos = reinterpret_cast<double *>(tsm + TILE_HDR_BYTES);  // {x,y} pairs, then z
  double *posz = pos + (size_t)2 * scap;
  int *chunk_ctr = reinterpret_cast<int *>(&H->pad0);
  const unsigned pos_s = (unsigned)__cvta_generic_to_shared(pos);
  const unsigned posz_s = (unsigned)__cvta_generic_to_shared(posz);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int S = tile_rows_cached(hdrs, tile, H);
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {
    const int nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k, s = base + k;
        const double *p = reinterpret_cast<const double *>(xt + src);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(pos_s + (unsigned)s * 16u), "l"(p)
                     : "memory");
        cp_async8(posz + s, p + 2);
      }
    }
    if (tid < 2) pos[2 * S + tid] = TILE2_FAR;  // the dummy atom padding entries name
    if (tid == 2) posz[S] = TILE2_FAR;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  const int ni = H->ni, ibase = tile_ibase[tile], W = maxslots >> 3;
  const double cutsq = P.cutforcesq, rdr = P.rdr;
  const int nr = P.nr;
  double phisum = 0.0;
  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(chunk_ctr, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk * 32 >= ni) break;
    const int ti = chunk * 32 + lane;
    if (ti >= ni) continue;
    const int g = ibase + ti;
    const uint4 *lp = list + g;
    const int li = iloc[g];
    const int ntot = min((int)tnum[g], maxslots), nfar = min((int)tfar[g], ntot), nnear = ntot - nfar;
    const int gi = tgi[g];
    const double pix = pos[2 * li], piy = pos[2 * li + 1], piz = posz[li];
    double rhoi = 0.0;

    struct P3 { double x, y, z; };
    auto ldpos = [&](unsigned e) -> P3 {
      P3 p;
      const unsigned j = e & TILE_IDX;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.x), "=d"(p.y) : "r"(pos_s + j * 16u));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(p.z) : "r"(posz_s + j * 8u));
      return p;
    };
    auto entry = [&](const uint4 &w, int k) -> unsigned {
      const unsigned v = k < 2 ? w.x : (k < 4 ? w.y : (k < 6 ? w.z : w.w));
      return (k & 1) ? (v >> 16) : (v & 0xffffu);
    };
    // density of one partner at squared distance rsq (pair_eam.cpp:190-206), 0 outside the cutoff
    auto dens = [&](double rsq) -> double {
      double rinv;
      const double r = sqrt_fast(rsq, rinv);
      int m;
      double p;
      eam2_knot(r, rdr, nr, m, p);
      const double4 a = ld_xt(reinterpret_cast<const double4 *>(F.rho4 + 2 * m));  // one LDG.256
      return fma(fma(fma(a.x, p, a.y), p, a.z), p, a.w);
    };
    // NEAR words: every entry evaluated, two at a time, branch-free
    for (int k0 = 0; k0 < nnear; k0 += 8) {
      const uint4 c = __ldg(lp + (size_t)(k0 >> 3) * NI);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const P3 a = ldpos(entry(c, k)), b = ldpos(entry(c, k + 1));
        const double ra = rsq_ref(pix - a.x, piy - a.y, piz - a.z), rb = rsq_ref(pix - b.x, piy - b.y, piz - b.z);
        const double da = dens(fmin(ra, cutsq)), db = dens(fmin(rb, cutsq));
        rhoi += ra < cutsq ? da : 0.0;
        rhoi += rb < cutsq ? db : 0.0;
      }
    }
    // FAR words (stored from the end of the row): the pair function only where a partner came in
    for (int k0 = 0; k0 < nfar; k0 += 8) {
      const uint4 c = __ldg(lp + (size_t)(W - 1 - (k0 >> 3)) * NI);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const P3 a = ldpos(entry(c, k)), b = ldpos(entry(c, k + 1));
        const double ra = rsq_ref(pix - a.x, piy - a.y, piz - a.z), rb = rsq_ref(pix - b.x, piy - b.y, piz - b.z);
        if (ra < cutsq) rhoi += dens(ra);
        if (rb < cutsq) rhoi += dens(rb);
      }
    }
    // embedding (pair_eam.cpp:219-231, compute_embedding): fp = F'(rho), phi = F(rho)
    {
      double p = rhoi * P.rdrho + 1.0;
      int m = (int)p;
      m = max(1, min(m, P.nrho - 1));
      p -= m;
      p = fmin(p, 1.0);
      const double *c = P.frho + ((size_t)P.type2frho[1] * (P.nrho + 1) + m) * 7;
      const double fpi = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
      rho[gi] = rhoi;
      fp[gi] = fpi;
      if (EV) {
        double phi = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
        if (rhoi > P.rhomax) {
          phi += fpi * (rhoi - P.rhomax);
          atomicOr(err, 4);
        }
        phisum += phi * F.scale;
      }
    }
  }
  if (EV) {
    double v[1] = {phisum};
    __syncthreads();
    block_sum<1>(v, pos);
    if (tid == 0) atomicAdd(&ev[0], v[0]);
  }
}

// --------------------------------------------------------------------------------------- force
template <bool EV, int MAXT, int MINB, bool NVE = false>
__global__ void __launch_bounds__(MAXT, MINB) k_tile_eam2_force(
    int nlocal, const double4 *__restrict__ xt, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const unsigned short *__restrict__ tfar, const int *__restrict__ tgi, const uint4 *__restrict__ list,
    EAMParams P, EAMFast F, const double *__restrict__ fp, double *__restrict__ fx,
    double *__restrict__ fy, double *__restrict__ fz, double *__restrict__ ev, int scap,
    int *__restrict__ tflags, const int *__restrict__ tile_ids, NveFuse nv,
    const unsigned char *__restrict__ hdrs) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  // staged atom s: {x,y} at pxy[s], {z,fp} at pzf[s] -- two LDS.128 per entry
  double2 *pxy = reinterpret_cast<double2 *>(tsm + TILE_HDR_BYTES);
  double2 *pzf = pxy + scap;
  int *chunk_ctr = reinterpret_cast<int *>(&H->pad0);
  const unsigned pxy_s = (unsigned)__cvta_generic_to_shared(pxy);
  const unsigned pzf_s = (unsigned)__cvta_generic_to_shared(pzf);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int S = tile_rows_cached(hdrs, tile, H);
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {
    const int nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k, s = base + k;
        const double *p = reinterpret_cast<const double *>(xt + src);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(pxy_s + (unsigned)s * 16u), "l"(p)
                     : "memory");
        cp_async8(&pzf[s].x, p + 2);
        cp_async8(&pzf[s].y, fp + src);
      }
    }
    if (tid == 0) {
      pxy[S] = make_double2(TILE2_FAR, TILE2_FAR);
      pzf[S] = make_double2(TILE2_FAR, 0.0);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  const int ni = H->ni, ibase = tile_ibase[tile], W = maxslots >> 3;
  const double cutsq = P.cutforcesq, rdr = P.rdr, nscale = -F.scale;
  const int nr = P.nr;
  double evdwl = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(chunk_ctr, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk * 32 >= ni) break;
    const int ti = chunk * 32 + lane;
    if (ti >= ni) continue;
    const int g = ibase + ti;
    const uint4 *lp = list + g;
    const int li = iloc[g];
    const int ntot = min((int)tnum[g], maxslots), nfar = min((int)tfar[g], ntot), nnear = ntot - nfar;
    const int gi = tgi[g];
    const double2 mxy = pxy[li], mzf = pzf[li];
    const double pix = mxy.x, piy = mxy.y, piz = mzf.x, fpi = mzf.y;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int imask = 0;
    if (NVE) {
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v0) : "l"(nv.vx + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v1) : "l"(nv.vy + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v2) : "l"(nv.vz + gi));
      asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(imask) : "l"(nv.mask + gi));
    }
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;

    struct P4 { double x, y, z, f; };
    auto ldrec = [&](unsigned e) -> P4 {
      P4 p;
      const unsigned j = (e & TILE_IDX) * 16u;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.x), "=d"(p.y) : "r"(pxy_s + j));
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.z), "=d"(p.f) : "r"(pzf_s + j));
      return p;
    };
    auto entry = [&](const uint4 &w, int k) -> unsigned {
      const unsigned v = k < 2 ? w.x : (k < 4 ? w.y : (k < 6 ? w.z : w.w));
      return (k & 1) ? (v >> 16) : (v & 0xffffu);
    };
    // fpair of one partner at squared distance rsq (pair_eam.cpp:262-297) and the pair energy phi
    auto pairf = [&](double rsq, double fpj, double &phi) -> double {
      double recip;
      const double r = sqrt_fast(rsq, recip);
      recip = fma(recip, fma(-r, recip, 1.0), recip);
      int m;
      double p;
      eam2_knot(r, rdr, nr, m, p);
      const double4 *c = reinterpret_cast<const double4 *>(F.frc8 + 4 * m);  // two LDG.256
      const double4 qa = ld_xt(c), qb = ld_xt(c + 1);
      const double rhop = rdr * fma(fma(3.0 * qa.x, p, 2.0 * qa.y), p, qa.z);  // rhoip == rhojp
      const double t = qb.x * p;
      const double z2 = fma(fma(t + qb.y, p, qb.z), p, qb.w);
      const double z2p = rdr * fma(fma(3.0, t, qb.y + qb.y), p, qb.z);
      phi = z2 * recip;
      const double phip = (z2p - phi) * recip;
      const double psip = fma(fpi + fpj, rhop, phip);
      return nscale * psip * recip;
    };
    auto tally = [&](unsigned e, double dx, double dy, double dz, double f, double phi) {
      const bool fwd = (e & TILE_FWD) != 0;
      evdwl += fwd ? phi * F.scale : 0.0;
      const double w = fwd ? f : 0.0;
      vir[0] = fma(dx * dx, w, vir[0]); vir[1] = fma(dy * dy, w, vir[1]);
      vir[2] = fma(dz * dz, w, vir[2]); vir[3] = fma(dx * dy, w, vir[3]);
      vir[4] = fma(dx * dz, w, vir[4]); vir[5] = fma(dy * dz, w, vir[5]);
    };
    // one entry, branch-free (NEAR words) or only when inside the cutoff (FAR words)
    auto body = [&](unsigned e, const P4 &pj, bool always) {
      const double dx = pix - pj.x, dy = piy - pj.y, dz = piz - pj.z;
      const double rsq = rsq_ref(dx, dy, dz);
      const bool in = rsq < cutsq;
      if (always || in) {
        double phi;
        const double fp0 = pairf(always ? fmin(rsq, cutsq) : rsq, pj.f, phi);
        const double f = in ? fp0 : 0.0;
        fxi = fma(dx, f, fxi);
        fyi = fma(dy, f, fyi);
        fzi = fma(dz, f, fzi);
        if (EV) tally(e, dx, dy, dz, f, in ? phi : 0.0);
      }
    };
    for (int k0 = 0; k0 < nnear; k0 += 8) {
      const uint4 c = __ldg(lp + (size_t)(k0 >> 3) * NI);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const unsigned ea = entry(c, k), eb = entry(c, k + 1);
        const P4 a = ldrec(ea), b = ldrec(eb);
        body(ea, a, true);
        body(eb, b, true);
      }
    }
    for (int k0 = 0; k0 < nfar; k0 += 8) {
      const uint4 c = __ldg(lp + (size_t)(W - 1 - (k0 >> 3)) * NI);
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const unsigned ea = entry(c, k), eb = entry(c, k + 1);
        const P4 a = ldrec(ea), b = ldrec(eb);
        body(ea, a, false);
        body(eb, b, false);
      }
    }
    if (NVE) {
      double px = pix, py = piy, pz = piz;
      if (imask & nv.groupbit) {
        const double dtfm = nv.dtf / nv.mass[1];
        const double ka = __dmul_rn(dtfm, fxi), kb = __dmul_rn(dtfm, fyi), kc = __dmul_rn(dtfm, fzi);
        double a = v0, b = v1, c = v2;
        a = __dadd_rn(__dadd_rn(a, ka), ka);  // final_integrate(n), then the half-kick of n+1
        b = __dadd_rn(__dadd_rn(b, kb), kb);
        c = __dadd_rn(__dadd_rn(c, kc), kc);
        nv.vx[gi] = a; nv.vy[gi] = b; nv.vz[gi] = c;
        px = __dadd_rn(px, __dmul_rn(nv.dtv, a));
        py = __dadd_rn(py, __dmul_rn(nv.dtv, b));
        pz = __dadd_rn(pz, __dmul_rn(nv.dtv, c));
      }
      nv.xt_out[gi] = make_double4(px, py, pz, type2d(1));
      if (nv.do_check) {  // Neighbor::check_distance for the next step's decide()
        const double dx = px - nv.xhx[gi], dy = py - nv.xhy[gi], dz = pz - nv.xhz[gi];
        if (rsq_ref(dx, dy, dz) > nv.triggersq) *nv.moved = 1;
      }
    } else {
      fx[gi] = fxi;
      fy[gi] = fyi;
      fz[gi] = fzi;
    }
  }
  if (EV) {
    double v[7] = {evdwl, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5]};
    __syncthreads();
    block_sum<7>(v, reinterpret_cast<double *>(pxy));
    if (tid == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ev[k], v[k]);
  }
}
