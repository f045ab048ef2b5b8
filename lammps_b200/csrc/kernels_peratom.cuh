// kernels_peratom.cuh -- per-atom energy and virial of the pair style: Pair::ev_tally's
// eatom[] / vatom[][6] (pair.cpp:1087-1182) as compute pe/atom, stress/atom and the reference's own
// pair-style unit test (unittest/force-styles/test_pair_style.cpp:143) read them.
//
// With Newton on the reference gives each atom of a pair half of the pair energy and half of the
// pair virial del (x) del * fpair, ghosts included, and the computes return the ghost shares to
// their owners with a reverse communication.  These are output-step quantities: a separate pass
// over the list in force, not a variant of the per-step force kernels.
//   k_peratom_tile   FULLGHOST rows (lj/cut on tiles, eam on the second-generation tile kernels):
//                    a row holds EVERY partner of its owned atom, so the thread sums half of every
//                    pair term itself -- no scatter, no reverse halo; SPLIT rows (tfar != null)
//                    are walked near words first, far words from the end.
//   k_peratom_flat   half lists (flat int32 list, or the tile list without ghost partners is not
//                    used here): half to i in registers, half to j with RED.F64 onto owned+ghost
//                    arrays; the engine then runs the reverse halo on the seven arrays.
// STYLE 1 = lj/cut (PairLJCut::compute, pair_lj_cut.cpp:104-134, reference operation order),
// STYLE 2 = eam (PairEAM::compute, pair_eam.cpp:262-304: phi and fpair from the general tables;
// the embedding energy F(rho_i) of pair_eam.cpp:219-231 is added by k_peratom_embed).
// No tensor cores: nothing here is a dense contraction.
#pragma once
#include "kernels_tile.cuh"

struct PerAtomOut {
  double *e;     // [n] or null
  double *v[6];  // xx, yy, zz, xy, xz, yz; each [n] or null (all or none)
};

// pair energy and fpair of one interacting pair; returns false outside the cutoff
template <int STYLE>
__device__ __forceinline__ bool peratom_pair(double rsq, int itype, int jtype, int n1,
                                             const double *__restrict__ ljtab, const EAMParams &P,
                                             double fpi, double fpj, double &epair, double &fpair) {
  if (STYLE == 1) {
    const int n2 = n1 * n1, tij = itype * n1 + jtype;
    if (!(rsq < ljtab[tij])) return false;
    const double r2inv = 1.0 / rsq;
    const double r6inv = r2inv * r2inv * r2inv;
    const double forcelj = r6inv * (ljtab[n2 + tij] * r6inv - ljtab[2 * n2 + tij]);
    fpair = forcelj * r2inv;
    epair = r6inv * (ljtab[3 * n2 + tij] * r6inv - ljtab[4 * n2 + tij]) - ljtab[5 * n2 + tij];
    return true;
  } else {
    if (!(rsq < P.cutforcesq)) return false;
    const double r = sqrt(rsq);
    double p = r * P.rdr + 1.0;
    int m = (int)p;
    m = min(m, P.nr - 1);
    p -= m;
    p = fmin(p, 1.0);
    const int tij = P.type2rhor[itype * n1 + jtype], tji = P.type2rhor[jtype * n1 + itype];
    const double *c = P.rhor + ((size_t)tij * (P.nr + 1) + m) * 7;
    const double rhoip = (c[0] * p + c[1]) * p + c[2];
    c = P.rhor + ((size_t)tji * (P.nr + 1) + m) * 7;
    const double rhojp = (c[0] * p + c[1]) * p + c[2];
    c = P.z2r + ((size_t)P.type2z2r[itype * n1 + jtype] * (P.nr + 1) + m) * 7;
    const double z2p = (c[0] * p + c[1]) * p + c[2];
    const double z2 = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
    const double recip = 1.0 / r;
    const double phi = z2 * recip;
    const double phip = z2p * recip - phi * recip;
    const double psip = fpi * rhojp + fpj * rhoip + phip;
    const double sc = P.scale[itype * n1 + jtype];
    fpair = -sc * psip * recip;
    epair = sc * phi;
    return true;
  }
}

template <int STYLE>
__global__ void __launch_bounds__(352) k_peratom_tile(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const double *__restrict__ fp,
    const int *__restrict__ ostart, const int *__restrict__ gstart, const int *__restrict__ tile_ibase,
    int NI, int maxslots, const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const unsigned short *__restrict__ tfar, const uint4 *__restrict__ list,
    const double *__restrict__ ljtab, EAMParams P, int ntypes, PerAtomOut out, int scap) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  const TileS T = tile_carve(tsm, scap, true);
  const int tile = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos Tp = tile_pos(G, tile);
  const int S = tile_rows(G, Tp, ostart, gstart, H);
  if (S + 1 > scap) return;  // (the build and the force kernels report this; nothing to add here)
  if (STYLE == 2) tile_stage<true>(nlocal, xt, fp, H, T);
  else tile_stage<false>(nlocal, xt, nullptr, H, T);
  const int ni = H->ni, ibase = tile_ibase[tile], n1 = ntypes + 1, W = maxslots >> 3;
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    const int nfar = tfar ? min((int)tfar[g], n) : 0, nnear = n - nfar;
    const double3 pi = tile_pos3(T, li);
    const int gi = T.gmap[li], itype = T.type[li];
    const double fpi = STYLE == 2 ? T.fp[li] : 0.0;
    double e = 0.0, v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int kk = 0; kk < n; kk++) {
      const bool far = kk >= nnear;
      const int k = far ? kk - nnear : kk;
      const uint4 q = list[(size_t)(far ? W - 1 - (k >> 3) : (k >> 3)) * NI + g];
      const unsigned w = (k & 4) ? ((k & 2) ? q.w : q.z) : ((k & 2) ? q.y : q.x);
      const int j = (w >> ((k & 1) * 16)) & TILE_IDX;
      if (j >= S) continue;  // padding entry
      const double3 pj = tile_pos3(T, j);
      const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      double ep, fpair;
      if (!peratom_pair<STYLE>(rsq_ref(dx, dy, dz), itype, T.type[j], n1, ljtab, P, fpi,
                               STYLE == 2 ? T.fp[j] : 0.0, ep, fpair))
        continue;
      e += 0.5 * ep;
      v[0] += 0.5 * dx * dx * fpair; v[1] += 0.5 * dy * dy * fpair; v[2] += 0.5 * dz * dz * fpair;
      v[3] += 0.5 * dx * dy * fpair; v[4] += 0.5 * dx * dz * fpair; v[5] += 0.5 * dy * dz * fpair;
    }
    if (out.e) out.e[gi] = e;
    if (out.v[0])
      for (int k = 0; k < 6; k++) out.v[k][gi] = v[k];
  }
}

template <int STYLE>
__global__ void __launch_bounds__(128) k_peratom_flat(int nlocal, int nstride, int T,
                                                      const double4 *__restrict__ xt,
                                                      const double *__restrict__ fp,
                                                      const int *__restrict__ numneigh,
                                                      const int *__restrict__ neigh,
                                                      const double *__restrict__ ljtab, EAMParams P,
                                                      int ntypes, PerAtomOut out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const double4 pi = xt[i];
  const int itype = d2type(pi.w), n1 = ntypes + 1, jnum = numneigh[i];
  const double fpi = STYLE == 2 ? fp[i] : 0.0;
  double e = 0.0, v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int n = 0; n < jnum; n++) {
    // slot n of atom i in the transposed list of T lanes per atom (kernels_neigh.cuh, k_build_half)
    const int j = neigh[(size_t)(n / T) * nstride * T + (size_t)i * T + (n % T)] & NEIGHMASK;
    const double4 pj = ld_xt(xt + j);
    const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
    double ep, fpair;
    if (!peratom_pair<STYLE>(rsq_ref(dx, dy, dz), itype, d2type(pj.w), n1, ljtab, P, fpi,
                             STYLE == 2 ? fp[j] : 0.0, ep, fpair))
      continue;
    const double h[7] = {0.5 * ep, 0.5 * dx * dx * fpair, 0.5 * dy * dy * fpair, 0.5 * dz * dz * fpair,
                         0.5 * dx * dy * fpair, 0.5 * dx * dz * fpair, 0.5 * dy * dz * fpair};
    e += h[0];
    for (int k = 0; k < 6; k++) v[k] += h[1 + k];
    if (out.e) atomicAdd(&out.e[j], h[0]);
    if (out.v[0])
      for (int k = 0; k < 6; k++) atomicAdd(&out.v[k][j], h[1 + k]);
  }
  if (out.e) atomicAdd(&out.e[i], e);
  if (out.v[0])
    for (int k = 0; k < 6; k++) atomicAdd(&out.v[k][i], v[k]);
}

// eam: the embedding energy F(rho_i) * scale joins the atom's energy (pair_eam.cpp:219-231)
__global__ void __launch_bounds__(256) k_peratom_embed(int nlocal, const double4 *__restrict__ xt,
                                                       EAMParams P, const double *__restrict__ rho,
                                                       double *__restrict__ eatom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int itype = d2type(xt[i].w), n1 = P.ntypes + 1;
  const double r = rho[i];
  double p = r * P.rdrho + 1.0;
  int m = (int)p;
  m = max(1, min(m, P.nrho - 1));
  p -= m;
  p = fmin(p, 1.0);
  const double *c = P.frho + ((size_t)P.type2frho[itype] * (P.nrho + 1) + m) * 7;
  double phi = ((c[3] * p + c[4]) * p + c[5]) * p + c[6];
  if (r > P.rhomax) phi += ((c[0] * p + c[1]) * p + c[2]) * (r - P.rhomax);
  eatom[i] += phi * P.scale[itype * n1 + itype];
}
