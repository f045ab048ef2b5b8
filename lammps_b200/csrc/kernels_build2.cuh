// kernels_build2.cuh -- second-generation build of the bin-tile list (same list, same layout, same
// SET of entries per row as k_tile_build in kernels_tile.cuh; only the order inside a row differs).
//
// k_tile_build gives every owned atom a thread that walks its candidates one by one: lanes of a
// warp own atoms of ~14 different bins, so the candidate runs they walk have different lengths
// (the warp pays the longest), every accepted candidate runs the divergent `push`, and every
// scattered candidate costs three LDS.64 with bank conflicts -- ncu: 27 issue slots per distance
// test, 2.66 ms per build at 4 M atoms (profiles/r01k_ncu_full_k_tile_build.txt).
// Here a WARP owns a bin.  All owned atoms of a bin see the same candidates (the atoms of the
// stencil bins: ~50 runs of consecutive staged atoms), so
//   1. the warp expands the runs once per bin into a flat candidate list in shared memory
//      (one lane per run; 16-bit staged index + the FWD/GHOST flags the run implies),
//   2. for each owned atom of the bin the 32 lanes test 32 candidates at a time: consecutive
//      staged atoms -> conflict-free LDS, the atom's own position is a broadcast, no lane idles,
//   3. the accepted candidates are compacted with a ballot + popc into the atom's row, staged in
//      shared memory, and the finished row leaves as 16-byte words.
// The reference's membership rules are unchanged (npair_bin.cpp:52-253): upper half stencil and
// own-bin successors are FWD (the half/Newton-on list), own-bin ghosts by the (z,y,x) rule,
// everything else is the transposed copy (FULLGHOST: also the non-member ghosts), rsq by rsq_ref.
// No tensor cores: nothing here is a dense contraction.
#pragma once
#include "kernels_tile.cuh"

#define B2_WARPS 16
#ifndef B2_MINB
#define B2_MINB 1
#endif
#define B2_CAND 640  // candidates expanded at once per warp (a bin of the LJ melt sees ~230)

__host__ __device__ __forceinline__ size_t build2_smem_bytes(int scap, int rows, int sbx, bool with_type,
                                                            int maxslots) {
  size_t b = TILE_HDR_BYTES + (size_t)scap * 3 * sizeof(double);
  if (with_type) b += (size_t)scap * sizeof(int);
  b += (size_t)2 * rows * (sbx + 1) * sizeof(unsigned short);
  b = (b + 15) / 16 * 16;
  b += (size_t)B2_WARPS * (B2_CAND + maxslots) * sizeof(unsigned short);
  return (b + 127) / 128 * 128;
}

template <bool ONETYPE, bool FULLGHOST, bool SPLIT>
__global__ void __launch_bounds__(B2_WARPS * 32, B2_MINB) k_tile_build2(
    TileGeom G, FullStencil F, int nlocal, const double4 *__restrict__ xt,
    const int *__restrict__ ostart, const int *__restrict__ gstart,
    const int *__restrict__ tile_ibase, int NI, int maxslots, double cut1,
    const double *__restrict__ cutneighsq, int ntypes, unsigned short *__restrict__ iloc,
    unsigned short *__restrict__ tnum, int *__restrict__ tgi, uint4 *__restrict__ list,
    int *__restrict__ numneigh_half, int scap, int *__restrict__ tflags, double splitsq,
    unsigned short *__restrict__ tfar) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  // staged positions: {x,y} pairs, then z (as in k_tile_lj2), then the types
  double *pos = reinterpret_cast<double *>(tsm + TILE_HDR_BYTES);
  double *posz = pos + (size_t)2 * scap;
  int *stype = reinterpret_cast<int *>(pos + (size_t)3 * scap);
  const int ncol = G.sbx + 1, nrows_s = G.srow_y * G.srow_z;
  unsigned short *sbo = reinterpret_cast<unsigned short *>(stype + (ONETYPE ? 0 : scap));
  unsigned short *sbg = sbo + (size_t)nrows_s * ncol;
  unsigned short *wbuf = reinterpret_cast<unsigned short *>(
      (reinterpret_cast<size_t>(sbg + (size_t)nrows_s * ncol) + 15) / 16 * 16);
  int *bin_ctr = reinterpret_cast<int *>(&H->pad0);

  // the stencil rows are indexed per lane below: shared memory, not the (serialising) constant bank
  __shared__ signed char fdy[FST_MAXROWS], fdz[FST_MAXROWS], fxlo[FST_MAXROWS], fxhi[FST_MAXROWS];

  const int tile = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = bd >> 5;
  const TilePos P = tile_pos(G, tile);
  if (tid < F.nrows) {
    fdy[tid] = F.dy[tid];
    fdz[tid] = F.dz[tid];
    fxlo[tid] = F.dxlo[tid];
    fxhi[tid] = F.dxhi[tid];
  }
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S + 1 > scap || S + 1 > TILE_MAXSTAGE) {  // + 1: the dummy atom that padding entries name
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {  // staged index of the first owned / first ghost atom of every staged bin (+ end sentinel)
    int xa, xb;
    tile_xrange(G, P, xa, xb);
    const int xs = P.tx0 - G.s[0];
    for (int e = tid; e < nrows_s * ncol; e += bd) {
      const int r = e / ncol, c = e % ncol;
      const int y = P.ty0 - G.s[1] + r % G.srow_y, z = P.tz0 - G.s[2] + r / G.srow_y;
      int so = H->rowbase[r], sg = so + H->row_no[r];
      if (y >= 0 && y < G.mbin[1] && z >= 0 && z < G.mbin[2] && xb > xa) {
        const int b = (z * G.mbin[1] + y) * G.mbin[0];
        const int xc = min(max(xs + c, xa), xb);
        so += ostart[b + xc] - H->row_o0[r];
        sg += gstart[b + xc] - H->row_g0[r];
      }
      sbo[e] = (unsigned short)so;
      sbg[e] = (unsigned short)sg;
    }
    const unsigned pos_s = (unsigned)__cvta_generic_to_shared(pos);
    for (int r = warp; r < H->nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k, s = base + k;
        const double *p = reinterpret_cast<const double *>(xt + src);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(pos_s + (unsigned)s * 16u), "l"(p)
                     : "memory");
        cp_async8(posz + s, p + 2);
        if (!ONETYPE) cp_async4(stype + s, p + 3);
      }
    }
    if (tid == 0) *bin_ctr = 0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }

  unsigned short *cand = wbuf + (size_t)warp * (B2_CAND + maxslots);
  unsigned short *rowbuf = cand + B2_CAND;  // (B2_CAND and maxslots are multiples of 8: 16-byte aligned)
  const int ni = H->ni, ibase = tile_ibase[tile], nipad = (ni + 31) / 32 * 32;
  const int n1 = ntypes + 1, W = maxslots >> 3;
  const int xs = P.tx0 - G.s[0], ys = P.ty0 - G.s[1], zs = P.tz0 - G.s[2];
  const int nbin_tile = G.t[0] * G.t[1] * G.t[2];
  const unsigned lt = (1u << lane) - 1u;
  int wmax = 0, wmaxf = 0;
  // the runs of a bin: 2 per stencil row (owned, ghost) + the two right-hand pieces of row (0,0)
  const int nq = 2 * F.nrows + 2;
  int r00 = -1;
  for (int k = 0; k < F.nrows; k++)
    if (F.dy[k] == 0 && F.dz[k] == 0) r00 = k;

  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(bin_ctr, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= nbin_tile) break;
    const int lx = b % G.t[0], ly = (b / G.t[0]) % G.t[1], lz = b / (G.t[0] * G.t[1]);
    const int bx = P.tx0 + lx, by = P.ty0 + ly, bz = P.tz0 + lz;
    if (bx >= G.mbin[0] || by >= G.mbin[1] || bz >= G.mbin[2]) continue;
    const int row0 = (by - ys) + (bz - zs) * G.srow_y, c0 = bx - xs;
    const int io0 = sbo[row0 * ncol + c0], io1 = sbo[row0 * ncol + c0 + 1];
    if (io1 == io0) continue;  // no owned atom in this bin
    const int ig0 = sbg[row0 * ncol + c0], ig1 = sbg[row0 * ncol + c0 + 1];
    const int nso = io1 - io0, nsg = ig1 - ig0;  // own-bin candidates, first in the list

    // run q of this bin: [lo, hi) staged atoms with the flags the stencil position implies
    auto run_of = [&](int q, int &lo, int &hi, unsigned &flags) {
      lo = hi = 0;
      flags = 0;
      const bool right = q >= 2 * F.nrows;  // right-hand pieces of row (0,0)
      int r = q >> 1;
      const int ghost = q & 1;
      if (right) {
        r = r00;
        if (r < 0) return;
      }
      const int dy = fdy[r], dz = fdz[r];
      const int srow = (by + dy - ys) + (bz + dz - zs) * G.srow_y;
      const unsigned short *tb = (ghost ? sbg : sbo) + srow * ncol;
      const int ca = bx + fxlo[r] - xs, cb = bx + fxhi[r] + 1 - xs;
      if (dz > 0 || (dz == 0 && dy > 0)) {  // upper half stencil: members of i's half list
        lo = tb[ca]; hi = tb[cb];
        flags = ghost ? (TILE_FWD | TILE_GHOST) : TILE_FWD;
      } else if (dz < 0 || dy < 0) {         // lower half: the partner holds the pair
        if (ghost && !FULLGHOST) return;
        lo = tb[ca]; hi = tb[cb];
        flags = ghost ? TILE_GHOST : 0u;
      } else if (!right) {                   // row (0,0), bins left of the own bin
        if (ghost && !FULLGHOST) return;
        lo = tb[ca]; hi = tb[c0];
        flags = ghost ? TILE_GHOST : 0u;
      } else {                               // row (0,0), bins right of the own bin
        lo = tb[c0 + 1]; hi = tb[cb];
        flags = ghost ? (TILE_FWD | TILE_GHOST) : TILE_FWD;
      }
      if (hi < lo) hi = lo;
    };

    // total number of candidates: decides between one expansion for all atoms of the bin and the
    // (rare: very dense bins) segmented walk
    int mylen = 0;
    for (int q = lane; q < nq; q += 32) {
      int lo, hi;
      unsigned fl;
      run_of(q, lo, hi, fl);
      mylen += hi - lo;
    }
    int incl = mylen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int nrun_cand = __shfl_sync(0xffffffffu, incl, 31);
    const int ncand_all = nso + nsg + nrun_cand;
    const bool single = ncand_all <= B2_CAND;

    // expand candidates [seg0, seg0 + B2_CAND) of the bin's flat candidate sequence into cand[]
    // (sequence: own-bin owned, own-bin ghosts, then the runs lane by lane); returns the count
    auto expand = [&](int seg0) -> int {
      __syncwarp();
      const int seg1 = seg0 + B2_CAND;
      for (int c = lane; c < nso + nsg; c += 32)
        if (c >= seg0 && c < seg1) cand[c - seg0] = (unsigned short)(c < nso ? io0 + c : ig0 + (c - nso));
      int off = nso + nsg + incl - mylen;
      for (int q = lane; q < nq; q += 32) {
        int lo, hi;
        unsigned fl;
        run_of(q, lo, hi, fl);
        for (int s = lo; s < hi; s++, off++)
          if (off >= seg0 && off < seg1) cand[off - seg0] = (unsigned short)(s | fl);
      }
      __syncwarp();
      return min(ncand_all, seg1) - seg0;
    };

    int nc = 0;
    if (single) nc = expand(0);
    for (int li = io0; li < io1; li++) {
      const int gi = H->row_o0[row0] + (li - H->rowbase[row0]);
      const int qrun = ly + lz * G.t[1];
      const int ti = H->runpre[qrun] + (gi - H->run_o0[qrun]);
      const int g = ibase + ti;
      const double pix = pos[2 * li], piy = pos[2 * li + 1], piz = posz[li];
      const double *cut_i = ONETYPE ? nullptr : cutneighsq + (size_t)stype[li] * n1;
      // the row starts as all-padding: entries name the dummy atom S
      for (int k = lane; k < W; k += 32) {
        const unsigned ss = (unsigned)S | ((unsigned)S << 16);
        reinterpret_cast<uint4 *>(rowbuf)[k] = make_uint4(ss, ss, ss, ss);
      }
      __syncwarp();
      int n = 0, nfar = 0, nf = 0;
      for (int seg0 = 0; seg0 < ncand_all; seg0 += B2_CAND) {
        if (!single) nc = expand(seg0);
        for (int k0 = 0; k0 < nc; k0 += 32) {
          const int c = k0 + lane;
          bool valid = c < nc;
          unsigned e = valid ? cand[c] : (unsigned)S;
          const int s = e & TILE_IDX;
          double2 pxy;
          double pz;
          pxy = reinterpret_cast<const double2 *>(pos)[s];
          pz = posz[s];
          const int cg = seg0 + c;  // position in the bin's candidate sequence
          if (cg < nso + nsg) {     // own bin: flags depend on the atom (npair_bin.cpp:156-171)
            if (cg < nso) {
              if (s == li) valid = false;
              e = (unsigned)s | (s > li ? TILE_FWD : 0u);
            } else {
              bool member = true;
              if (pz < piz) member = false;
              else if (pz == piz) {
                if (pxy.y < piy) member = false;
                else if (pxy.y == piy && pxy.x < pix) member = false;
              }
              if (member) e = (unsigned)s | TILE_FWD | TILE_GHOST;
              else if (FULLGHOST) e = (unsigned)s | TILE_GHOST;
              else valid = false;
            }
          }
          const double rsq = rsq_ref(pix - pxy.x, piy - pxy.y, piz - pz);
          const double cut = ONETYPE ? cut1 : (valid ? __ldg(cut_i + stype[s]) : 0.0);  // (slot S holds no type)
          const bool ok = valid && rsq <= cut;
          const bool isfar = SPLIT && rsq > splitsq;
          const unsigned mk = __ballot_sync(0xffffffffu, ok);
          const unsigned mf = SPLIT ? __ballot_sync(0xffffffffu, ok && isfar) : 0u;
          const unsigned mn = mk & ~mf;
          nf += __popc(__ballot_sync(0xffffffffu, ok && (e & TILE_FWD)));
          if (ok) {
            // far entry f lives in word W-1-f/8 at element f%8, as k_tile_build stores it
            const int f = nfar + __popc(mf & lt);
            const int p = (SPLIT && isfar) ? ((W - 1 - (f >> 3)) << 3) + (f & 7) : n + __popc(mn & lt);
            if (p >= 0 && p < maxslots) rowbuf[p] = (unsigned short)e;
          }
          n += __popc(mn);
          nfar += __popc(mf);
        }
      }
      __syncwarp();
      // the finished row: near words from the front, far words from the back
      const int nw = (n + 7) >> 3, nfw = (nfar + 7) >> 3;
      for (int k = lane; k < W; k += 32)
        if (k < nw || k >= W - nfw) list[(size_t)k * NI + g] = reinterpret_cast<const uint4 *>(rowbuf)[k];
      if (lane == 0) {
        iloc[g] = (unsigned short)li;
        tgi[g] = gi;
        numneigh_half[gi] = nf;
        tnum[g] = (unsigned short)min(n + nfar, 65535);
        if (SPLIT) tfar[g] = (unsigned short)min(nfar, 65535);
      }
      wmax = max(wmax, SPLIT ? (nw + nfw) << 3 : n);
      wmaxf = max(wmaxf, nf);
      __syncwarp();
    }
  }
  // rows of the padding threads of the last warp-sized chunk
  for (int ti = ni + tid; ti < nipad; ti += bd) {
    iloc[ibase + ti] = (unsigned short)TILE_NOATOM;
    tnum[ibase + ti] = 0;
    if (SPLIT) tfar[ibase + ti] = 0;
  }
  if (lane == 0 && wmax > 0) {
    atomicMax(&tflags[2], wmax);
    atomicMax(&tflags[3], wmaxf);
  }
}
