// kernels_tile.cuh -- bin-tile kernels: the Verlet list and the pair loops with the
// neighbourhood of a tile of bins staged in shared memory.
//
// Why: the flat kernels (kernels_pair.cuh) gather every neighbour record through L1TEX and
// scatter every Newton partner force with a global RED; ncu shows them pinned at ~1 L1TEX
// wavefront per cycle per SM (profiles/r01d_ncu_full_k_pair_lj*.txt), not at DRAM or the FP
// pipes.  Here one CTA owns a tile of tx*ty*tz bins.  Atoms are bin-sorted, so the atoms of the
// tile and of its +-s bin halo are (ty+2s)*(tz+2s) runs of consecutive records (an owned run and a
// ghost run per x-row), copied once per CTA with cp.async (LDGSTS) into SoA x[],y[],z[],type[] in
// shared memory.  Neighbour reads are then three LDS.64, and list entries are 16-bit indices into
// the staged tile (half the list traffic of int32 indices).
// (A first version staged the 32-byte AoS records with cp.async.bulk: correct, but a 32-byte
// stride maps every record onto 4 of the 8 16-byte bank groups and ncu counted 232.9 M conflict
// wavefronts on 66 M ideal ones -- profiles/r01e_ncu_full_k_tile_lj_aos.txt.  SoA doubles give
// 16 distinct bank pairs for 16 consecutive atoms.)
//
// The list: for owned atom i every partner within the neighbour cutoff is stored once in i's row.
// An entry carries two flags:
//   FWD   the pair (i,j) is a member of the reference's half/Newton-on list of atom i
//         (NPairBin<1,1,0,0,1>::build, npair_bin.cpp:52-253: same bin -> j after i, ghosts by
//         the (z,y,x) rule; other bins -> upper half stencil).  The FWD entries ARE that list:
//         b200_get_neighbor_list exports exactly them and the parity tests compare them
//         bit-for-bit with the reference's pair set.
//   GHOST j is a ghost atom (only FWD entries can be ghosts).
// Entries without FWD are the transposed copy of an owned-owned pair (the member (j,i) of j's
// half list).  With them atom i accumulates ALL of its pair forces in registers and stores f_i
// once: no atomics between owned atoms.  Only a FWD|GHOST entry scatters (RED.ADD.F64) onto
// the ghost, which the reverse halo returns to its owner exactly as the reference does with
// Newton on.  Energy is tallied on FWD entries only, so every pair counts once; the virial
// uses the f.x form over owned+ghost atoms as before (pair.cpp:1809-1825).
// rsq uses the reference's operation order (rsq_ref) in the build and in the FP64-staged pair
// kernels, so list membership and cutoff decisions are those of the CPU path; the fixed-point
// mixed kernel decides in FP32 and re-takes every decision within 2e-6 of the cutoff in FP64.
// Kernels: k_tile_count / k_tile_split (per-tile sizes, interior|boundary order for the
// halo/compute overlap), k_tile_build (list), k_tile_export (test hook), k_tile_lj (FP64 and
// FP64-staged mixed), k_tile_lj_fx (fixed-point staged mixed), k_tile_eam_rho / k_tile_eam_force
// (behind B200_LIST=tile).  No tensor cores: nothing here is a dense contraction.
#pragma once
#include "common.cuh"
#include "kernels_pair.cuh"
#include "kernels_pair_mixed.cuh"

#define TILE_MAXROWS 160  // staged x-rows of one tile: (ty+2sy)*(tz+2sz)
#define TILE_MAXRUNS 64   // x-rows of the tile itself: ty*tz
#define FST_MAXROWS 49    // (dy,dz) rows of the full stencil: (2sy+1)*(2sz+1), s <= 3
#define TILE_FWD 0x8000u
#define TILE_GHOST 0x4000u
#define TILE_IDX 0x3fffu
#define TILE_MAXSTAGE 16384
#define TILE_NOATOM 0xffffu
#ifndef TILE_MINB
#define TILE_MINB 2  // CTAs per SM the pair kernels are register-budgeted for
#endif

struct TileGeom {
  int t[3];     // tile size in bins
  int nt[3];    // tiles per dimension
  int ilo[3];   // first local bin (per dim) that can hold an owned atom
  int nib[3];   // number of such bins per dim (the first and last may also hold ghosts)
  int s[3];     // stencil half width in bins (NStencil::sx,sy,sz, nstencil.cpp:203-237)
  int mbin[3];  // local bin grid
  int ntiles, srow_y, srow_z, sbx;  // staged rows along y and z, staged bins per row
  // fixed-point staging of the mixed-precision lj kernel: q = rn((x - origin) * fxscale),
  // origin = lower corner of the tile's staged bins minus fxpad
  double bin0[3], bsize[3], fxpad, fxscale;
};

// all (dy,dz) rows of the stencil, both halves: bins with bin_distance < cutneighmaxsq
struct FullStencil {
  int nrows;
  signed char dy[FST_MAXROWS], dz[FST_MAXROWS], dxlo[FST_MAXROWS], dxhi[FST_MAXROWS];
};

struct TileHdr {
  unsigned long long pad0;
  int S, ni, nrows, nruns;
  int rowbase[TILE_MAXROWS + 1];  // staged index of the first atom of each row
  int row_o0[TILE_MAXROWS], row_no[TILE_MAXROWS], row_g0[TILE_MAXROWS], row_ng[TILE_MAXROWS];
  int runpre[TILE_MAXRUNS + 1];   // exclusive prefix of the owned-atom counts of the tile's rows
  int run_o0[TILE_MAXRUNS];       // global index of the first owned atom of each of them
};
#define TILE_HDR_BYTES ((sizeof(TileHdr) + 127) / 128 * 128)

__host__ __device__ __forceinline__ size_t tile_smem_bytes(int scap, int rows, int sbx, bool build,
                                                          bool with_fp) {
  size_t b = TILE_HDR_BYTES + (size_t)scap * (3 * sizeof(double) + 2 * sizeof(int));
  if (with_fp) b += (size_t)scap * sizeof(double);
  if (build) b += (size_t)2 * rows * (sbx + 1) * sizeof(unsigned short);
  return (b + 127) / 128 * 128;
}

// the staged tile in shared memory
struct TileS {
  double *x, *y, *z, *fp;
  int *type, *gmap;
  unsigned short *tail;  // build only: bin tables
};
__device__ __forceinline__ TileS tile_carve(unsigned char *tsm, int scap, bool with_fp) {
  TileS T;
  T.x = reinterpret_cast<double *>(tsm + TILE_HDR_BYTES);
  T.y = T.x + scap;
  T.z = T.y + scap;
  T.fp = T.z + scap;
  T.type = reinterpret_cast<int *>(T.fp + (with_fp ? scap : 0));
  T.gmap = T.type + scap;
  T.tail = reinterpret_cast<unsigned short *>(T.gmap + scap);
  return T;
}

struct TilePos {
  int tx0, ty0, tz0;
};

__device__ __forceinline__ TilePos tile_pos(const TileGeom &G, int tile) {
  TilePos p;
  p.tx0 = G.ilo[0] + (tile % G.nt[0]) * G.t[0];
  p.ty0 = G.ilo[1] + ((tile / G.nt[0]) % G.nt[1]) * G.t[1];
  p.tz0 = G.ilo[2] + (tile / (G.nt[0] * G.nt[1])) * G.t[2];
  return p;
}

// x-extent [xa,xb) of the staged bins of a row, clamped to the local grid
__device__ __forceinline__ void tile_xrange(const TileGeom &G, const TilePos &P, int &xa, int &xb) {
  xa = max(P.tx0 - G.s[0], 0);
  xb = min(P.tx0 + G.t[0] + G.s[0], G.mbin[0]);
  if (xb < xa) xb = xa;
}

// Row tables of a tile: which runs of atom records make up the staged neighbourhood.
// Returns the staged atom count S (uniform over the CTA).  Ends with a __syncthreads.
__device__ __forceinline__ int tile_rows(const TileGeom &G, const TilePos &P,
                                         const int *__restrict__ ostart,
                                         const int *__restrict__ gstart, TileHdr *H) {
  const int tid = threadIdx.x, bd = blockDim.x, lane = tid & 31;
  const int nrows = G.srow_y * G.srow_z, nruns = G.t[1] * G.t[2];
  int xa, xb;
  tile_xrange(G, P, xa, xb);
  for (int r = tid; r < nrows; r += bd) {
    const int y = P.ty0 - G.s[1] + r % G.srow_y, z = P.tz0 - G.s[2] + r / G.srow_y;
    int o0 = 0, no = 0, g0 = 0, ng = 0;
    if (y >= 0 && y < G.mbin[1] && z >= 0 && z < G.mbin[2] && xb > xa) {
      const int b = (z * G.mbin[1] + y) * G.mbin[0];
      o0 = ostart[b + xa];
      no = ostart[b + xb] - o0;
      g0 = gstart[b + xa];
      ng = gstart[b + xb] - g0;
    }
    H->row_o0[r] = o0;
    H->row_no[r] = no;
    H->row_g0[r] = g0;
    H->row_ng[r] = ng;
  }
  // the tile's own rows: owned atoms of bins [tx0, tx0+tx) (bins past the sub-domain hold none)
  const int oxa = min(P.tx0, G.mbin[0]), oxb = min(P.tx0 + G.t[0], G.mbin[0]);
  for (int q = tid; q < nruns; q += bd) {
    const int y = P.ty0 + q % G.t[1], z = P.tz0 + q / G.t[1];
    int o0 = 0, no = 0;
    if (y < G.mbin[1] && z < G.mbin[2]) {
      const int b = (z * G.mbin[1] + y) * G.mbin[0];
      o0 = ostart[b + oxa];
      no = ostart[b + oxb] - o0;
    }
    H->run_o0[q] = o0;
    H->runpre[q + 1] = no;  // count for now, prefix below
  }
  if (tid == 0) {
    H->nrows = nrows;
    H->nruns = nruns;
  }
  __syncthreads();
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      const int r = base + lane;
      const int c = r < nrows ? H->row_no[r] + H->row_ng[r] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (r < nrows) H->rowbase[r] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      H->rowbase[nrows] = carry;
      H->S = carry;
    }
    carry = 0;
    for (int base = 0; base < nruns; base += 32) {
      const int q = base + lane;
      const int c = q < nruns ? H->runpre[q + 1] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      __syncwarp();
      if (q < nruns) H->runpre[q + 1] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      H->runpre[0] = 0;
      H->ni = carry;
    }
  }
  __syncthreads();
  return H->S;
}

// The row tables of a tile depend only on the bin starts, i.e. they change at a rebuild and not
// in between: k_tile_count stores every tile's finished header (TILE_HDR_BYTES each) and the
// per-step kernels fetch it with cp.async instead of recomputing it (two dependent rounds of
// global loads, two scans and ~1000 instructions per CTA and launch).  Ends with a __syncthreads.
__device__ __forceinline__ int tile_rows_cached(const unsigned char *__restrict__ hdrs, int tile,
                                                TileHdr *H) {
  const uint4 *src = reinterpret_cast<const uint4 *>(hdrs + (size_t)tile * TILE_HDR_BYTES);
  uint4 *dst = reinterpret_cast<uint4 *>(H);
  for (int k = threadIdx.x; k < (int)(TILE_HDR_BYTES / 16); k += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + k)),
                 "l"(src + k)
                 : "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  return H->S;
}

__device__ __forceinline__ double3 tile_pos3(const TileS &T, int s) {
  return make_double3(T.x[s], T.y[s], T.z[s]);
}

// cp.async (LDGSTS): global -> shared without a register round trip, so a warp keeps the copies
// of all its rows in flight at once
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
               "l"(src)
               : "memory");
}

// Stage the tile: one warp per run of records, {x,y,z,type} of each record copied into the SoA
// arrays with cp.async, plus the staged->global index map.  Precondition: tile_rows() done,
// S <= scap.  Ends with a __syncthreads after the copies have landed.
template <bool WITH_FP>
__device__ __forceinline__ void tile_stage(int nlocal, const double4 *__restrict__ xt,
                                           const double *__restrict__ fp, const TileHdr *H,
                                           const TileS &T) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nrows = H->nrows;
  for (int r = warp; r < nrows; r += nwarp) {
    const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
    const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
    for (int k = lane; k < n; k += 32) {
      const int src = k < no ? o0 + k : g0 + k, s = base + k;
      const double *p = reinterpret_cast<const double *>(xt + src);
      cp_async8(T.x + s, p);
      cp_async8(T.y + s, p + 1);
      cp_async8(T.z + s, p + 2);
      cp_async4(T.type + s, p + 3);  // low word of the int64 type field (little endian)
      if (WITH_FP) cp_async8(T.fp + s, fp + src);
      T.gmap[s] = src;
    }
  }
  if (threadIdx.x == 0) {
    // slot S: the dummy atom that the padding entries of a list row name (k_tile_build); its
    // record must be readable (the branch-free bodies evaluate it before discarding it)
    const int S = H->S;
    T.x[S] = T.y[S] = T.z[S] = 1.0e10;
    if (WITH_FP) T.fp[S] = 0.0;
    T.type[S] = 1;
    T.gmap[S] = 0;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
}

// thread ti of a tile -> staged index of its owned atom (and the atom's global index)
__device__ __forceinline__ int tile_own_atom(const TileGeom &G, const TileHdr *H, int ti, int &gi) {
  int lo = 0, hi = H->nruns;  // find run q with runpre[q] <= ti < runpre[q+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (H->runpre[mid] <= ti) lo = mid; else hi = mid;
  }
  const int q = lo, k = ti - H->runpre[q];
  const int row = (q % G.t[1] + G.s[1]) + (q / G.t[1] + G.s[2]) * G.srow_y;
  gi = H->run_o0[q] + k;
  return H->rowbase[row] + (gi - H->row_o0[row]);
}


// Walk the n entries of list row g: 8 entries per 16-byte word, next word prefetched.
// body(e, valid) must be branch-free (entries past n are zero padding -> valid = false): the
// eight bodies of a word then interleave in the instruction stream.  ghost(e) runs only for
// FWD|GHOST entries, behind one test per word.
template <int ILP, bool SCATTER, class Body, class Ghost>
__device__ __forceinline__ void tile_walk(const uint4 *__restrict__ list, int g, int n, int NI,
                                          Body &&body, Ghost &&ghost) {
  const uint4 *lp = list + g;
  uint4 q = n > 0 ? __ldg(lp) : make_uint4(0, 0, 0, 0);
  for (int k0 = 0; k0 < n; k0 += 8) {
    const uint4 c = q;
    if (k0 + 8 < n) q = __ldg(lp + (size_t)((k0 >> 3) + 1) * NI);
    // ILP entries at a time (2 in FP64, 4 with FP32 pair math): enough independent chains to
    // cover the pipe latency without the register cost of all eight.  The empty asm statements
    // keep the compiler from hoisting the next step's shared-memory reads above this step.
    auto step2 = [&](unsigned w, int kb) {
      body(w & 0xffffu, kb < n);
      body(w >> 16, kb + 1 < n);
      if (SCATTER && (w & (TILE_GHOST | (TILE_GHOST << 16)))) {
        if ((w & TILE_GHOST) && kb < n) ghost(w & 0xffffu);
        if ((w & (TILE_GHOST << 16)) && kb + 1 < n) ghost(w >> 16);
      }
    };
    auto step4 = [&](unsigned w0, unsigned w1, int kb) {
      body(w0 & 0xffffu, kb < n);
      body(w0 >> 16, kb + 1 < n);
      body(w1 & 0xffffu, kb + 2 < n);
      body(w1 >> 16, kb + 3 < n);
      if (SCATTER && ((w0 | w1) & (TILE_GHOST | (TILE_GHOST << 16)))) {
        if ((w0 & TILE_GHOST) && kb < n) ghost(w0 & 0xffffu);
        if ((w0 & (TILE_GHOST << 16)) && kb + 1 < n) ghost(w0 >> 16);
        if ((w1 & TILE_GHOST) && kb + 2 < n) ghost(w1 & 0xffffu);
        if ((w1 & (TILE_GHOST << 16)) && kb + 3 < n) ghost(w1 >> 16);
      }
    };
    if (ILP == 4) {
      step4(c.x, c.y, k0);
      asm volatile("" ::: "memory");
      step4(c.z, c.w, k0 + 4);
    } else {
      step2(c.x, k0);
      asm volatile("" ::: "memory");
      step2(c.y, k0 + 2);
      asm volatile("" ::: "memory");
      step2(c.z, k0 + 4);
      asm volatile("" ::: "memory");
      step2(c.w, k0 + 6);
    }
    asm volatile("" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------------
// per-tile counts: owned atoms (padded to a warp) and staged atoms; sizes the list and the
// shared-memory request.  tflags: [0] max staged, [1] max owned per tile, [4] owned total.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tile_count(TileGeom G, const int *__restrict__ ostart,
                                                    const int *__restrict__ gstart,
                                                    int *__restrict__ tile_ibase,
                                                    int *__restrict__ tile_bflag,
                                                    int *__restrict__ tflags, int reverse_halo,
                                                    unsigned char *__restrict__ hdrs) {
  __shared__ __align__(16) unsigned char hraw[TILE_HDR_BYTES];
  TileHdr &H = *reinterpret_cast<TileHdr *>(hraw);
  const TilePos P = tile_pos(G, blockIdx.x);
  const int S = tile_rows(G, P, ostart, gstart, &H);
  // "boundary" tile: it stages a ghost (must wait for the forward halo) or -- only when the pair
  // style scatters onto ghosts and a reverse halo follows (eam on tiles) -- its staged bins reach
  // the first/last owned bin of a dimension: then it may own atoms within the ghost cutoff of a
  // face, whose forces the reverse halo adds to (2 bins >= cutghost, nbin_standard.cpp:82-214).
  // Every other tile touches owned atoms only and can run beside the halo.
  int ghosts = 0;
  for (int r = threadIdx.x; r < H.nrows; r += blockDim.x) ghosts |= H.row_ng[r] > 0;
  const int t0[3] = {P.tx0, P.ty0, P.tz0};
#pragma unroll
  for (int d = 0; d < 3; d++)
    if (reverse_halo)
      ghosts |= (t0[d] - G.s[d] <= G.ilo[d]) || (t0[d] + G.t[d] - 1 + G.s[d] >= G.ilo[d] + G.nib[d] - 1);
  ghosts = __syncthreads_or(ghosts);
  {  // the finished header, for the per-step kernels (tile_rows_cached)
    if (threadIdx.x == 0) H.pad0 = 0;
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(hdrs + (size_t)blockIdx.x * TILE_HDR_BYTES);
    const uint4 *src = reinterpret_cast<const uint4 *>(hraw);
    for (int k = threadIdx.x; k < (int)(TILE_HDR_BYTES / 16); k += blockDim.x) dst[k] = src[k];
  }
  if (threadIdx.x == 0) {
    tile_bflag[blockIdx.x] = ghosts ? 1 : 0;
    const int ni = H.ni;
    tile_ibase[blockIdx.x] = (ni + 31) / 32 * 32;
    atomicMax(&tflags[0], S);
    atomicMax(&tflags[1], ni);
    atomicAdd(&tflags[4], ni);
  }
}

// tile ids ordered interior first, boundary last (both ascending, i.e. still in spatial order);
// bpos = exclusive scan of the boundary flags, bpos[ntiles] = number of boundary tiles
__global__ void __launch_bounds__(256) k_tile_split(int ntiles, const int *__restrict__ bflag,
                                                    const int *__restrict__ bpos,
                                                    int *__restrict__ ids) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const int nint = ntiles - bpos[ntiles];
  if (bflag[t]) ids[nint + bpos[t]] = t;
  else ids[t - bpos[t]] = t;
}

// ---------------------------------------------------------------------------------------
// List build.  One CTA per tile, one thread per owned atom; candidates are read from the staged
// tile in the reference's bin order.  Eight 16-bit entries are packed into one 16-byte store:
// entry n of list row g lives in word (n/8)*NI + g (uint4), lane-contiguous for a warp.
// tflags: [2] max entries per atom, [3] max FWD entries per atom, [5] tile overflow.
// ---------------------------------------------------------------------------------------
// FULLGHOST: a row also holds the ghost partners that are NOT members of atom i's half list
// (ghosts of the lower half stencil, of the bins left of the own bin, own-bin ghosts below i in
// the (z,y,x) order), flagged GHOST without FWD.  Every pair that crosses the sub-domain
// boundary is then evaluated from both sides -- on this rank for the owned atom, on the ghost's
// owner for its own atom -- exactly like owned-owned pairs are inside a sub-domain: the pair
// kernel stores complete forces for its owned atoms and there is neither a Newton scatter onto
// ghosts nor a reverse halo (lj/cut on tiles).  The FWD entries are unchanged: still exactly the
// reference's half/Newton-on list.  Without FULLGHOST (eam on tiles) only FWD ghosts are stored
// and the pair kernels scatter onto them.
// SPLIT (eam on tiles, kernels_eam2.cuh): a row is filled from both ends.  Partners whose distance
// at build time is <= sqrt(splitsq) (force cutoff + a margin) are the NEAR entries, words 0, 1, ...
// as before; the others are the FAR entries, stored from the last word of the row downwards
// (far entry f lives in word maxslots/8 - 1 - f/8), tfar[g] counts them, tnum[g] stays the total.
// The pair kernels evaluate the near entries unconditionally and the far ones behind a cutoff
// test that almost never fires, so that a warp does not run the expensive eam pair function for
// the ~30 % of the stored partners that sit in the skin.  The SET of entries is unchanged.
// TRI (triclinic box, npair_bin.cpp:133-155): the half list is not cut out by stencil halves and
// coordinates but by the local order (owned j after owned i: here the global index, any
// antisymmetric order stores each owned pair once) and, for a ghost j, by the parity of
// itag + jtag; a ghost image of i itself by the (z,y,x) comparison with tolerance `tri_delta`.
// The rule is antisymmetric between the two owners of a boundary pair, so FULLGHOST rows flag
// the pair FWD on exactly one side, as in the orthogonal case.
// EXTRA: the rarely used list rules (newton off membership, neigh_modify exclude group) are compiled
// into separate instantiations so that the default build carries none of their tests.
template <bool ONETYPE, bool FULLGHOST, bool SPLIT = false, bool TRI = false, bool EXTRA = false>
__global__ void __launch_bounds__(512) k_tile_build(
    TileGeom G, FullStencil F, int nlocal, const double4 *__restrict__ xt,
    const int *__restrict__ ostart, const int *__restrict__ gstart,
    const int *__restrict__ atombin, const int *__restrict__ tile_ibase, int NI, int maxslots,
    double cut1, const double *__restrict__ cutneighsq, int ntypes,
    unsigned short *__restrict__ iloc, unsigned short *__restrict__ tnum, int *__restrict__ tgi,
    uint4 *__restrict__ list, int *__restrict__ numneigh_half, int scap, int *__restrict__ tflags,
    double splitsq = 0.0, unsigned short *__restrict__ tfar = nullptr,
    const int *__restrict__ tag = nullptr, double tri_delta = 0.0, int newtoff = 0,
    ExGroups ex = ExGroups{0, {0}, {0}}, const int *__restrict__ mask = nullptr) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  const TileS T = tile_carve(tsm, scap, false);
  unsigned short *sbo = T.tail;
  const int ncol = G.sbx + 1;
  unsigned short *sbg = sbo + (size_t)G.srow_y * G.srow_z * ncol;

  const int tile = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S + 1 > scap || S + 1 > TILE_MAXSTAGE) {  // + 1: the dummy atom that padding entries name
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  // staged index of the first owned / first ghost atom of every staged bin (+ end sentinel)
  {
    int xa, xb;
    tile_xrange(G, P, xa, xb);
    const int xs = P.tx0 - G.s[0], nrows = H->nrows;
    for (int e = tid; e < nrows * ncol; e += bd) {
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
  }
  tile_stage<false>(nlocal, xt, nullptr, H, T);

  const int ni = H->ni, ibase = tile_ibase[tile], nipad = (ni + 31) / 32 * 32;
  const int n1 = ntypes + 1;
  const int xs = P.tx0 - G.s[0], ys = P.ty0 - G.s[1], zs = P.tz0 - G.s[2];
  int wmax = 0, wmaxf = 0;
  for (int ti = tid; ti < nipad; ti += bd) {
    const int g = ibase + ti;
    int n = 0, nf = 0;
    if (ti < ni) {
      int gi;
      const int li = tile_own_atom(G, H, ti, gi);
      const double3 pi = tile_pos3(T, li);
      const double *cut_i = ONETYPE ? nullptr : cutneighsq + (size_t)T.type[li] * n1;
      const int b = atombin[gi];
      const int bx = b % G.mbin[0], by = (b / G.mbin[0]) % G.mbin[1], bz = b / (G.mbin[0] * G.mbin[1]);
      unsigned long long qlo = 0, qhi = 0, flo = 0, fhi = 0;
      int nfar = 0;  // SPLIT: far entries so far (n counts the near ones until the row is done)
      const int W = maxslots >> 3;
      auto push = [&](int s, unsigned flags, bool isfar) {
        const unsigned long long e = (unsigned)s | flags;
        // members of the reference's list: the FWD entries -- or, with newton off
        // (NPairBin<HALF,!NEWTON>, npair_bin.cpp:126-131: "stores own/ghost pairs on both procs"),
        // the FWD owned partners and EVERY ghost partner
        nf += (EXTRA && newtoff && (flags & TILE_GHOST)) ? 1 : (flags >> 15);
        if (SPLIT && isfar) {
          flo = (flo >> 16) | (fhi << 48);
          fhi = (fhi >> 16) | (e << 48);
          nfar++;
          if ((nfar & 7) == 0 && (nfar >> 3) <= W)
            list[(size_t)(W - (nfar >> 3)) * NI + g] = make_uint4((unsigned)flo, (unsigned)(flo >> 32),
                                                                (unsigned)fhi, (unsigned)(fhi >> 32));
          return;
        }
        qlo = (qlo >> 16) | (qhi << 48);
        qhi = (qhi >> 16) | (e << 48);
        n++;
        if ((n & 7) == 0 && n <= maxslots)
          list[(size_t)((n >> 3) - 1) * NI + g] = make_uint4((unsigned)qlo, (unsigned)(qlo >> 32),
                                                           (unsigned)qhi, (unsigned)(qhi >> 32));
      };
      auto dist = [&](int s) -> double {  // rsq in the reference's operation order
        const double3 pj = tile_pos3(T, s);
        return rsq_ref(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
      };
      const int mi = (EXTRA && ex.n) ? mask[gi] : 0;
      auto cutof = [&](int s) -> double {  // the reference's test: rsq <= cutneighsq[itype][jtype]
        // (neigh_modify exclude group: an excluded partner fails the test at every distance)
        if (EXTRA && ex.n && ex_group(ex, mi, mask[T.gmap[s]])) return -1.0;
        return ONETYPE ? cut1 : __ldg(cut_i + T.type[s]);
      };
      auto test = [&](int s, unsigned flags) {
        const double rsq = dist(s);
        if (rsq <= cutof(s)) push(s, flags, rsq > splitsq);
      };
      // a run of consecutive staged atoms with the same flags: two distance tests in flight
      // (four were measured slower: 2.71 vs 2.57 ms per build at 4 M atoms)
      // (they are independent; the stores of `push` are not on their path)
      auto run = [&](int lo, int hi, unsigned flags) {
        int s = lo;
        for (; s + 1 < hi; s += 2) {
          const double ra = dist(s), rb = dist(s + 1);
          if (ra <= cutof(s)) push(s, flags, ra > splitsq);
          if (rb <= cutof(s + 1)) push(s + 1, flags, rb > splitsq);
        }
        if (s < hi) test(s, flags);
      };
      for (int r = 0; r < F.nrows; r++) {
        const int dy = F.dy[r], dz = F.dz[r];
        const int srow = (by + dy - ys) + (bz + dz - zs) * G.srow_y;
        const unsigned short *bo = sbo + srow * ncol, *bg = sbg + srow * ncol;
        const int ca = bx + F.dxlo[r] - xs, cb = bx + F.dxhi[r] + 1 - xs;
        if (TRI) {
          const int itag = tag[gi];
          // owned partners: a staged row is a run of consecutive global indices, so "after i in
          // the local order" splits the candidate run at one point (and leaves out i itself)
          {
            const int lo = bo[ca], hi = bo[cb];
            const int sstar = H->rowbase[srow] + (gi + 1 - H->row_o0[srow]);  // first with index > gi
            const int sfwd = min(max(sstar, lo), hi);
            int sback = sfwd;
            if (sback > lo && sback - 1 == li) sback--;
            run(lo, sback, 0u);
            run(sfwd, hi, TILE_FWD);
          }
          for (int s = bg[ca]; s < bg[cb]; s++) {
            const int jtag = tag[T.gmap[s]];
            bool member = true;
            if (itag > jtag) member = ((itag + jtag) % 2) != 0;
            else if (itag < jtag) member = ((itag + jtag) % 2) != 1;
            else {
              const double3 pj = tile_pos3(T, s);
              if (fabs(pj.z - pi.z) > tri_delta) member = !(pj.z < pi.z);
              else if (fabs(pj.y - pi.y) > tri_delta) member = !(pj.y < pi.y);
              else member = !(pj.x < pi.x);
            }
            if (member) test(s, TILE_FWD | TILE_GHOST);
            else if (FULLGHOST) test(s, TILE_GHOST);
          }
          continue;
        }
        if (dz > 0 || (dz == 0 && dy > 0)) {  // upper half stencil: members of i's half list
          run(bo[ca], bo[cb], TILE_FWD);
          run(bg[ca], bg[cb], TILE_FWD | TILE_GHOST);
        } else if (dz < 0 || dy < 0) {        // lower half: owned j holds (j,i) in ITS half list
          run(bo[ca], bo[cb], 0u);
          if (FULLGHOST) run(bg[ca], bg[cb], TILE_GHOST);
        } else {
          // row (0,0): bins left of own bin -> transposed; own bin -> by list position;
          // right -> members.  Owned atoms of a row are staged in index order, so that is s > li.
          for (int s = bo[ca]; s < bo[cb]; s++)
            if (s != li) test(s, s > li ? TILE_FWD : 0u);
          const int c0 = bx - xs;
          if (FULLGHOST) run(bg[ca], bg[c0], TILE_GHOST);
          for (int s = bg[c0]; s < bg[c0 + 1]; s++) {  // own-bin ghosts: npair_bin.cpp:156-171
            const double3 pj = tile_pos3(T, s);
            bool member = true;
            if (pj.z < pi.z) member = false;
            else if (pj.z == pi.z) {
              if (pj.y < pi.y) member = false;
              else if (pj.y == pi.y && pj.x < pi.x) member = false;
            }
            if (member) test(s, TILE_FWD | TILE_GHOST);
            else if (FULLGHOST) test(s, TILE_GHOST);
          }
          run(bg[c0 + 1], bg[cb], TILE_FWD | TILE_GHOST);
        }
      }
      if ((n & 7) && (n >> 3) < (maxslots >> 3)) {
        // the rest of the last word names staged slot S: a dummy atom the pair kernels park far
        // away, so they need no per-entry bounds test (kernels_tile2.cuh)
        for (int k = n & 7; k < 8; k++) {
          qlo = (qlo >> 16) | (qhi << 48);
          qhi = (qhi >> 16) | ((unsigned long long)S << 48);
        }
        list[(size_t)(n >> 3) * NI + g] = make_uint4((unsigned)qlo, (unsigned)(qlo >> 32),
                                                   (unsigned)qhi, (unsigned)(qhi >> 32));
      }
      if (SPLIT) {
        if ((nfar & 7) && (nfar >> 3) < W) {
          for (int k = nfar & 7; k < 8; k++) {
            flo = (flo >> 16) | (fhi << 48);
            fhi = (fhi >> 16) | ((unsigned long long)S << 48);
          }
          list[(size_t)(W - 1 - (nfar >> 3)) * NI + g] = make_uint4((unsigned)flo, (unsigned)(flo >> 32),
                                                                 (unsigned)fhi, (unsigned)(fhi >> 32));
        }
        tfar[g] = (unsigned short)min(nfar, 65535);
        // slots the row needs: whole words from both ends (the host grows maxslots to this)
        const int need = (((n + 7) >> 3) + ((nfar + 7) >> 3)) << 3;
        n += nfar;
        wmax = max(wmax, need);
      }
      iloc[g] = (unsigned short)li;
      tgi[g] = gi;
      numneigh_half[gi] = nf;
    } else {
      iloc[g] = (unsigned short)TILE_NOATOM;
      if (SPLIT) tfar[g] = 0;
    }
    tnum[g] = (unsigned short)min(n, 65535);
    wmax = max(wmax, n);
    wmaxf = max(wmaxf, nf);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    wmaxf = max(wmaxf, __shfl_xor_sync(0xffffffffu, wmaxf, o));
  }
  if ((tid & 31) == 0 && wmax > 0) {
    atomicMax(&tflags[2], wmax);
    atomicMax(&tflags[3], wmaxf);
  }
}

// half list (FWD entries) as CSR over the owned atoms in global indices (test hook)
__global__ void __launch_bounds__(512) k_tile_export(
    TileGeom G, int nlocal, const int *__restrict__ ostart, const int *__restrict__ gstart,
    const int *__restrict__ tile_ibase, int NI, int maxslots, const unsigned short *__restrict__ tnum,
    const uint4 *__restrict__ list, const long long *__restrict__ first, int *__restrict__ flat,
    int scap, const unsigned short *__restrict__ tfar = nullptr, int newtoff = 0) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  int *gmap = reinterpret_cast<int *>(tsm + TILE_HDR_BYTES);
  const int tile = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S > scap) return;
  for (int r = tid >> 5; r < H->nrows; r += bd >> 5) {
    const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
    const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
    for (int k = tid & 31; k < n; k += 32) gmap[base + k] = k < no ? o0 + k : g0 + k;
  }
  __syncthreads();
  const int ni = H->ni, ibase = tile_ibase[tile];
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    int gi;
    tile_own_atom(G, H, ti, gi);
    const int n = min((int)tnum[g], maxslots);
    long long o = first[gi];
    // (SPLIT rows: the last tfar[g] entries are stored from the end of the row downwards)
    const int nfar = tfar ? min((int)tfar[g], n) : 0, W = maxslots >> 3;
    for (int kk = 0; kk < n; kk++) {
      const bool far = kk >= n - nfar;
      const int k = far ? kk - (n - nfar) : kk;
      const uint4 q = list[(size_t)(far ? W - 1 - (k >> 3) : (k >> 3)) * NI + g];
      const unsigned w = (k & 4) ? ((k & 2) ? q.w : q.z) : ((k & 2) ? q.y : q.x);
      const unsigned e = (w >> ((k & 1) * 16)) & 0xffffu;
      if ((e & TILE_FWD) || (newtoff && (e & TILE_GHOST) && (e & TILE_IDX) != (unsigned)S))
        flat[o++] = gmap[e & TILE_IDX];
    }
  }
}

// ---------------------------------------------------------------------------------------
// lj/cut over a tile.  PairLJCut::compute (pair_lj_cut.cpp:71-141); MIXED evaluates the pair
// function in FP32 (del, rsq and the cutoff test stay FP64) like k_pair_lj_mixed.
// ---------------------------------------------------------------------------------------
template <bool EV, bool ONETYPE, bool MIXED>
__global__ void __launch_bounds__(352, TILE_MINB) k_tile_lj(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const uint4 *__restrict__ list, double *__restrict__ fx, double *__restrict__ fy,
    double *__restrict__ fz, LJOne one, LJOneF onef, const double *__restrict__ tab,
    const float *__restrict__ tabf, int ntypes, double *__restrict__ ev, int scap,
    int *__restrict__ tflags,
    const int *__restrict__ tile_ids) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  const TileS T = tile_carve(tsm, scap, false);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S + 1 > scap) {  // cannot happen between rebuilds (the rows are those the build staged)
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  tile_stage<false>(nlocal, xt, nullptr, H, T);
  const int ni = H->ni, ibase = tile_ibase[tile];
  const int n1 = ntypes + 1, n2 = n1 * n1;
  double evdwl = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    const double3 pi = tile_pos3(T, li);
    const int gi = T.gmap[li];
    const int itype = T.type[li];
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    // pair force of one entry (times del gives the force on i); 0 outside the cutoff
    auto pair = [&](unsigned e, bool valid, double &delx, double &dely, double &delz, double &fp64,
                    float &fp32, double &epair) {
      const int j = e & TILE_IDX;
      const double3 pj = tile_pos3(T, j);
      delx = pi.x - pj.x; dely = pi.y - pj.y; delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      int tij = 0;
      double cutsq = one.cutsq;
      if (!ONETYPE) {
        tij = itype * n1 + T.type[j];
        cutsq = __ldg(tab + tij);
      }
      const bool in = valid && rsq < cutsq;
      if (MIXED) {
        const float lj1 = ONETYPE ? onef.lj1 : __ldg(tabf + tij);
        const float lj2 = ONETYPE ? onef.lj2 : __ldg(tabf + n2 + tij);
        const float r2inv = rcp_f((float)rsq);
        const float r6inv = r2inv * r2inv * r2inv;
        fp32 = in ? r6inv * (lj1 * r6inv - lj2) * r2inv : 0.0f;
        if (EV) {
          const float lj3 = ONETYPE ? onef.lj3 : __ldg(tabf + 2 * n2 + tij);
          const float lj4 = ONETYPE ? onef.lj4 : __ldg(tabf + 3 * n2 + tij);
          const float off = ONETYPE ? onef.offset : __ldg(tabf + 4 * n2 + tij);
          epair = (in && (e & TILE_FWD)) ? (double)(r6inv * (lj3 * r6inv - lj4) - off) : 0.0;
        }
      } else {
        const double lj1 = ONETYPE ? one.lj1 : __ldg(tab + n2 + tij);
        const double lj2 = ONETYPE ? one.lj2 : __ldg(tab + 2 * n2 + tij);
        const double r2inv = rcp_nr(rsq);
        const double r6inv = r2inv * r2inv * r2inv;
        fp64 = in ? r6inv * (lj1 * r6inv - lj2) * r2inv : 0.0;
        if (EV) {
          const double lj3 = ONETYPE ? one.lj3 : __ldg(tab + 3 * n2 + tij);
          const double lj4 = ONETYPE ? one.lj4 : __ldg(tab + 4 * n2 + tij);
          const double off = ONETYPE ? one.offset : __ldg(tab + 5 * n2 + tij);
          epair = (in && (e & TILE_FWD)) ? r6inv * (lj3 * r6inv - lj4) - off : 0.0;
        }
      }
    };
    // (the list is FULLGHOST: ghost partners are ordinary entries, nothing is scattered)
    tile_walk<MIXED ? 4 : 2, false>(
        list, g, n, NI,
        [&](unsigned e, bool valid) {
          double delx, dely, delz, f64 = 0.0, ep = 0.0;
          float f32 = 0.0f;
          pair(e, valid, delx, dely, delz, f64, f32, ep);
          if (MIXED) {  // FP32 pair math, FP64 accumulation
            fxi += (double)((float)delx * f32); fyi += (double)((float)dely * f32);
            fzi += (double)((float)delz * f32);
          } else {
            fxi += delx * f64; fyi += dely * f64; fzi += delz * f64;
          }
          if (EV) {
            // energy and virial are tallied once per pair, on the half-list (FWD) copy
            // (Pair::ev_tally, pair.cpp:1087-1182: v = del (x) del * fpair)
            evdwl += ep;
            const double w = (e & TILE_FWD) ? (MIXED ? (double)f32 : f64) : 0.0;
            vir[0] += delx * delx * w; vir[1] += dely * dely * w; vir[2] += delz * delz * w;
            vir[3] += delx * dely * w; vir[4] += delx * delz * w; vir[5] += dely * delz * w;
          }
        },
        [&](unsigned) {});
    fx[gi] = fxi;
    fy[gi] = fyi;
    fz[gi] = fzi;
  }
  if (EV) {
    double v[7] = {evdwl, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5]};
    __syncthreads();
    block_sum<7>(v, T.x);
    if (tid == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ev[k], v[k]);
  }
}

// ---------------------------------------------------------------------------------------
// eam over a tile, PairEAM::compute (pair_eam.cpp:124-327).  Phase 1: rho_i from every entry of
// the row (stored, not accumulated); a FWD|GHOST entry also adds this atom's density onto the
// ghost (RED), which the rho reverse halo returns to its owner (pair_eam.cpp:215,1625-1646).
// ---------------------------------------------------------------------------------------
template <bool MIXED>
__global__ void __launch_bounds__(352, TILE_MINB) k_tile_eam_rho(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const uint4 *__restrict__ list, EAMParams P, EAMParamsF F, double *__restrict__ rho, int scap,
    int *__restrict__ tflags,
    const int *__restrict__ tile_ids) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  const TileS T = tile_carve(tsm, scap, false);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos Tp = tile_pos(G, tile);
  const int S = tile_rows(G, Tp, ostart, gstart, H);
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  tile_stage<false>(nlocal, xt, nullptr, H, T);
  const int ni = H->ni, ibase = tile_ibase[tile], n1 = P.ntypes + 1;
  const bool onetype = P.ntypes == 1;
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    const double3 pi = tile_pos3(T, li);
    const int itype = T.type[li];
    double rhoi = 0.0;
    // density of type `from` felt at distance sqrt(rsq) by type `to`; 0 outside the cutoff
    auto dens = [&](double rsq, bool valid, int from, int to) -> double {
      const bool in = valid && rsq < P.cutforcesq;
      const int tr = onetype ? P.type2rhor[n1 + 1] : P.type2rhor[from * n1 + to];
      if (MIXED) {
        float p = sqrtf((float)rsq) * F.rdr + 1.0f;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= (float)m;
        p = fminf(p, 1.0f);
        // knot = {c0,c1,c2,c3 | c4,c5,c6,pad}: the value cubic is c3..c6
        const float4 *kc = reinterpret_cast<const float4 *>(F.rhor + ((size_t)tr * (P.nr + 1) + m) * 8);
        const float4 c0 = __ldg(kc), c1 = __ldg(kc + 1);
        return in ? (double)(((c0.w * p + c1.x) * p + c1.y) * p + c1.z) : 0.0;
      } else {
        double rinv;
        double p = sqrt_nr(fmax(rsq, 1.0e-300), rinv) * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        const double *c = P.rhor + ((size_t)tr * (P.nr + 1) + m) * 7;
        return in ? ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6) : 0.0;
      }
    };
    tile_walk<MIXED ? 4 : 2, true>(
        list, g, n, NI,
        [&](unsigned e, bool valid) {
          const int j = e & TILE_IDX;
          const double3 pj = tile_pos3(T, j);
          const double rsq = rsq_ref(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
          rhoi += dens(rsq, valid, onetype ? 1 : T.type[j], itype);
        },
        [&](unsigned e) {
          const int j = e & TILE_IDX;
          const double3 pj = tile_pos3(T, j);
          const double rsq = rsq_ref(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
          if (rsq < P.cutforcesq) atomicAdd(&rho[T.gmap[j]], dens(rsq, true, itype, T.type[j]));
        });
    rho[T.gmap[li]] = rhoi;
  }
}

// Phase 3 (pair_eam.cpp:233-314).  fp of the staged atoms rides along in shared memory.
template <bool EV, bool MIXED>
__global__ void __launch_bounds__(352, TILE_MINB) k_tile_eam_force(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const uint4 *__restrict__ list, EAMParams P, EAMParamsF F, const double *__restrict__ fp,
    double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
    double *__restrict__ ev, int scap, int *__restrict__ tflags,
    const int *__restrict__ tile_ids) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  const TileS T = tile_carve(tsm, scap, true);
  const double *sfp = T.fp;
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos Tp = tile_pos(G, tile);
  const int S = tile_rows(G, Tp, ostart, gstart, H);
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  tile_stage<true>(nlocal, xt, fp, H, T);
  const int ni = H->ni, ibase = tile_ibase[tile], n1 = P.ntypes + 1;
  const bool onetype = P.ntypes == 1;
  const int t11 = n1 + 1;
  double evdwl = 0.0;
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    const double3 pi = tile_pos3(T, li);
    const int gi = T.gmap[li];
    const int itype = T.type[li];
    const double fpi = sfp[li];
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    float gxi = 0.0f, gyi = 0.0f, gzi = 0.0f;
    // fpair of one entry (times del gives the force on i) and its pair energy; 0 outside the cutoff
    auto pair = [&](unsigned e, bool valid, double &delx, double &dely, double &delz, double &f64,
                    float &f32, double &epair) {
      const int j = e & TILE_IDX;
      const double3 pj = tile_pos3(T, j);
      delx = pi.x - pj.x; dely = pi.y - pj.y; delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      const bool in = valid && rsq < P.cutforcesq;
      const int jtype = onetype ? 1 : T.type[j];
      const int tt = onetype ? t11 : itype * n1 + jtype;
      const int tij = P.type2rhor[tt], tji = onetype ? tij : P.type2rhor[jtype * n1 + itype];
      const int tz = P.type2z2r[tt];
      const double fpj = sfp[j];
      if (MIXED) {
        const float r = sqrtf((float)rsq);
        float p = r * F.rdr + 1.0f;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= (float)m;
        p = fminf(p, 1.0f);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(F.rhor + ((size_t)tij * (P.nr + 1) + m) * 8));
        const float rhoip = (a.x * p + a.y) * p + a.z;
        float rhojp = rhoip;
        if (tji != tij) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(F.rhor + ((size_t)tji * (P.nr + 1) + m) * 8));
          rhojp = (b.x * p + b.y) * p + b.z;
        }
        const float4 *zc = reinterpret_cast<const float4 *>(F.z2r + ((size_t)tz * (P.nr + 1) + m) * 8);
        const float4 z0 = __ldg(zc), z1 = __ldg(zc + 1);
        const float z2p = (z0.x * p + z0.y) * p + z0.z;
        const float z2 = ((z0.w * p + z1.x) * p + z1.y) * p + z1.z;
        const float recip = rcp_f(r);
        const float phi = z2 * recip;
        const float phip = z2p * recip - phi * recip;
        const float psip = (float)fpi * rhojp + (float)fpj * rhoip + phip;
        const float sc = (float)P.scale[tt];
        f32 = in ? -sc * psip * recip : 0.0f;
        if (EV) epair = (in && (e & TILE_FWD)) ? (double)(sc * phi) : 0.0;
      } else {
        double recip;
        const double r = sqrt_nr(fmax(rsq, 1.0e-300), recip);
        recip = fma(recip, fma(-r, recip, 1.0), recip);  // one Newton step: 1/r to <= 1 ulp
        double p = r * P.rdr + 1.0;
        int m = (int)p;
        m = min(m, P.nr - 1);
        p -= m;
        p = fmin(p, 1.0);
        const double *c = P.rhor + ((size_t)tij * (P.nr + 1) + m) * 7;
        const double rhoip = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        double rhojp = rhoip;
        if (tji != tij) {
          c = P.rhor + ((size_t)tji * (P.nr + 1) + m) * 7;
          rhojp = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        }
        c = P.z2r + ((size_t)tz * (P.nr + 1) + m) * 7;
        const double z2p = (__ldg(c) * p + __ldg(c + 1)) * p + __ldg(c + 2);
        const double z2 = ((__ldg(c + 3) * p + __ldg(c + 4)) * p + __ldg(c + 5)) * p + __ldg(c + 6);
        const double phi = z2 * recip;
        const double phip = z2p * recip - phi * recip;
        const double psip = fpi * rhojp + fpj * rhoip + phip;
        const double sc = P.scale[tt];
        f64 = in ? -sc * psip * recip : 0.0;
        if (EV) epair = (in && (e & TILE_FWD)) ? sc * phi : 0.0;
      }
    };
    tile_walk<MIXED ? 4 : 2, true>(
        list, g, n, NI,
        [&](unsigned e, bool valid) {
          double delx, dely, delz, f64 = 0.0, ep = 0.0;
          float f32 = 0.0f;
          pair(e, valid, delx, dely, delz, f64, f32, ep);
          if (MIXED) {
            gxi += (float)delx * f32; gyi += (float)dely * f32; gzi += (float)delz * f32;
          } else {
            fxi += delx * f64; fyi += dely * f64; fzi += delz * f64;
          }
          if (EV) evdwl += ep;
        },
        [&](unsigned e) {
          double delx, dely, delz, f64 = 0.0, ep = 0.0;
          float f32 = 0.0f;
          pair(e, true, delx, dely, delz, f64, f32, ep);
          const int gj = T.gmap[e & TILE_IDX];
          if (MIXED) {
            atomicAdd(&fx[gj], -(double)((float)delx * f32));
            atomicAdd(&fy[gj], -(double)((float)dely * f32));
            atomicAdd(&fz[gj], -(double)((float)delz * f32));
          } else {
            atomicAdd(&fx[gj], -(delx * f64));
            atomicAdd(&fy[gj], -(dely * f64));
            atomicAdd(&fz[gj], -(delz * f64));
          }
        });
    if (MIXED) {
      fxi = (double)gxi; fyi = (double)gyi; fzi = (double)gzi;
    }
    fx[gi] = fxi;
    fy[gi] = fyi;
    fz[gi] = fzi;
  }
  if (EV) {
    double v[1] = {evdwl};
    __syncthreads();
    block_sum<1>(v, T.x);
    if (tid == 0) atomicAdd(&ev[0], v[0]);
  }
}

// tile_walk for the fixed-point kernel: four branch-free bodies per step; a body returns true
// when its cutoff decision must be re-taken in FP64 (then `exact(e)` adds that entry's
// contribution; rare).  `fold()` runs once per list word (FP32 partial sums -> FP64).
template <class Body, class Exact, class Fold>
__device__ __forceinline__ void tile_walk_fx(const uint4 *__restrict__ list, int g, int n, int NI,
                                             Body &&body, Exact &&exact, Fold &&fold) {
  const uint4 *lp = list + g;
  uint4 q = n > 0 ? __ldg(lp) : make_uint4(0, 0, 0, 0);
  for (int k0 = 0; k0 < n; k0 += 8) {
    const uint4 c = q;
    if (k0 + 8 < n) q = __ldg(lp + (size_t)((k0 >> 3) + 1) * NI);
    auto step4 = [&](unsigned w0, unsigned w1, int kb) {
      const bool a0 = body(w0 & 0xffffu, kb < n), a1 = body(w0 >> 16, kb + 1 < n);
      const bool a2 = body(w1 & 0xffffu, kb + 2 < n), a3 = body(w1 >> 16, kb + 3 < n);
      if (a0 | a1 | a2 | a3) {
        if (a0) exact(w0 & 0xffffu);
        if (a1) exact(w0 >> 16);
        if (a2) exact(w1 & 0xffffu);
        if (a3) exact(w1 >> 16);
      }
    };
    step4(c.x, c.y, k0);
    asm volatile("" ::: "memory");
    step4(c.z, c.w, k0 + 4);
    fold();
    asm volatile("" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------------
// lj/cut over a tile, mixed precision with fixed-point staging.
// Positions are staged as int32 fixed point relative to the tile corner (resolution
// 1/fxscale ~ 1.5e-8 sigma for 8x8x4 tiles: finer than an FP32 coordinate, and differences of
// two staged coordinates are exact), 12 bytes per atom instead of 24: half the shared-memory
// wavefronts that bound the FP64 kernel, and no FP64 instruction in the common path.
// del = float(q_i - q_j) / fxscale, rsq and the pair function in FP32.  A cutoff decision
// within 2e-6 (relative) of the cutoff is re-taken in FP64 from the global positions with the
// reference's operation order, so the set of interacting pairs is still the CPU path's.
// Tolerances (BASELINE.json): forces <= 1e-5 relative, energy/pressure <= 1e-6.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t tile_smem_bytes_fx(int scap) {
  return (TILE_HDR_BYTES + (size_t)scap * 5 * sizeof(int) + 127) / 128 * 128;
}

#ifndef TILE_FX_MINB
#define TILE_FX_MINB 2
#endif
template <bool EV, bool ONETYPE>
__global__ void __launch_bounds__(352, TILE_FX_MINB) k_tile_lj_fx(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const uint4 *__restrict__ list, double *__restrict__ fx, double *__restrict__ fy,
    double *__restrict__ fz, LJOne one, LJOneF onef, const double *__restrict__ tab,
    const float *__restrict__ tabf, int ntypes, double *__restrict__ ev, int scap,
    int *__restrict__ tflags, const int *__restrict__ tile_ids) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  int *qx = reinterpret_cast<int *>(tsm + TILE_HDR_BYTES), *qy = qx + scap, *qz = qy + scap;
  int *stype = qz + scap, *gmap = stype + scap;
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {  // stage: one warp per run, LDG.256 of the record, fixed-point conversion, STS.32
    const double ox = G.bin0[0] + (P.tx0 - G.s[0]) * G.bsize[0] - G.fxpad,
                 oy = G.bin0[1] + (P.ty0 - G.s[1]) * G.bsize[1] - G.fxpad,
                 oz = G.bin0[2] + (P.tz0 - G.s[2]) * G.bsize[2] - G.fxpad;
    const int lane = tid & 31, warp = tid >> 5, nwarp = bd >> 5, nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k, s = base + k;
        const double4 p = ld_xt(xt + src);
        qx[s] = __double2int_rn((p.x - ox) * G.fxscale);
        qy[s] = __double2int_rn((p.y - oy) * G.fxscale);
        qz[s] = __double2int_rn((p.z - oz) * G.fxscale);
        stype[s] = d2type(p.w);
        gmap[s] = src;
      }
    }
    if (tid == 0) {  // the dummy atom of padding entries (never `valid`, but its type is read)
      qx[S] = qy[S] = qz[S] = 0;
      stype[S] = 1;
      gmap[S] = 0;
    }
    __syncthreads();
  }
  const int ni = H->ni, ibase = tile_ibase[tile];
  const int n1 = ntypes + 1, n2 = n1 * n1;
  const float inv = (float)(1.0 / G.fxscale);
  double evdwl = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int ti = tid; ti < ni; ti += bd) {
    const int g = ibase + ti;
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    const int xi = qx[li], yi = qy[li], zi = qz[li];
    const int gi = gmap[li];
    const int itype = stype[li];
    // FP32 pair math, FP64 accumulation: the FP32 partial sums of one list word (8 entries) are
    // added to FP64 accumulators, so the rounding of the sum does not grow with the list length
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    float gxi = 0.0f, gyi = 0.0f, gzi = 0.0f, ei = 0.0f;
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f, w4 = 0.0f, w5 = 0.0f;
    // FP32 pair function of entry e; `in` decided by the caller
    auto lj = [&](float rsq, int tij, bool in, bool fwd, float &fpair, float &epair) {
      const float lj1 = ONETYPE ? onef.lj1 : __ldg(tabf + tij);
      const float lj2 = ONETYPE ? onef.lj2 : __ldg(tabf + n2 + tij);
      const float r2inv = rcp_f(rsq);
      const float r6inv = r2inv * r2inv * r2inv;
      fpair = in ? r6inv * (lj1 * r6inv - lj2) * r2inv : 0.0f;
      if (EV) {
        const float lj3 = ONETYPE ? onef.lj3 : __ldg(tabf + 2 * n2 + tij);
        const float lj4 = ONETYPE ? onef.lj4 : __ldg(tabf + 3 * n2 + tij);
        const float off = ONETYPE ? onef.offset : __ldg(tabf + 4 * n2 + tij);
        epair = (in && fwd) ? r6inv * (lj3 * r6inv - lj4) - off : 0.0f;
      }
    };
    auto geom = [&](int j, float &delx, float &dely, float &delz, float &rsq, int &tij, float &cutsq) {
      delx = (float)(xi - qx[j]) * inv;
      dely = (float)(yi - qy[j]) * inv;
      delz = (float)(zi - qz[j]) * inv;
      rsq = delx * delx + dely * dely + delz * delz;
      tij = 0;
      cutsq = (float)one.cutsq;
      if (!ONETYPE) {
        tij = itype * n1 + stype[j];
        cutsq = (float)__ldg(tab + tij);
      }
    };
    // the reference's FP64 cutoff test, from the global positions (rare: |rsq - cutsq| tiny)
    auto exact_in = [&](int j, int tij) -> bool {
      const double4 a = ld_xt(xt + gi), b = ld_xt(xt + gmap[j]);
      return rsq_ref(a.x - b.x, a.y - b.y, a.z - b.z) < (ONETYPE ? one.cutsq : __ldg(tab + tij));
    };
    // energy and virial: once per pair, on the half-list (FWD) copy (Pair::ev_tally)
    auto tally = [&](unsigned e, float delx, float dely, float delz, float f32, float ep) {
      ei += ep;
      const float w = (e & TILE_FWD) ? f32 : 0.0f;
      w0 += delx * delx * w; w1 += dely * dely * w; w2 += delz * delz * w;
      w3 += delx * dely * w; w4 += delx * delz * w; w5 += dely * delz * w;
    };
    tile_walk_fx(
        list, g, n, NI,
        [&](unsigned e, bool valid) -> bool {  // fast path; returns "too close to call"
          float delx, dely, delz, rsq, cutsq, f32, ep = 0.0f;
          int tij;
          geom(e & TILE_IDX, delx, dely, delz, rsq, tij, cutsq);
          const bool amb = valid && fabsf(rsq - cutsq) < 2.0e-6f * cutsq;
          lj(rsq, tij, valid && !amb && rsq < cutsq, e & TILE_FWD, f32, ep);
          gxi += delx * f32; gyi += dely * f32; gzi += delz * f32;
          if (EV) tally(e, delx, dely, delz, f32, ep);
          return amb;
        },
        [&](unsigned e) {  // an ambiguous entry: decide in FP64, then the same FP32 pair function
          float delx, dely, delz, rsq, cutsq, f32, ep = 0.0f;
          int tij;
          geom(e & TILE_IDX, delx, dely, delz, rsq, tij, cutsq);
          lj(rsq, tij, exact_in(e & TILE_IDX, tij), e & TILE_FWD, f32, ep);
          gxi += delx * f32; gyi += dely * f32; gzi += delz * f32;
          if (EV) tally(e, delx, dely, delz, f32, ep);
        },
        [&]() {  // end of a list word: fold the FP32 partial sums into the FP64 accumulators
          fxi += (double)gxi; fyi += (double)gyi; fzi += (double)gzi;
          gxi = gyi = gzi = 0.0f;
        });
    fx[gi] = fxi;
    fy[gi] = fyi;
    fz[gi] = fzi;
    if (EV) {
      evdwl += (double)ei;  // <= ~60 terms per atom in FP32, FP64 across atoms
      vir[0] += (double)w0; vir[1] += (double)w1; vir[2] += (double)w2;
      vir[3] += (double)w3; vir[4] += (double)w4; vir[5] += (double)w5;
    }
  }
  if (EV) {
    double v[7] = {evdwl, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5]};
    __syncthreads();
    block_sum<7>(v, reinterpret_cast<double *>(tsm + TILE_HDR_BYTES));
    if (tid == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ev[k], v[k]);
  }
}
