// kernels_neigh.cuh -- rebuild-step kernels: periodic wrap, counting sort by bin, ghost
// (border) construction, and the half/Newton Verlet-list build.  All are HBM-bound
// streaming or gather kernels; no tensor-core work exists on this path.
#pragma once
#include "common.cuh"
#include "kernels_halo.cuh"
#include "kernels_pair.cuh"

// ---------------------------------------------------------------------------------------
// exclusive scan of int32 (bin counts -> bin starts), 3 launches, hand-written.
// ---------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int block_excl_scan_int(int v, int *smem /*[32]*/, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive prefix of warp totals
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int res = incl - v + smem[warp];
  *total = smem[32];
  __syncthreads();
  return res;
}

// pass 1: per-tile exclusive scan in place (out may alias in), tile totals to tilesum
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int *in, int *out, int n,
                                                             int *__restrict__ tilesum) {
  __shared__ int smem[33];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int pre = block_excl_scan_int(s, smem, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = pre;
    pre += v[k];
  }
  if (threadIdx.x == 0) tilesum[blockIdx.x] = total;
}

// pass 2: one block scans the tile totals (sequential chunks of blockDim.x with a carry)
__global__ void __launch_bounds__(1024) k_scan_sums(int *__restrict__ tilesum, int ntiles,
                                                    int *__restrict__ grand_total) {
  __shared__ int smem[33];
  int carry = 0;
  for (int b = 0; b < ntiles; b += blockDim.x) {
    int i = b + threadIdx.x;
    int v = i < ntiles ? tilesum[i] : 0, total;
    int pre = block_excl_scan_int(v, smem, &total);
    if (i < ntiles) tilesum[i] = pre + carry;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;  // also written to out[n] by pass 3
}

// pass 3: add tile offsets; element n receives the grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int *__restrict__ out, int n,
                                                           const int *__restrict__ tilesum,
                                                           const int *__restrict__ grand_total) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  const int off = tilesum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] += off;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand_total;
}

// ---------------------------------------------------------------------------------------
// Domain::pbc (domain.cpp:769-887) fused with NBin::coord2bin + the count pass of the
// counting sort.  One thread per owned atom.  32 B read + 32 B write of position, 4+4 B image.
// err[0] |= 1 non-finite coordinate, |= 2 atom outside the local bin grid (lost).
// ---------------------------------------------------------------------------------------
template <bool MULTI>
__global__ void __launch_bounds__(256) k_pbc_bin(int nlocal, double4 *__restrict__ xt,
                                                 int *__restrict__ image, Geom g, Owner own,
                                                 int *__restrict__ atombin, int *__restrict__ slot,
                                                 int *__restrict__ bincount,
                                                 int *__restrict__ dircount, int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  int img = image[i];
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
    atomicOr(err, 1);
    atombin[i] = 0;
    slot[i] = atomicAdd(&bincount[0], 1);
    return;
  }
  // triclinic: wrap, ownership and (later) the ghost slabs are decided in lamda coordinates; the
  // atom stays in lamda coordinates until the ghosts exist (Domain::x2lamda, verlet.cpp:293)
  if (g.tri) x2lamda(g, p.x, p.y, p.z);
  int idim, otherdims;
  if (g.periodic[0]) {
    if (p.x < g.boxlo[0]) {
      p.x += g.prd[0];
      idim = img & IMGMASK; otherdims = img ^ idim; idim--; idim &= IMGMASK;
      img = otherdims | idim;
    }
    if (p.x >= g.boxhi[0]) {
      p.x -= g.prd[0];
      p.x = fmax(p.x, g.boxlo[0]);
      idim = img & IMGMASK; otherdims = img ^ idim; idim++; idim &= IMGMASK;
      img = otherdims | idim;
    }
  }
  if (g.periodic[1]) {
    if (p.y < g.boxlo[1]) {
      p.y += g.prd[1];
      idim = (img >> IMGBITS) & IMGMASK; otherdims = img ^ (idim << IMGBITS); idim--; idim &= IMGMASK;
      img = otherdims | (idim << IMGBITS);
    }
    if (p.y >= g.boxhi[1]) {
      p.y -= g.prd[1];
      p.y = fmax(p.y, g.boxlo[1]);
      idim = (img >> IMGBITS) & IMGMASK; otherdims = img ^ (idim << IMGBITS); idim++; idim &= IMGMASK;
      img = otherdims | (idim << IMGBITS);
    }
  }
  if (g.periodic[2]) {
    if (p.z < g.boxlo[2]) {
      p.z += g.prd[2];
      idim = ((unsigned)img) >> IMG2BITS; otherdims = img ^ (idim << IMG2BITS); idim--; idim &= IMGMASK;
      img = otherdims | (idim << IMG2BITS);
    }
    if (p.z >= g.boxhi[2]) {
      p.z -= g.prd[2];
      p.z = fmax(p.z, g.boxlo[2]);
      idim = ((unsigned)img) >> IMG2BITS; otherdims = img ^ (idim << IMG2BITS); idim++; idim &= IMGMASK;
      img = otherdims | (idim << IMG2BITS);
    }
  }
  xt[i] = p;
  image[i] = img;
  if (MULTI) {
    // CommBrick::exchange: an atom outside [sublo,subhi) leaves to the sub-domain owning it
    const int ox = owner_dim(own, 0, p.x), oy = owner_dim(own, 1, p.y), oz = owner_dim(own, 2, p.z);
    if (ox | oy | oz) {
      if (ox == 2 || oy == 2 || oz == 2) {
        atomicOr(err, 2);
        atombin[i] = -1 - 13;
        slot[i] = 0;
        return;
      }
      const int dir = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
      atombin[i] = -1 - dir;
      slot[i] = atomicAdd(&dircount[dir], 1);
      return;
    }
  }
  if (g.tri) lamda2x(g, p.x, p.y, p.z);  // the position it will have when it is binned (verlet.cpp:313)
  int b = coord2bin(g, p.x, p.y, p.z);
  if (b < 0) {
    atomicOr(err, 2);
    b = 0;
  }
  atombin[i] = b;
  slot[i] = atomicAdd(&bincount[b], 1);
}

// Domain::lamda2x over the owned atoms once the ghosts have been created from their lamda
// coordinates (verlet.cpp:313); the ghosts were converted as they were made (k_ghost_make)
__global__ void __launch_bounds__(256) k_lamda2x(int n, double4 *__restrict__ xt, Geom g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xt[i];
  lamda2x(g, p.x, p.y, p.z);
  xt[i] = p;
}

// The slot an atom takes inside its bin (atomicAdd in k_pbc_bin) depends on the order in which
// thread blocks happen to run.  To make the atom order -- and with it the order of every list
// row and of every force sum -- reproducible from run to run, the atoms of a bin are ranked by
// their tag instead: pass 1 drops the tags at the provisional slots, pass 2 (k_permute_owned)
// counts the smaller tags among the ~2 bin mates.  4 + 4 B/atom of extra traffic per rebuild.
__global__ void __launch_bounds__(256) k_bin_keys(int n, const int *__restrict__ atombin,
                                                  const int *__restrict__ slot,
                                                  const int *__restrict__ binstart,
                                                  const int *__restrict__ tag, int *__restrict__ okey) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = atombin[i];
  if (b >= 0) okey[binstart[b] + slot[i]] = tag[i];
}

// Scatter pass of the counting sort: owned atoms are physically reordered by bin so that
// every later gather (list build, pair kernels) walks nearly-contiguous memory.  xhold
// (Neighbor::build prologue, neighbor.cpp:2520-2532) is written in the same pass.
__global__ void __launch_bounds__(256) k_permute_owned(
    int nlocal, const int *__restrict__ atombin, const int *__restrict__ slot,
    const int *__restrict__ okey,
    const int *__restrict__ binstart, const double4 *__restrict__ xt_in,
    double4 *__restrict__ xt_out, const double *__restrict__ vx_in, const double *__restrict__ vy_in,
    const double *__restrict__ vz_in, double *__restrict__ vx_out, double *__restrict__ vy_out,
    double *__restrict__ vz_out, const int *__restrict__ tag_in, int *__restrict__ tag_out,
    const int *__restrict__ mask_in, int *__restrict__ mask_out, const int *__restrict__ image_in,
    int *__restrict__ image_out, int *__restrict__ bin_out, double *__restrict__ xhx,
    double *__restrict__ xhy, double *__restrict__ xhz, Geom g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int b = atombin[i];
  if (b < 0) return;  // left for another sub-domain (k_pack_migrate shipped it)
  int d = binstart[b] + slot[i];
  if (okey) {  // rank by tag among the atoms of the bin (tags are unique)
    const int lo = binstart[b], hi = binstart[b + 1], mine = tag_in[i];
    int rank = 0;
    for (int q = lo; q < hi; q++) rank += okey[q] < mine;
    d = lo + rank;
  }
  const double4 p = xt_in[i];
  xt_out[d] = p;
  vx_out[d] = vx_in[i];
  vy_out[d] = vy_in[i];
  vz_out[d] = vz_in[i];
  tag_out[d] = tag_in[i];
  mask_out[d] = mask_in[i];
  image_out[d] = image_in[i];
  bin_out[d] = b;
  double hx = p.x, hy = p.y, hz = p.z;
  if (g.tri) lamda2x(g, hx, hy, hz);  // xhold is taken after lamda2x (neighbor.cpp:2520-2532)
  xhx[d] = hx;
  xhy[d] = hy;
  xhz[d] = hz;
}

// ---------------------------------------------------------------------------------------
// CommBrick::borders (comm_brick.cpp:720-899) restated as ONE pass over owned atoms instead
// of three dependent swap stages: an atom is sent in direction (dx,dy,dz) iff it lies in the
// corresponding slab in every non-zero dimension (x <= sublo+cutghost for -1, x >= subhi-
// cutghost for +1; inclusive compares as comm_brick.cpp:778).  The *set* of ghosts equals the
// reference's x->y->z cascade because each stage tests only its own coordinate.
// mode 0: count per direction; mode 1: fill sendlist at diroffset[dir] + running cursor.
// Warp-aggregated: one atomic per (warp, direction).
// ---------------------------------------------------------------------------------------
template <int FILL>
__global__ void __launch_bounds__(256) k_border(int nlocal, const double4 *__restrict__ xt, Geom g,
                                                int *__restrict__ dircount /*[27]*/,
                                                const int *__restrict__ diroffset /*[28]*/,
                                                int *__restrict__ sendlist,
                                                unsigned char *__restrict__ senddir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  if (i < nlocal) {
    const double4 p = xt[i];
    const double c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      lo[d] = g.send_left[d] && (c[d] <= g.slab_left_hi[d]);
      hi[d] = g.send_right[d] && (c[d] >= g.slab_right_lo[d]);
    }
  }
  const int any = lo[0] | lo[1] | lo[2] | hi[0] | hi[1] | hi[2];
  if (!__any_sync(0xffffffffu, any)) return;
  for (int dir = 0; dir < NDIR; dir++) {
    if (dir == 13) continue;
    const int dx = dir % 3 - 1, dy = (dir / 3) % 3 - 1, dz = dir / 9 - 1;
    const int ok = (dx == 0 || (dx < 0 ? lo[0] : hi[0])) && (dy == 0 || (dy < 0 ? lo[1] : hi[1])) &&
                   (dz == 0 || (dz < 0 ? lo[2] : hi[2]));
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (!m) continue;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&dircount[dir], __popc(m));
    if (FILL) {
      base = __shfl_sync(0xffffffffu, base, leader);
      if (ok) {
        const int p = diroffset[dir] + base + __popc(m & ((1u << lane) - 1u));
        sendlist[p] = i;
        senddir[p] = (unsigned char)dir;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// NPairBin<HALF=1,NEWTON=1,TRI=0,SIZE=0,ATOMONLY=1>::build (npair_bin.cpp:52-253).
// One thread per owned atom i (atoms are bin-sorted, so a warp covers 1-3 adjacent bins and
// its candidate reads hit L1).  Candidates = the half stencil regrouped into rows of
// x-contiguous bins, i.e. contiguous index ranges [ostart[b0], ostart[b1+1]) of owned atoms
// and [gstart[b0], gstart[b1+1]) of ghosts.  Own bin: owned j > i (list position), ghosts only
// if "above/right" by exact (z,y,x) compare (npair_bin.cpp:156-171).  Distance test in FP64
// with `rsq <= cutneighsq[itype][jtype]` (npair_bin.cpp:219) -> bit-exact pair set.
// Output: slot-major list, entry n of atom i at list_index(n, i, nstride, T) (kernels_pair.cuh);
// numneigh[i] = full count even when it
// exceeds maxneigh (then nothing past maxneigh is written and the host regrows the list).
// ---------------------------------------------------------------------------------------
template <bool ONETYPE, bool EXG = false>
__global__ void __launch_bounds__(128) k_build_half(
    int nlocal, int nstride, int maxneigh, int T, const double4 *__restrict__ xt,
    const int *__restrict__ atombin, const int *__restrict__ ostart,
    const int *__restrict__ gstart, Stencil st, double cutneighsq_one,
    const double *__restrict__ cutneighsq, int ntypes, int *__restrict__ numneigh,
    int *__restrict__ neigh, int *__restrict__ maxcount, ExGroups ex, const int *__restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w);
    const int b = atombin[i];
    const double *cut_i = ONETYPE ? nullptr : cutneighsq + (size_t)itype * (ntypes + 1);
    const int mi = (EXG && ex.n) ? mask[i] : 0;

    auto test = [&](int j) {
      if (EXG && ex.n && ex_group(ex, mi, mask[j])) return;
      const double4 pj = ld_xt(xt + j);
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      const double c = ONETYPE ? cutneighsq_one : cut_i[d2type(pj.w)];
      if (rsq <= c) {
        if (n < maxneigh) neigh[list_index(n, i, nstride, T)] = j;
        n++;
      }
    };

    for (int r = 0; r < st.nrows; r++) {
      const int b0 = b + st.rowoff[r] + st.dxlo[r];
      const int b1 = b + st.rowoff[r] + st.dxhi[r] + 1;
      if (r == 0) {
        // row (dz=0,dy=0): own bin first (dxlo == 0), then bins to the right
        for (int j = i + 1; j < ostart[b1]; j++) test(j);
        const int gown_end = gstart[b + 1];
        for (int gj = gstart[b]; gj < gown_end; gj++) {
          const int j = nlocal + gj;
          const double4 pj = ld_xt(xt + j);
          if (pj.z < pi.z) continue;
          if (pj.z == pi.z) {
            if (pj.y < pi.y) continue;
            if (pj.y == pi.y && pj.x < pi.x) continue;
          }
          test(j);
        }
        for (int gj = gown_end; gj < gstart[b1]; gj++) test(nlocal + gj);
      } else {
        for (int j = ostart[b0]; j < ostart[b1]; j++) test(j);
        for (int gj = gstart[b0]; gj < gstart[b1]; gj++) test(nlocal + gj);
      }
    }
    numneigh[i] = n;
  }
  // one atomicMax per warp
  int m = n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxcount, m);
}

// ---------------------------------------------------------------------------------------
// NPairBin<HALF=1,NEWTON=1,TRI=1,SIZE=0,ATOMONLY=1>::build (npair_bin.cpp:52-253, active branch
// :133-155): triclinic box.  The stencil is full in all three dimensions (nstencil_bin.cpp:36-62;
// `st` holds every (dz,dy) row, the row (0,0) includes the own bin).  An owned j is stored when it
// comes after i in the local order (each owned pair once); a ghost j is stored by the parity of
// itag+jtag (itag > jtag: odd sums, itag < jtag: even sums), a ghost image of i itself by the
// (z,y,x) comparison with the reference's tolerance `delta`.  Which of two owned atoms holds a pair
// depends on the local order, which is the reference's business as much as ours; the SET of pairs
// (as unordered tag pairs + image) does not, and that is what parity compares.
// ---------------------------------------------------------------------------------------
template <bool ONETYPE>
__global__ void __launch_bounds__(128) k_build_half_tri(
    int nlocal, int nstride, int maxneigh, int T, const double4 *__restrict__ xt,
    const int *__restrict__ tag, const int *__restrict__ atombin, const int *__restrict__ ostart,
    const int *__restrict__ gstart, Stencil st, double cutneighsq_one,
    const double *__restrict__ cutneighsq, int ntypes, double delta, int *__restrict__ numneigh,
    int *__restrict__ neigh, int *__restrict__ maxcount, ExGroups ex, const int *__restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 0;
  if (i < nlocal) {
    const double4 pi = xt[i];
    const int itype = d2type(pi.w), itag = tag[i];
    const int b = atombin[i];
    const double *cut_i = ONETYPE ? nullptr : cutneighsq + (size_t)itype * (ntypes + 1);
    const int mi = ex.n ? mask[i] : 0;

    auto test = [&](int j, const double4 &pj) {
      if (ex.n && ex_group(ex, mi, mask[j])) return;
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = rsq_ref(delx, dely, delz);
      const double c = ONETYPE ? cutneighsq_one : cut_i[d2type(pj.w)];
      if (rsq <= c) {
        if (n < maxneigh) neigh[list_index(n, i, nstride, T)] = j;
        n++;
      }
    };

    for (int r = 0; r < st.nrows; r++) {
      const int b0 = b + st.rowoff[r] + st.dxlo[r];
      const int b1 = b + st.rowoff[r] + st.dxhi[r] + 1;
      for (int j = max(ostart[b0], i + 1); j < ostart[b1]; j++) test(j, ld_xt(xt + j));
      for (int gj = gstart[b0]; gj < gstart[b1]; gj++) {
        const int j = nlocal + gj;
        const int jtag = tag[j];
        const double4 pj = ld_xt(xt + j);
        if (itag > jtag) {
          if ((itag + jtag) % 2 == 0) continue;
        } else if (itag < jtag) {
          if ((itag + jtag) % 2 == 1) continue;
        } else {
          if (fabs(pj.z - pi.z) > delta) {
            if (pj.z < pi.z) continue;
          } else if (fabs(pj.y - pi.y) > delta) {
            if (pj.y < pi.y) continue;
          } else {
            if (pj.x < pi.x) continue;
          }
        }
        test(j, pj);
      }
    }
    numneigh[i] = n;
  }
  int m = n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxcount, m);
}

// sum of numneigh (stored pairs) for statistics / tests
__global__ void __launch_bounds__(256) k_sum_int(int n, const int *__restrict__ a,
                                                 unsigned long long *__restrict__ out) {
  unsigned long long s = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256) k_sum_u16(int n, const unsigned short *__restrict__ a,
                                                 unsigned long long *__restrict__ out) {
  unsigned long long s = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// CSR export of the transposed list (test hook b200_get_neighbor_list)
__global__ void __launch_bounds__(256) k_export_csr(int nlocal, int nstride, int T,
                                                    const int *__restrict__ numneigh,
                                                    const int *__restrict__ neigh,
                                                    const long long *__restrict__ first,
                                                    int *__restrict__ flat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int n = numneigh[i];
  long long o = first[i];
  for (int k = 0; k < n; k++) flat[o + k] = neigh[list_index(k, i, nstride, T)];
}
