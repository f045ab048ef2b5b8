// kernels_halo.cuh -- ghost-atom halo across sub-domains (one sub-domain per GPU/process).
//
// The reference moves ghosts with 2 x maxneed dependent swaps per dimension, x then y then z
// (CommBrick::borders/forward_comm/reverse_comm, comm_brick.cpp:485-586, 720-899).  Here every
// sub-domain talks to its (up to) 26 neighbours directly: an owned atom is sent in direction
// (dx,dy,dz) iff it lies inside the ghost slab of every non-zero dimension, which yields the
// same SET of ghosts as the cascade (each stage tests only its own coordinate).  One grouped
// NCCL send/recv per halo instead of three dependent phases.
//
// Index spaces:
//   p  "send order":  sendlist[p] = owned atom, senddir[p] = direction, segments by direction
//   q  "recv order":  ghosts as they arrive, segments by the SENDER's direction
//   g  ghost slot:    ghosts are stored bin-sorted at xt[nlocal + g]; gsrc[g] >= 0 is the owner
//                     on this device (periodic self image), gsrc[g] = -1-q a ghost received
//                     from another rank.
// Directions whose neighbour is this very rank (1 rank along that dimension) never touch NCCL.
#pragma once
#include "common.cuh"

#define MIG_W 9  // doubles per migrating atom: x y z type | vx vy vz | (tag,mask) | (image,0)

__device__ __forceinline__ double pack2i(int lo, int hi) {
  return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo));
}
__device__ __forceinline__ void unpack2i(double d, int &lo, int &hi) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  lo = (int)(unsigned)(u & 0xffffffffull);
  hi = (int)(unsigned)(u >> 32);
}

// which neighbour sub-domain owns coordinate c in one dimension: 0 own, -1 left, +1 right,
// 2 = none of them (the atom moved further than one sub-domain: lost).
// Ownership is [sublo, subhi) as in CommBrick::exchange (comm_brick.cpp:655,702).
struct Owner {
  double lo[3][3], hi[3][3];  // [dim][0 own, 1 left neighbour, 2 right neighbour]
  int has[3][3];
};

__device__ __forceinline__ int owner_dim(const Owner &o, int d, double c) {
  if (c >= o.lo[d][0] && c < o.hi[d][0]) return 0;
  if (o.has[d][1] && c >= o.lo[d][1] && c < o.hi[d][1]) return -1;
  if (o.has[d][2] && c >= o.lo[d][2] && c < o.hi[d][2]) return 1;
  return 2;
}

// ---------------------------------------------------------------------------------------
// CommBrick::exchange (comm_brick.cpp:599-708), device side.  k_pbc_bin (kernels_neigh.cuh)
// classifies every owned atom after the periodic wrap: it either stays (binned, takes a slot
// in its bin) or leaves to the neighbour that now owns it (atombin = -1-dir, slot = position
// inside that direction's message).
// ---------------------------------------------------------------------------------------
// AtomVec::pack_exchange (atom_vec.cpp:1158-1173): x, v, tag, type, mask, image
__global__ void __launch_bounds__(256) k_pack_migrate(
    int nlocal, const int *__restrict__ atombin, const int *__restrict__ slot,
    const int *__restrict__ diroffset, const double4 *__restrict__ xt,
    const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
    const int *__restrict__ tag, const int *__restrict__ mask, const int *__restrict__ image,
    double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const int b = atombin[i];
  if (b >= 0) return;
  const int dir = -1 - b;
  if (dir == 13) return;
  double *o = buf + (size_t)(diroffset[dir] + slot[i]) * MIG_W;
  const double4 p = xt[i];
  o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
  o[4] = vx[i]; o[5] = vy[i]; o[6] = vz[i];
  o[7] = pack2i(tag[i], mask[i]);
  o[8] = pack2i(image[i], 0);
}

// AtomVec::unpack_exchange (atom_vec.cpp:1249-1267): arrivals are appended behind the current
// owned atoms, binned, and then take part in the counting sort like everybody else.
__global__ void __launch_bounds__(256) k_unpack_migrate(
    int narrive, int base, const double *__restrict__ buf, Geom g, Owner own,
    double4 *__restrict__ xt, double *__restrict__ vx, double *__restrict__ vy,
    double *__restrict__ vz, int *__restrict__ tag, int *__restrict__ mask,
    int *__restrict__ image, int *__restrict__ atombin, int *__restrict__ slot,
    int *__restrict__ bincount, int *__restrict__ err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= narrive) return;
  const double *o = buf + (size_t)k * MIG_W;
  const int i = base + k;
  const double4 p = make_double4(o[0], o[1], o[2], o[3]);
  xt[i] = p;
  vx[i] = o[4]; vy[i] = o[5]; vz[i] = o[6];
  int a, b2;
  unpack2i(o[7], a, b2);
  tag[i] = a; mask[i] = b2;
  unpack2i(o[8], a, b2);
  image[i] = a;
  if (owner_dim(own, 0, p.x) != 0 || owner_dim(own, 1, p.y) != 0 || owner_dim(own, 2, p.z) != 0)
    atomicOr(err, 2);  // arrived at the wrong sub-domain: moved more than one sub-domain
  double bx = p.x, by = p.y, bz = p.z;
  if (g.tri) lamda2x(g, bx, by, bz);  // arrivals are in lamda coordinates like everybody else
  int b = coord2bin(g, bx, by, bz);
  if (b < 0) {
    atomicOr(err, 2);
    b = 0;
  }
  atombin[i] = b;
  slot[i] = atomicAdd(&bincount[b], 1);
}

// ---------------------------------------------------------------------------------------
// Borders: AtomVec::pack_border / unpack_border (atom_vec.cpp:796-830, 1026-1042).
// Message record = {x+shift, y+shift, z+shift, (type,tag)}: 4 doubles per ghost.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_border(int nsend, const int *__restrict__ sendlist,
                                                     const unsigned char *__restrict__ senddir,
                                                     unsigned remote_mask, Geom g,
                                                     const double4 *__restrict__ xt,
                                                     const int *__restrict__ tag,
                                                     double4 *__restrict__ buf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsend) return;
  const int dir = senddir[p];
  if (!((remote_mask >> dir) & 1u)) return;
  const int src = sendlist[p];
  double4 q = xt[src];
  q.x = q.x + g.shift[dir][0];
  q.y = q.y + g.shift[dir][1];
  q.z = q.z + g.shift[dir][2];
  q.w = pack2i(d2type(q.w), tag[src]);
  buf[p] = q;
}

// group masks of the border atoms, one double each, same order as the border records
// (AtomVec::pack_border sends mask with every ghost, atom_vec.cpp:796-830; here only on request)
__global__ void __launch_bounds__(256) k_pack_mask(int nsend, const int *__restrict__ sendlist,
                                                   const unsigned char *__restrict__ senddir,
                                                   unsigned remote_mask, const int *__restrict__ mask,
                                                   double *__restrict__ buf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsend) return;
  if (!((remote_mask >> senddir[p]) & 1u)) return;
  buf[p] = (double)mask[sendlist[p]];
}

// Ghost creation, pass 1 over recv order q: take the ghost from the local owner (self image)
// or from the received border record, bin it, take a slot in the ghost histogram.
__global__ void __launch_bounds__(256) k_ghost_make(
    int nghost, const int *__restrict__ recvoffset /*[28]*/, const int *__restrict__ sendoffset,
    unsigned remote_mask, const int *__restrict__ sendlist, Geom g, const double4 *__restrict__ xt,
    const int *__restrict__ tag, const double4 *__restrict__ rbuf, double4 *__restrict__ gtmp,
    int *__restrict__ gtag_tmp, int *__restrict__ gsrc_tmp, int *__restrict__ gbin,
    int *__restrict__ gslot, unsigned char *__restrict__ gdir_tmp, int *__restrict__ gbincount,
    int *__restrict__ err) {
  __shared__ int roff[NDIR + 1], soff[NDIR + 1];
  if (threadIdx.x <= NDIR) {
    roff[threadIdx.x] = recvoffset[threadIdx.x];
    soff[threadIdx.x] = sendoffset[threadIdx.x];
  }
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nghost) return;
  int dir = 0;
#pragma unroll 1
  while (dir < NDIR - 1 && q >= roff[dir + 1]) dir++;
  double4 r;
  int t, src;
  if ((remote_mask >> dir) & 1u) {
    r = rbuf[q];
    int ty;
    unpack2i(r.w, ty, t);
    r.w = type2d(ty);
    src = -1 - q;
  } else {
    src = sendlist[soff[dir] + (q - roff[dir])];
    r = xt[src];
    r.x = r.x + g.shift[dir][0];
    r.y = r.y + g.shift[dir][1];
    r.z = r.z + g.shift[dir][2];
    t = tag[src];
  }
  if (g.tri) lamda2x(g, r.x, r.y, r.z);  // Domain::lamda2x(nlocal+nghost), verlet.cpp:313
  int b = coord2bin(g, r.x, r.y, r.z);
  if (b < 0) {
    atomicOr(err, 2);
    b = 0;
  }
  gtmp[q] = r;
  gtag_tmp[q] = t;
  gsrc_tmp[q] = src;
  gbin[q] = b;
  gdir_tmp[q] = (unsigned char)dir;
  gslot[q] = atomicAdd(&gbincount[b], 1);
}

// Reproducible ghost order (see k_bin_keys): a ghost is identified by (tag, direction it came
// from); the ghosts of a bin are ranked by that key instead of by the order of the atomicAdd.
__global__ void __launch_bounds__(256) k_ghost_keys(int nghost, const int *__restrict__ gbin,
                                                    const int *__restrict__ gslot,
                                                    const int *__restrict__ gstart,
                                                    const int *__restrict__ gtag_tmp,
                                                    const unsigned char *__restrict__ gdir_tmp,
                                                    long long *__restrict__ gkey) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nghost) return;
  gkey[gstart[gbin[q]] + gslot[q]] = ((long long)gtag_tmp[q] << 5) | gdir_tmp[q];
}

// Ghost creation, pass 2: place ghosts bin-sorted behind the owned atoms; gsrc/gdir are the
// receiver-side equivalent of sendlist/firstrecv/pbc_flag of comm_brick, reused every step.
__global__ void __launch_bounds__(256) k_ghost_place(
    int nghost, int nlocal, const double4 *__restrict__ gtmp, const int *__restrict__ gtag_tmp,
    const int *__restrict__ gsrc_tmp, const int *__restrict__ gbin, const int *__restrict__ gslot,
    const unsigned char *__restrict__ gdir_tmp, const int *__restrict__ gstart,
    double4 *__restrict__ xt, int *__restrict__ tag, int *__restrict__ mask,
    int *__restrict__ gsrc, unsigned char *__restrict__ gdir, double4 *__restrict__ xt_alt,
    const long long *__restrict__ gkey, const double *__restrict__ rmask) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nghost) return;
  int gi = gstart[gbin[q]] + gslot[q];
  if (gkey) {
    const int lo = gstart[gbin[q]], hi = gstart[gbin[q] + 1];
    const long long mine = ((long long)gtag_tmp[q] << 5) | gdir_tmp[q];
    int rank = 0;
    for (int k = lo; k < hi; k++) rank += gkey[k] < mine;
    gi = lo + rank;
  }
  const int src = gsrc_tmp[q];
  xt[nlocal + gi] = gtmp[q];
  // the other position buffer takes the record too: a pair kernel with the fused integrator
  // writes owned positions there and the halo then refreshes x,y,z only -- the type must be in place
  xt_alt[nlocal + gi] = gtmp[q];
  tag[nlocal + gi] = gtag_tmp[q];
  // (group masks of ghosts owned elsewhere travel only when a list rule reads them: k_pack_mask)
  mask[nlocal + gi] = src >= 0 ? mask[src] : (rmask ? (int)rmask[q] : 1);
  gsrc[gi] = src;
  gdir[gi] = gdir_tmp[q];
}

// ---------------------------------------------------------------------------------------
// Per-step halo.  forward = CommBrick::forward_comm + pack_comm/unpack_comm
// (comm_brick.cpp:485-538, atom_vec.cpp:354-440,561); reverse = reverse_comm +
// pack_reverse/unpack_reverse (comm_brick.cpp:545-586, atom_vec.cpp:672,729).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_forward(int nsend, const int *__restrict__ sendlist,
                                                      const unsigned char *__restrict__ senddir,
                                                      unsigned remote_mask, Geom g,
                                                      const double4 *__restrict__ xt,
                                                      double *__restrict__ buf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsend) return;
  const int dir = senddir[p];
  if (!((remote_mask >> dir) & 1u)) return;
  const double4 q = xt[sendlist[p]];
  double *o = buf + 3 * (size_t)p;
  halo_shift(g, dir, q.x, q.y, q.z, o);
}

struct Vec3Ptr {
  double *a[3];
};

// ghost g <- local owner + periodic shift, or <- received record
// fclear (may be null): ghost forces are zeroed in the same pass -- Verlet::force_clear for the
// tile path, whose pair kernels accumulate into ghost slots only (owned f_i is stored)
__global__ void __launch_bounds__(256) k_unpack_forward(int nghost, int nlocal,
                                                        const int *__restrict__ gsrc,
                                                        const unsigned char *__restrict__ gdir,
                                                        Geom g, const double *__restrict__ rbuf,
                                                        double4 *__restrict__ xt, Vec3Ptr fclear) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nghost) return;
  if (fclear.a[0]) {
    fclear.a[0][nlocal + k] = 0.0;
    fclear.a[1][nlocal + k] = 0.0;
    fclear.a[2][nlocal + k] = 0.0;
  }
  const int src = gsrc[k];
  double *o = reinterpret_cast<double *>(xt + nlocal + k);
  if (src >= 0) {
    const int dir = gdir[k];
    const double4 q = xt[src];
    halo_shift(g, dir, q.x, q.y, q.z, o);
  } else {
    const double *r = rbuf + 3 * (size_t)(-1 - src);
    o[0] = r[0];
    o[1] = r[1];
    o[2] = r[2];
  }
}

// ghost contributions: local images are added to their owner at once (an owner has up to 7
// images -> RED.ADD.F64), ghosts owned elsewhere are packed for the way back.
template <int W>
__global__ void __launch_bounds__(256) k_pack_reverse(int nghost, int nlocal,
                                                      const int *__restrict__ gsrc, Vec3Ptr f,
                                                      double *__restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nghost) return;
  const int src = gsrc[k];
  if (src >= 0) {
#pragma unroll
    for (int d = 0; d < W; d++) atomicAdd(&f.a[d][src], f.a[d][nlocal + k]);
  } else {
    double *o = buf + (size_t)W * (-1 - src);
#pragma unroll
    for (int d = 0; d < W; d++) o[d] = f.a[d][nlocal + k];
  }
}

template <int W>
__global__ void __launch_bounds__(256) k_unpack_reverse(int nsend, const int *__restrict__ sendlist,
                                                        const unsigned char *__restrict__ senddir,
                                                        unsigned remote_mask,
                                                        const double *__restrict__ buf, Vec3Ptr f) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsend) return;
  if (!((remote_mask >> senddir[p]) & 1u)) return;
  const int i = sendlist[p];
  const double *r = buf + (size_t)W * p;
#pragma unroll
  for (int d = 0; d < W; d++) atomicAdd(&f.a[d][i], r[d]);
}

// scalar forward (EAM fp, pair_eam.cpp:1600-1621)
__global__ void __launch_bounds__(256) k_pack_forward_scalar(int nsend,
                                                             const int *__restrict__ sendlist,
                                                             const unsigned char *__restrict__ senddir,
                                                             unsigned remote_mask,
                                                             const double *__restrict__ a,
                                                             double *__restrict__ buf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nsend) return;
  if (!((remote_mask >> senddir[p]) & 1u)) return;
  buf[p] = a[sendlist[p]];
}
__global__ void __launch_bounds__(256) k_unpack_forward_scalar(int nghost, int nlocal,
                                                               const int *__restrict__ gsrc,
                                                               const double *__restrict__ rbuf,
                                                               double *__restrict__ a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nghost) return;
  const int src = gsrc[k];
  a[nlocal + k] = src >= 0 ? a[src] : rbuf[-1 - src];
}

// =======================================================================================
// Peer-memory halo (NVLink / NVSwitch): the pack kernel stores straight into the neighbour's
// staging buffer (a CUDA-IPC mapping of the peer's arena) and raises an arrival flag there;
// the neighbour's unpack kernel waits for its arrival flags, reads, and acknowledges.  One
// kernel does "pack + transfer", no send/recv rendezvous, no intermediate copy.  NCCL
// (halo_exchange) stays for the rebuild-time traffic and as the fallback transport.
//
// Protocol, per direction and per buffer class (F: messages into peers' rbuf, R: into sbuf):
//   sender  : wait ack_in[dir]  >= seq-1   (peer consumed my previous message)
//             store records; __threadfence_system(); last block: flag_out[dir] = seq
//   receiver: wait flag_in[dir] >= seq; read; last block: ack_out[dir] = seq
// seq advances identically on every rank (all ranks execute the same halo sequence).  Spins
// give up after P2P_SPIN_LIMIT cycles and raise err |= 8 instead of hanging the GPU.
// =======================================================================================
#define P2P_SPIN_LIMIT (4000000000ll)

struct P2PMap {
  double *dst[NDIR];          // peer staging buffer (base of the peer's rbuf or sbuf)
  int dstoff[NDIR];           // first record of my message inside it
  long long *flag_out[NDIR];  // peer's arrival flag for my message
  long long *ack_out[NDIR];   // peer's ack flag for the message I consume
  const long long *flag_in;   // own arrival flags [32]
  const long long *ack_in;    // own ack flags [32]
  unsigned out_mask, in_mask; // directions with data to send / to receive
};

__device__ __forceinline__ long long ld_relaxed_sys(const long long *p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(long long *p, long long v) {
  asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// lanes 0..26 of the block poll one direction each until flags[dir] >= need, then one
// system-scope fence orders the flag reads before the data reads of the whole block
__device__ __forceinline__ void p2p_wait(const long long *flags, unsigned mask, long long need,
                                         int *err) {
  if (need > 0 && threadIdx.x < NDIR && ((mask >> threadIdx.x) & 1u)) {
    const long long t0 = clock64();
    while (ld_relaxed_sys(flags + threadIdx.x) < need) {
      if (clock64() - t0 > P2P_SPIN_LIMIT) {
        atomicOr(err, 8);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

// the last block to get here publishes `value` to out[dir] for every dir in mask: every
// thread fences its own stores (system scope) before the block is counted, the last block
// fences once more and then lanes 0..26 store one flag each
__device__ __forceinline__ void p2p_publish(long long *const *out, unsigned mask, long long value,
                                            unsigned *counter) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x < NDIR && ((mask >> threadIdx.x) & 1u)) {
    __threadfence_system();
    st_relaxed_sys(out[threadIdx.x], value);
  }
}

// forward: MODE 0 = positions (3 doubles, + periodic shift), MODE 1 = one scalar per atom
template <int MODE>
__global__ void __launch_bounds__(256) k_p2p_pack_forward(
    int nsend, const int *__restrict__ sendlist, const unsigned char *__restrict__ senddir,
    const int *__restrict__ sendoffset, Geom g, const double4 *__restrict__ xt,
    const double *__restrict__ a, P2PMap pm, long long seq, unsigned *counter, int *err) {
  p2p_wait(pm.ack_in, pm.out_mask, seq - 1, err);
  // (grid-stride: sub-domains that share a GPU launch these kernels with a capped grid so that
  // every spinning block of every sub-domain is resident at once -- engine.cu p2p_grid)
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nsend; p += gridDim.x * blockDim.x) {
    const int dir = senddir[p];
    if ((pm.out_mask >> dir) & 1u) {
      const size_t q = (size_t)pm.dstoff[dir] + (p - sendoffset[dir]);
      if (MODE == 0) {
        const double4 r = xt[sendlist[p]];
        double *o = pm.dst[dir] + 3 * q;
        halo_shift(g, dir, r.x, r.y, r.z, o);
      } else {
        pm.dst[dir][q] = a[sendlist[p]];
      }
    }
  }
  p2p_publish(pm.flag_out, pm.out_mask, seq, counter);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_p2p_unpack_forward(
    int nghost, int nlocal, const int *__restrict__ gsrc, const unsigned char *__restrict__ gdir,
    Geom g, const double *__restrict__ rbuf, double4 *__restrict__ xt, double *__restrict__ a,
    P2PMap pm, long long seq, unsigned *counter, int *err, Vec3Ptr fclear) {
  p2p_wait(pm.flag_in, pm.in_mask, seq, err);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nghost; k += gridDim.x * blockDim.x) {
    const int src = gsrc[k];
    if (MODE == 0 && fclear.a[0]) {
      fclear.a[0][nlocal + k] = 0.0;
      fclear.a[1][nlocal + k] = 0.0;
      fclear.a[2][nlocal + k] = 0.0;
    }
    if (MODE == 0) {
      double *o = reinterpret_cast<double *>(xt + nlocal + k);
      if (src >= 0) {
        const int dir = gdir[k];
        const double4 q = xt[src];
        halo_shift(g, dir, q.x, q.y, q.z, o);
      } else {
        const double *r = rbuf + 3 * (size_t)(-1 - src);
        o[0] = __ldcv(r);
        o[1] = __ldcv(r + 1);
        o[2] = __ldcv(r + 2);
      }
    } else {
      a[nlocal + k] = src >= 0 ? a[src] : __ldcv(rbuf + (-1 - src));
    }
  }
  p2p_publish(pm.ack_out, pm.in_mask, seq, counter);
}

// reverse: ghost contributions of W SoA arrays; local images are added to their owner at
// once, ghosts owned elsewhere are stored into the owner's sbuf (send order of the owner)
template <int W>
__global__ void __launch_bounds__(256) k_p2p_pack_reverse(
    int nghost, int nlocal, const int *__restrict__ gsrc, const unsigned char *__restrict__ gdir,
    const int *__restrict__ recvoffset, Vec3Ptr f, P2PMap pm, long long seq, unsigned *counter,
    int *err) {
  p2p_wait(pm.ack_in, pm.out_mask, seq - 1, err);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nghost; k += gridDim.x * blockDim.x) {
    const int src = gsrc[k];
    if (src >= 0) {
#pragma unroll
      for (int d = 0; d < W; d++) atomicAdd(&f.a[d][src], f.a[d][nlocal + k]);
    } else {
      const int dir = gdir[k];
      const size_t q = (size_t)pm.dstoff[dir] + ((-1 - src) - recvoffset[dir]);
      double *o = pm.dst[dir] + (size_t)W * q;
#pragma unroll
      for (int d = 0; d < W; d++) o[d] = f.a[d][nlocal + k];
    }
  }
  p2p_publish(pm.flag_out, pm.out_mask, seq, counter);
}

template <int W>
__global__ void __launch_bounds__(256) k_p2p_unpack_reverse(
    int nsend, const int *__restrict__ sendlist, const unsigned char *__restrict__ senddir,
    const double *__restrict__ sbuf, Vec3Ptr f, P2PMap pm, long long seq, unsigned *counter,
    int *err) {
  p2p_wait(pm.flag_in, pm.in_mask, seq, err);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nsend; p += gridDim.x * blockDim.x) {
    if ((pm.in_mask >> senddir[p]) & 1u) {
      const int i = sendlist[p];
      const double *r = sbuf + (size_t)W * p;
#pragma unroll
      for (int d = 0; d < W; d++) atomicAdd(&f.a[d][i], __ldcv(r + d));
    }
  }
  p2p_publish(pm.ack_out, pm.in_mask, seq, counter);
}
