// engine.cu -- context, device memory, rebuild/step orchestration and the C ABI
// (include/b200_md.h) of the B200-native short-range MD hot path.  sm_100a only.
//
// Device layout (all resident in HBM for the whole run):
//   xt     double4[nmax]   {x,y,z,type}  owned atoms [0,nlocal) sorted by bin, then ghosts
//                                        [nlocal,nlocal+nghost) sorted by bin
//   v      3 x double[nmax] SoA,  f 3 x double[nmax] SoA (owned + ghost, Newton on)
//   tag/mask/image int32[nmax], xhold 3 x double[nmax]
//   ostart/gstart  int32[mbins+1]  bin -> first owned / first ghost index (counting sort)
//   tile list (lj/cut): uint16 entries into the shared-memory staging of a bin tile, 8 per
//          16-byte word, [slot/8][list row]; see kernels_tile.cuh
//   neigh  (eam, B200_LIST=flat) int32[maxneigh][nstride] transposed half list; numneigh
//          int32[nlocal] holds the half-list count in both layouts
// Streams: `stream` (high priority) carries the timestep and the halo; `stream2` only runs the
// interior tiles beside the halo on multi-GPU runs; the plain timestep of a single sub-domain is
// replayed from a CUDA graph (graph_step).
//   gsrc   int32[nghost], gdir uint8[nghost]: owner index and direction of every ghost
#include "common.cuh"
#include "kernels_halo.cuh"
#include "kernels_neigh.cuh"
#include "kernels_pair.cuh"
#include "kernels_pair_mixed.cuh"
#include "kernels_step.cuh"
#include "kernels_tile.cuh"
#include "kernels_tile2.cuh"
#include "kernels_eam2.cuh"
#include "kernels_build2.cuh"
#include "kernels_peratom.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <type_traits>

namespace {

template <class T>
struct DBuf {
  T *p = nullptr;
  size_t cap = 0;
};

struct PhaseRec {
  int ph;
  cudaEvent_t a, b;
};

}  // namespace

// ------------------------------------------------------------------ in-process group
// Several sub-domains driven by ONE host process (the LAMMPS package's `package b200 gpus N`,
// no MPI in the image; precedent for one process owning several devices:
// GPU/fix_gpu.cpp:123-247).  Every sub-domain is an ordinary context with its own host thread;
// the few collectives of a rebuild are rendezvous through shared host memory, bulk rebuild
// traffic is cudaMemcpyAsync between the contexts' buffers, and the per-step halo is the same
// peer-memory kernels as between processes (plain pointers instead of CUDA IPC mappings).
// Sub-domains may share a device (how the 1-GPU test box exercises migration and borders).
struct b200_group {
  int n = 0;
  std::vector<b200_ctx *> ctx;
  int grid[3] = {1, 1, 1};
  std::string err;
  bool shared_dev = false;  // two or more sub-domains on one device
  // sense-reversing barrier of the n context threads
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long phase = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const unsigned long long my = phase;
    if (++arrived == n) {
      arrived = 0;
      phase++;
      cv.notify_all();
    } else
      cv.wait(lk, [&] { return phase != my; });
  }
  // exchange tables (written between two barriers, read after the first)
  std::vector<int> counts;      // [n][32]
  std::vector<double> red;      // [n][8]
  std::vector<int> redi;        // [n]
  struct Pub {
    const double *src = nullptr;
    int off[NDIR + 1] = {0};
    cudaEvent_t ready = nullptr;
  };
  std::vector<Pub> pub;
  // one worker thread per context: run(fn) executes fn(i) on all of them and joins
  std::vector<std::thread> workers;
  std::function<int(int)> job;
  std::vector<int> rc;
  unsigned long long job_seq = 0, job_done = 0;
  int job_finished = 0;
  bool quit = false;
  std::mutex jmu;
  std::condition_variable jcv, dcv;
  // atoms handed to b200_group_set_atoms, split by owning sub-domain (kept for re-upload)
  std::vector<int> owner_of;
};

struct b200_ctx {
  int device = 0, prec = 0;
  b200_group *grp = nullptr;  // in-process group this context belongs to (rank = its index)
  cudaStream_t stream = nullptr;
  std::string err;
  size_t dev_bytes = 0;
  int64_t launches = 0;

  // domain
  bool have_box = false;
  double boxlo[3], boxhi[3], prd[3];
  int periodic[3] = {1, 1, 1};
  int procgrid[3] = {1, 1, 1}, myloc[3] = {0, 0, 0};
  double sublo[3], subhi[3];  // comm frame (lamda coordinates in a triclinic box)
  double osublo[3], osubhi[3];  // extent of the sub-domain in box coordinates (bounding box if triclinic)
  bool tri = false;           // triclinic box: tilt factors, Force::angstrom for the list rule's delta
  double xy = 0.0, xz = 0.0, yz = 0.0, angstrom = 1.0;
  // neighbor settings
  double skin = 0.3;
  int every = 1, delay = 0, dist_check = 1, one = 2000;
  double cutneighmax = 0, cutneighmaxsq = 0, triggersq = 0, cutghost = 0;
  std::vector<double> cutneighsq_h;
  std::vector<int> ex_type;  // neigh_modify exclude type: [(ntypes+1)^2] flags, empty = none
  bool build_once = false;   // neigh_modify once yes: the list of setup is never rebuilt
  // a box that changes during the run (fix npt/b200 -> b200_remap): the displacement check moves
  // into decide() and its trigger distance shrinks with the box corners (neighbor.cpp:2443-2455)
  bool box_changes = false;
  double boxlo_hold[3] = {0, 0, 0}, boxhi_hold[3] = {0, 0, 0}, deltasq = 0.0;
  DBuf<double> cutneighsq_d;
  int64_t ago = 0, nbuilds = 0, ndanger = 0;
  Geom geom;
  Stencil stencil;
  int nstencil = 0;
  bool geom_ready = false;
  // atoms
  int ntypes = 0, nlocal = 0, nghost = 0, nmax = 0;
  std::vector<double> mass_h;
  DBuf<double> mass_d;
  DBuf<double> lang_d, lang_u;  // fix langevin: per-type prefactors (+ 3 sums), host-drawn uniforms
  std::vector<double> lang_h;
  double4 *xt[2] = {nullptr, nullptr};
  // mixed precision: sub-domain-wide fixed-point records {qx,qy,qz,type} of all atoms (k_tile_lj2f
  // stages them with one cp.async each); [cur] is live like xt[cur]; kept current by the fused
  // integrator's epilogue (owned) and k_xt_to_q (ghosts after a halo, everything after a rebuild)
  int4 *qrec[2] = {nullptr, nullptr};
  QGeom qgeom;
  bool gq_ok = false, q_owned_valid = false, q_ghost_valid = false;
  double *v[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
  double *f[3] = {nullptr, nullptr, nullptr};
  double *xh[3] = {nullptr, nullptr, nullptr};
  int *tag[2] = {nullptr, nullptr}, *mask[2] = {nullptr, nullptr}, *image[2] = {nullptr, nullptr};
  int *atombin[2] = {nullptr, nullptr}, *slot = nullptr;
  int cur = 0;  // which of the ping-pong buffers holds the live atoms
  // bins
  DBuf<int> ostart, gstart, tilesum;
  // ghosts (index spaces p / q / g: see kernels_halo.cuh)
  DBuf<int> sendlist, gsrc, gbin, gslot, gtag_tmp, gsrc_tmp;
  DBuf<int> okey;          // tags at the provisional sort slots (reproducible atom order, k_bin_keys)
  DBuf<long long> gkey;    // (tag, direction) keys of the ghosts, likewise
  bool stable_order = true;  // B200_STABLE=0: keep the arrival order of the atomics
  DBuf<unsigned char> senddir, gdir, gdir_tmp;
  DBuf<double4> gtmp;
  int *counts = nullptr;      // [32] device: per-direction counters [0,27), [27] error flags
  int *diroffset = nullptr;   // [28] device: send-order segment offsets
  int *recvoffset = nullptr;  // [28] device: recv-order segment offsets
  int nsend = 0;
  int sendoff[NDIR + 1] = {0}, recvoff[NDIR + 1] = {0};
  // multi-GPU halo: one sub-domain per rank, NCCL send/recv between the 26 neighbours
  ncclComm_t nccl = nullptr;
  int nranks = 1, rank = 0;
  int nbr[NDIR];             // rank of the neighbour sub-domain in each direction (-1: none)
  std::vector<int> rankmap;  // optional grid location -> rank map supplied by the host
  unsigned remote_mask = 0;  // directions whose neighbour is another rank
  Owner owner;
  DBuf<double> sbuf, rbuf;   // halo staging in send order / recv order (up to 4 doubles per atom)
  // peer-memory halo (CUDA IPC over NVLink): sbuf/rbuf live in one arena that neighbours map
  bool p2p = false;               // transport of the per-step halo: peer stores (true) or NCCL
  char *arena = nullptr;          // [sbuf | rbuf], one cudaMalloc, exported through IPC
  size_t arena_bytes = 0, arena_roff = 0;
  unsigned arena_gen = 0;         // bumped at every (re)allocation
  long long *pflags = nullptr;    // [4][32] own flags: arrival F, arrival R, ack F, ack R
  unsigned *p2p_counter = nullptr;  // [4] last-block counters
  std::vector<char *> peer_arena;   // per rank: mapping of its arena (nullptr = not mapped)
  std::vector<void *> peer_base;    // per rank: what cudaIpcOpenMemHandle returned (to close it)
  std::vector<long long *> peer_flags;
  std::vector<size_t> peer_roff;
  std::vector<unsigned> peer_gen;
  std::vector<int> peer_off;        // per rank: sendoff[28] then recvoff[28]
  P2PMap fwdP, revP;
  long long seqF = 0, seqR = 0;
  DBuf<double> mig_send, mig_recv;
  int *allcounts = nullptr;  // [nranks*32] device (all-gathered counters)
  int *h_counts = nullptr;   // pinned [nranks*32]
  char *ag_dev = nullptr;    // persistent staging of allgather_host: [(nranks+1) * AG_BYTES] device
  char *ag_host = nullptr;   // ... and pinned host
  // every rank tracks every rank's halo-arena capacity (doubles): the growth rule is
  // deterministic in the all-gathered counts, so all ranks know without talking when some arena
  // is about to move and the IPC handles must be exchanged again
  std::vector<size_t> peer_cap_s, peer_cap_r;
  std::vector<int> rank_nbr;  // [nranks][27] neighbour ranks of every rank
  // list
  DBuf<int> neigh, numneigh;
  int maxneigh = 0, nstride = 0, max_numneigh = 0;
  int tpa = 2;  // lanes per atom in the pair kernels = interleave factor of the list (1,2,4,8)
  // bin-tile list (kernels_tile.cuh): the default; B200_LIST=flat selects the flat int32 list
  bool use_tiles = true;
  int list_mode = 0;            // B200_LIST: 0 auto (tiles for lj/cut, flat for eam), 1 tile, 2 flat
  bool tiles_active = false;    // the current list is the tile list
  // the tile list holds every ghost partner of an owned atom (k_tile_build<.,FULLGHOST>): the
  // pair kernel stores complete owned forces, nothing is scattered onto ghosts, and the step has
  // no force clear and no reverse halo; energy and virial are tallied inside the pair kernel
  bool full_ghost = false;
  int tile_req[3] = {0, 0, 0};  // B200_TILE=tx,ty,tz override of the tile size
  int tile_level = -1;
  TileGeom tg;
  FullStencil fst;
  int ibin_lo[3] = {0, 0, 0}, ibin_n[3] = {0, 0, 0};  // local bins that can hold owned atoms
  DBuf<int> tile_ibase;
  DBuf<unsigned char> tile_hdrs;  // every tile's row tables, written at each rebuild (tile_rows_cached)
  DBuf<unsigned short> tl_iloc, tl_num;
  DBuf<int> tl_gi;              // global index of the owned atom of every list row
  // FP64 lj/cut on tiles: k_tile_lj2 launch shape {threads, CTAs/SM, pair bodies in flight}
  // (B200_LJ2=threads,minb,ilp; B200_LJ2=0 selects the first-generation k_tile_lj)
  int lj2[3] = {352, 2, 2};
  bool use_lj2 = true;
  // eam on tiles, second generation (kernels_eam2.cuh): FULLGHOST + NEAR/FAR split rows, density
  // and embedding in one kernel, fix nve fused into the force kernel.  Single-element potentials
  // in FP64 (B200_EAM2=0 / `package b200 eam2 no` selects the flat half list kernels).
  ExGroups exg = ExGroups{0, {0}, {0}};  // neigh_modify exclude group
  DBuf<double> mask_s, mask_r;          // group masks of border atoms / remote ghosts (only with exg)
  VOps vops = VOps{};         // deferred per-atom integrator operations (k_vops), in issue order
  int vops_chk = 0;           // the queued drift also takes the displacement vote
  bool lazy_ops = true;       // package b200 lazy yes|no: queue them (default) or run each at once
  int newton = 1;               // Force::newton_pair; 0: lists hold every owned-ghost pair on both sides
  int eam2 = 2;                 // 0 never, 1 whenever usable, 2 auto: small sub-domains (see eam2_usable)
  long long eam2_max_bins = 60000;
  bool build2 = false;          // warp-per-bin list build (kernels_build2.cuh; B200_BUILD2=1): measured slower
                                // than k_tile_build (4.5 vs 2.7 ms per build at 4 M atoms), kept as a cross-check
  bool eam2_active = false;     // the current list was built for those kernels
  double eam2_margin = 0.35;    // NEAR = stored within force cutoff + eam2_margin * skin (B200_EAM2_MARGIN)
  DBuf<unsigned short> tl_far;  // FAR entries per row (SPLIT rows)
  DBuf<double> peratom;         // b200_pair_peratom: eatom + 6 vatom arrays over owned (+ ghost) atoms
  int lj2f[2] = {256, 4};       // k_tile_lj2f (mixed): threads, CTAs per SM (B200_LJ2F=threads,minb)
  DBuf<uint4> tl_list;
  int *tflags = nullptr;  // [8] device: max staged, max owned/tile, max entries, max FWD, owned total, overflow
  int tile_NI = 0, tile_scap = 0, tile_slots = 0, tile_threads = 256, tile_maxfull = 0, tile_maxown = 0;
  // halo/compute overlap (multi-GPU, tile list): tiles that stage no ghost ("interior") run on
  // stream2 while the forward halo, the boundary tiles and the reverse halo run on `stream`
  DBuf<int> tile_bflag, tile_bpos, tile_ids;
  int tile_nint = 0, tile_nbnd = 0;
  bool overlap = true;  // B200_OVERLAP=0 disables
  // CUDA graph of one plain timestep (no rebuild, no tallies, one sub-domain): small systems
  // are launch-bound (bench/in.lj: 32 k atoms, ~9 launches of a few microseconds each)
  bool ghost_f_clean = false;  // ghost forces are zero (forward halo / rebuild did it; tile path)
  bool use_graph = true;  // B200_GRAPH=0 disables
  cudaGraphExec_t step_graph = nullptr;
  int graph_launches = 0;
  bool mixed_fx = true; // mixed lj/cut on tiles: fixed-point staged positions (B200_MIXED_FX=0: FP64 staging)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // pair
  int pair_style = 0;  // 1 lj/cut, 2 eam
  std::vector<double> cutsq_h;
  LJOne lj_one;
  DBuf<double> lj_tab;
  LJOneF lj_onef;
  DBuf<float> lj_tabf;   // mixed mode: lj1..lj4, offset as float tables
  EAMParamsF eamf;
  DBuf<float> eam_f;     // mixed mode: rhor / z2r splines, 8 floats per knot
  float4 *ff = nullptr;  // mixed mode: float4 Newton-scatter force array [nmax]
  EAMParams eam;
  EAMFast eam_one;            // packed r-space tables of a single-element potential
  DBuf<double> eam_one_d;
  bool eam_one_ok = false;    // B200_EAM_ONE=0 disables
  DBuf<int> eam_i;
  DBuf<double> eam_d;
  double *rho = nullptr, *fp = nullptr;
  // nve
  double dtv = 0, dtf = 0;
  int groupbit = 1;
  bool have_nve = false;
  bool pending_final = false;  // final_integrate of the last step is deferred (fused into the next initial)
  // fix nve fused into the pair kernel (k_tile_lj2<..., NVE>): on such a step the kernel applies
  // final_integrate(n) and initial_integrate(n+1) itself and writes the new positions to the
  // other xt buffer; `ahead` = the atoms already carry step n+1's initial_integrate.  Only inside
  // b200_run on steps nobody looks at (no tallies, not the last).  B200_FUSE=0 disables.
  bool fuse_nve = true, ahead = false;
  bool fuse_now = false;   // the pair launches of the step being enqueued carry the integrator
  int fuse_check = 0;      // ... and the displacement check for the next step's decide()
  int fuse_min_atoms = 0;  // (measured: fusing beats the CUDA graph of the unfused step even at 32 k atoms,
                           //  1.03e9 vs 9.3e8 atom-steps/s on bench/in.lj; B200_FUSE_MIN raises the threshold)
  // tallies / flags
  double *ev = nullptr;   // [8] device: eng, virial[6], ke
  double *ke7 = nullptr;  // [7] device: b200_ke_group accumulators
  int *flags = nullptr;   // [4] device: moved, err, maxcount, grand_total
  unsigned long long *cnt64 = nullptr;
  double *h_ev = nullptr; // pinned [24]: tallies [0,8), ke_group [8,15), langevin sums [16,19)
  int *h_flags = nullptr; // pinned [32]
  double eng_vdwl = 0, virial[6] = {0, 0, 0, 0, 0, 0};
  bool setup_done = false;
  // tallies stay per sub-domain (a host that reduces them itself: LAMMPS + MPI)
  bool local_tallies = false;
  cudaEvent_t run_a = nullptr, run_b = nullptr;
  double last_run_ms = 0;
  // profiling
  bool profiling = false;
  std::vector<PhaseRec> recs;
  std::vector<cudaEvent_t> evpool;
  double ph_ms[B200_NPHASE] = {0};
  int64_t ph_calls[B200_NPHASE] = {0};

  int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return ctx->fail(B200_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                       __LINE__);                                                         \
  } while (0)
#define TRY(call)            \
  do {                       \
    int r_ = (call);         \
    if (r_ != B200_OK) return r_; \
  } while (0)
#define LAUNCH_CHECK() CK(cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <class T>
static int dalloc(b200_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc((void **)p, n * sizeof(T)));
  ctx->dev_bytes += n * sizeof(T);
  return B200_OK;
}
template <class T>
static void dfree(b200_ctx *ctx, T *&p, size_t n) {
  if (p) {
    cudaFree(p);
    ctx->dev_bytes -= std::min(ctx->dev_bytes, (n ? n : 1) * sizeof(T));
    p = nullptr;
  }
}
// grow-only buffer, contents NOT preserved
template <class T>
static int reserve(b200_ctx *ctx, DBuf<T> &b, size_t n) {
  if (n <= b.cap && b.p) return B200_OK;
  size_t cap = n + n / 8 + 64;
  if (b.p) dfree(ctx, b.p, b.cap);
  b.cap = 0;
  TRY(dalloc(ctx, &b.p, cap));
  b.cap = cap;
  return B200_OK;
}

// ------------------------------------------------------------------ profiling helpers
static int ph_begin(b200_ctx *ctx, int ph) {
  if (!ctx->profiling) return -1;
  PhaseRec r;
  r.ph = ph;
  auto get = [&]() {
    cudaEvent_t e;
    if (!ctx->evpool.empty()) {
      e = ctx->evpool.back();
      ctx->evpool.pop_back();
    } else
      cudaEventCreate(&e);
    return e;
  };
  r.a = get();
  r.b = get();
  cudaEventRecord(r.a, ctx->stream);
  ctx->recs.push_back(r);
  return (int)ctx->recs.size() - 1;
}
static void ph_end(b200_ctx *ctx, int h) {
  if (h < 0 || h >= (int)ctx->recs.size()) return;
  cudaEventRecord(ctx->recs[h].b, ctx->stream);
}
static void ph_collect(b200_ctx *ctx) {
  if (ctx->recs.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto &r : ctx->recs) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ctx->ph_ms[r.ph] += ms;
    ctx->ph_calls[r.ph]++;
    ctx->evpool.push_back(r.a);
    ctx->evpool.push_back(r.b);
  }
  ctx->recs.clear();
}


// ------------------------------------------------------------------ NCCL (loaded on demand)
// libnccl.so.2 is dlopen'ed only when b200_comm_init is called, so a single-GPU process has no
// NCCL dependency.  Inside a process that already loaded NCCL (torch) the same copy is reused.
namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
} g_nccl;

bool load_nccl(std::string &why) {
  if (g_nccl.ok) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    why = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
#define SYM(field, name)                                               \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);              \
  if (!g_nccl.field) {                                                 \
    why = std::string("libnccl lacks ") + name;                        \
    return false;                                                      \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(AllGather, "ncclAllGather")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.ok = true;
  return true;
}
}  // namespace

#define NK(call)                                                                           \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != ncclSuccess)                                                                 \
      return ctx->fail(B200_ECUDA, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, \
                       __LINE__);                                                          \
  } while (0)

// One grouped exchange between the 26 neighbours.  forward: segment `dir` of the send-order
// buffer goes to nbr[dir]; segment `dir` of the recv-order buffer comes from nbr[26-dir] (the
// rank that sent in direction dir).  reverse: the same segments travel the other way.  Both
// sides walk the directions in ascending order, so several messages between one pair of ranks
// (2 ranks along a dimension) match up in issue order.
static int halo_exchange(b200_ctx *ctx, const double *src, const int *srcoff, double *dst,
                         const int *dstoff, int width, bool reverse) {
  if (ctx->nranks == 1) return B200_OK;
  if (ctx->grp) {
    // in-process: publish where my segments are, then pull mine from the neighbours' buffers
    b200_group *g = ctx->grp;
    b200_group::Pub &me = g->pub[ctx->rank];
    me.src = src;
    memcpy(me.off, srcoff, sizeof me.off);
    CK(cudaEventRecord(me.ready, ctx->stream));
    g->barrier();
    for (int dir = 0; dir < NDIR; dir++) {
      if (!((ctx->remote_mask >> dir) & 1u)) continue;
      const int from = reverse ? ctx->nbr[dir] : ctx->nbr[NDIR - 1 - dir];
      const size_t nr = (size_t)(dstoff[dir + 1] - dstoff[dir]) * width;
      if (!nr || from < 0) continue;
      const b200_group::Pub &p = g->pub[from];
      CK(cudaStreamWaitEvent(ctx->stream, p.ready, 0));
      CK(cudaMemcpyAsync(dst + (size_t)dstoff[dir] * width, p.src + (size_t)p.off[dir] * width,
                         nr * sizeof(double), cudaMemcpyDefault, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    g->barrier();  // every pull is done: the source buffers may be reused
    return B200_OK;
  }
  if (!ctx->remote_mask) return B200_OK;
  NK(g_nccl.GroupStart());
  for (int dir = 0; dir < NDIR; dir++) {
    if (!((ctx->remote_mask >> dir) & 1u)) continue;
    const int to = reverse ? ctx->nbr[NDIR - 1 - dir] : ctx->nbr[dir];
    const int from = reverse ? ctx->nbr[dir] : ctx->nbr[NDIR - 1 - dir];
    const size_t ns = (size_t)(srcoff[dir + 1] - srcoff[dir]) * width;
    const size_t nr = (size_t)(dstoff[dir + 1] - dstoff[dir]) * width;
    if (ns && to >= 0)
      NK(g_nccl.Send(src + (size_t)srcoff[dir] * width, ns, ncclDouble, to, ctx->nccl, ctx->stream));
    if (nr && from >= 0)
      NK(g_nccl.Recv(dst + (size_t)dstoff[dir] * width, nr, ncclDouble, from, ctx->nccl, ctx->stream));
  }
  NK(g_nccl.GroupEnd());
  return B200_OK;
}

// The 32 device counters of every rank, on the host: [r*32 + dir] and the error word [r*32+27].
// One all-gather + one D2H copy + one stream sync (rebuild steps only).
static int sync_counts(b200_ctx *ctx) {
  if (ctx->grp) {
    b200_group *g = ctx->grp;
    CK(cudaMemcpyAsync(ctx->h_counts + 32 * (size_t)ctx->nranks, ctx->counts, sizeof(int) * 32,
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_flags + 40, ctx->p2p_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(&g->counts[(size_t)ctx->rank * 32], ctx->h_counts + 32 * (size_t)ctx->nranks, sizeof(int) * 32);
    g->barrier();
    memcpy(ctx->h_counts, g->counts.data(), sizeof(int) * 32 * ctx->nranks);
    g->barrier();
    if (ctx->h_flags[40]) return ctx->fail(B200_ECUDA, "peer-memory halo timed out waiting for a neighbour");
    return B200_OK;
  }
  if (ctx->nranks > 1) {
    NK(g_nccl.AllGather(ctx->counts, ctx->allcounts, 32, ncclInt, ctx->nccl, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->allcounts, sizeof(int) * 32 * ctx->nranks,
                       cudaMemcpyDeviceToHost, ctx->stream));
  } else
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->counts, sizeof(int) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->p2p)
    CK(cudaMemcpyAsync(ctx->h_flags + 40, ctx->p2p_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->p2p && ctx->h_flags[40])
    return ctx->fail(B200_ECUDA, "peer-memory halo timed out waiting for a neighbour");
  return B200_OK;
}
static int global_err(const b200_ctx *ctx) {
  int e = 0;
  for (int r = 0; r < ctx->nranks; r++) e |= ctx->h_counts[r * 32 + 27];
  return e;
}


// ------------------------------------------------------------------ peer-memory halo (host side)
namespace {
struct PeerInfo {  // what every rank publishes at a border build (all-gathered, 512 bytes)
  int sendoff[NDIR + 1], recvoff[NDIR + 1];
  unsigned gen;
  int ok;
  unsigned long long arena_delta;  // arena pointer minus the base of its cudaMalloc allocation
  unsigned long long roff;         // byte offset of rbuf inside the arena
  cudaIpcMemHandle_t arena;
  char pad[512 - 2 * (NDIR + 1) * 4 - 8 - 16 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(PeerInfo) == 512, "PeerInfo is one 512-byte record");

// offset of p inside the cudaMalloc allocation that contains it (IPC handles name allocations)
size_t alloc_delta(const void *p) {
  typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
  static range_fn fn = nullptr;
  if (!fn) {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess) fn = (range_fn)f;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (fn && fn(&base, &size, (unsigned long long)(uintptr_t)p) == 0) return (size_t)((uintptr_t)p - base);
  return 0;
}
}  // namespace

// all-gather `bytes` per rank from a host record (through device staging) back to the host
// (staging buffers are allocated once in b200_comm_init: a cudaMalloc/cudaFree pair per call
// used to put a device-wide synchronisation into every rebuild)
#define AG_BYTES 512
static int allgather_host(b200_ctx *ctx, const void *mine, void *all, size_t bytes) {
  if (bytes > AG_BYTES || !ctx->ag_dev) return ctx->fail(B200_EARG, "allgather_host: record of %zu bytes", bytes);
  char *stage = ctx->ag_dev, *hst = ctx->ag_host;
  memcpy(hst, mine, bytes);
  CK(cudaMemcpyAsync(stage, hst, bytes, cudaMemcpyHostToDevice, ctx->stream));
  NK(g_nccl.AllGather(stage, stage + AG_BYTES, bytes, ncclChar, ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(hst + AG_BYTES, stage + AG_BYTES, bytes * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(all, hst + AG_BYTES, bytes * ctx->nranks);
  return B200_OK;
}

// once, at b200_comm_init: flags + counters, IPC-map every rank's flags; decides ctx->p2p
static int p2p_init(b200_ctx *ctx) {
  const char *mode = getenv("B200_HALO");
  const bool want = !(mode && strcmp(mode, "nccl") == 0);
  const size_t fbytes = 2u << 20;  // own allocation (>= 1 MiB) so the IPC handle names just it
  CK(cudaMalloc((void **)&ctx->pflags, fbytes));
  CK(cudaMemset(ctx->pflags, 0, fbytes));
  CK(cudaMalloc((void **)&ctx->p2p_counter, 8 * sizeof(unsigned)));  // [0..3] counters, [4] error word
  CK(cudaMemset(ctx->p2p_counter, 0, 8 * sizeof(unsigned)));
  struct FlagInfo {
    cudaIpcMemHandle_t h;
    unsigned long long delta;
    int ok, pad;
  } mine, *all;
  memset(&mine, 0, sizeof mine);
  mine.ok = want && cudaIpcGetMemHandle(&mine.h, ctx->pflags) == cudaSuccess;
  mine.delta = alloc_delta(ctx->pflags);
  std::vector<FlagInfo> allv(ctx->nranks);
  all = allv.data();
  TRY(allgather_host(ctx, &mine, all, sizeof(FlagInfo)));
  ctx->peer_flags.assign(ctx->nranks, nullptr);
  ctx->peer_arena.assign(ctx->nranks, nullptr);
  ctx->peer_base.assign(ctx->nranks, nullptr);
  ctx->peer_roff.assign(ctx->nranks, 0);
  ctx->peer_gen.assign(ctx->nranks, 0);
  ctx->peer_off.assign((size_t)ctx->nranks * 2 * (NDIR + 1), 0);
  ctx->peer_cap_s.assign(ctx->nranks, 0);
  ctx->peer_cap_r.assign(ctx->nranks, 0);
  int ok = 1;
  for (int r = 0; r < ctx->nranks; r++) ok &= all[r].ok;
  if (ok)
    for (int r = 0; r < ctx->nranks && ok; r++) {
      if (r == ctx->rank) {
        ctx->peer_flags[r] = ctx->pflags;
        continue;
      }
      void *m = nullptr;
      if (cudaIpcOpenMemHandle(&m, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      ctx->peer_flags[r] = (long long *)((char *)m + all[r].delta);
    }
  // every rank must take the same transport
  int *d_ok = ctx->counts + 28;
  CK(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NK(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->p2p = ok != 0;
  return B200_OK;
}

// sbuf/rbuf capacity (doubles) inside the arena; (re)allocates when too small
// the growth rule, shared by the owner of an arena and by everybody tracking it: capacities in
// doubles for ns / nr doubles of send / recv staging; returns true when the arena must move
static bool arena_rule(size_t ns, size_t nr, size_t &cap_s, size_t &cap_r) {
  if (cap_s >= ns && cap_r >= nr && cap_s > 0) return false;
  const size_t need_s = (ns * sizeof(double) + 4095) / 4096 * 4096;
  const size_t need_r = (nr * sizeof(double) + 4095) / 4096 * 4096;
  cap_s = (need_s + need_s / 2 + (1u << 20)) / sizeof(double);
  cap_r = (need_r + need_r / 2 + (1u << 20)) / sizeof(double);
  return true;
}

static int ensure_arena(b200_ctx *ctx, size_t ns, size_t nr) {
  size_t cs = ctx->arena ? ctx->sbuf.cap : 0, cr = ctx->arena ? ctx->rbuf.cap : 0;
  if (!arena_rule(ns, nr, cs, cr)) return B200_OK;
  const size_t cap_s = cs * sizeof(double), cap_r = cr * sizeof(double);
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->arena) {
    cudaFree(ctx->arena);
    ctx->dev_bytes -= std::min(ctx->dev_bytes, ctx->arena_bytes);
  }
  ctx->arena_bytes = cap_s + cap_r;
  CK(cudaMalloc((void **)&ctx->arena, ctx->arena_bytes));
  ctx->dev_bytes += ctx->arena_bytes;
  ctx->arena_roff = cap_s;
  ctx->sbuf.p = (double *)ctx->arena;
  ctx->sbuf.cap = cap_s / sizeof(double);
  ctx->rbuf.p = (double *)(ctx->arena + cap_s);
  ctx->rbuf.cap = cap_r / sizeof(double);
  ctx->arena_gen++;
  return B200_OK;
}

// At every border build.  The per-direction offsets of EVERY rank follow from the all-gathered
// border counts (h_counts), so nothing but the counts crosses the wire on an ordinary rebuild;
// only when some rank's arena moved (all ranks know: arena_rule is deterministic in the counts)
// are the CUDA IPC handles exchanged again and the moved arenas re-mapped.  Then the per-direction
// tables the pack/unpack kernels take by value are laid out.
static int p2p_publish_borders(b200_ctx *ctx) {
  if (!ctx->p2p) return B200_OK;
  const int W2 = 2 * (NDIR + 1), nr = ctx->nranks;
  bool any_moved = false;
  for (int r = 0; r < nr; r++) {
    const int *nb = &ctx->rank_nbr[(size_t)r * NDIR];
    int *so = &ctx->peer_off[(size_t)r * W2], *ro = so + NDIR + 1;
    so[0] = ro[0] = 0;
    for (int dir = 0; dir < NDIR; dir++) {
      const bool rem = dir != 13 && nb[dir] >= 0 && nb[dir] != r;
      const int from = nb[NDIR - 1 - dir];
      const int mine = ctx->h_counts[r * 32 + dir];
      const int theirs = rem ? (from >= 0 ? ctx->h_counts[from * 32 + dir] : 0) : mine;
      so[dir + 1] = so[dir] + mine;
      ro[dir + 1] = ro[dir] + theirs;
    }
    any_moved |= arena_rule((size_t)so[NDIR] * 4, (size_t)ro[NDIR] * 4, ctx->peer_cap_s[r], ctx->peer_cap_r[r]);
  }
  if (ctx->peer_cap_s[ctx->rank] != ctx->sbuf.cap || ctx->peer_cap_r[ctx->rank] != ctx->rbuf.cap)
    return ctx->fail(B200_ECUDA, "halo arena bookkeeping out of step (%zu/%zu tracked, %zu/%zu allocated)",
                     ctx->peer_cap_s[ctx->rank], ctx->peer_cap_r[ctx->rank], ctx->sbuf.cap, ctx->rbuf.cap);
  std::vector<PeerInfo> all;
  if (ctx->grp) {
    // in-process: the neighbours' arenas are plain pointers; re-read them when anything moved
    b200_group *g = ctx->grp;
    g->barrier();  // every context has (re)allocated its arena for this build
    for (int r = 0; r < nr; r++) {
      ctx->peer_arena[r] = g->ctx[r]->arena;
      ctx->peer_roff[r] = g->ctx[r]->arena_roff;
    }
    any_moved = false;
  }
  if (any_moved) {
    PeerInfo mine;
    memset(&mine, 0, sizeof mine);
    mine.gen = ctx->arena_gen;
    mine.roff = ctx->arena_roff;
    mine.arena_delta = alloc_delta(ctx->arena);
    mine.ok = cudaIpcGetMemHandle(&mine.arena, ctx->arena) == cudaSuccess;
    all.resize(nr);
    TRY(allgather_host(ctx, &mine, all.data(), sizeof(PeerInfo)));
    for (int r = 0; r < nr; r++)
      if (!all[r].ok) return ctx->fail(B200_ECUDA, "rank %d cannot export its halo arena through CUDA IPC", r);
  }
  // map (or re-map) the arenas of the ranks I talk to
  for (int dir = 0; dir < NDIR && any_moved; dir++) {
    if (!((ctx->remote_mask >> dir) & 1u)) continue;
    const int r = ctx->nbr[dir];
    if (r < 0 || r == ctx->rank) continue;
    if (ctx->peer_arena[r] && ctx->peer_gen[r] == all[r].gen) continue;
    if (ctx->peer_base[r]) {
      cudaIpcCloseMemHandle(ctx->peer_base[r]);
      ctx->peer_base[r] = nullptr;
      ctx->peer_arena[r] = nullptr;
    }
    void *m = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&m, all[r].arena, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return ctx->fail(B200_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
    ctx->peer_base[r] = m;
    ctx->peer_arena[r] = (char *)m + all[r].arena_delta;
    ctx->peer_gen[r] = all[r].gen;
    ctx->peer_roff[r] = (size_t)all[r].roff;
  }
  // tables
  P2PMap &F = ctx->fwdP, &R = ctx->revP;
  memset(&F, 0, sizeof F);
  memset(&R, 0, sizeof R);
  long long *own = ctx->pflags;
  F.flag_in = own + 0 * 32;  // arrival F
  R.flag_in = own + 1 * 32;  // arrival R
  F.ack_in = own + 2 * 32;   // ack F
  R.ack_in = own + 3 * 32;   // ack R
  for (int dir = 0; dir < NDIR; dir++) {
    if (!((ctx->remote_mask >> dir) & 1u)) continue;
    const int to = ctx->nbr[dir], from = ctx->nbr[NDIR - 1 - dir];
    const int nsend = ctx->sendoff[dir + 1] - ctx->sendoff[dir];
    const int nrecv = ctx->recvoff[dir + 1] - ctx->recvoff[dir];
    if (nsend > 0 && to >= 0) {
      // forward: my send segment dir lands in `to`'s rbuf at its recvoff[dir]
      F.out_mask |= 1u << dir;
      F.dst[dir] = (double *)(ctx->peer_arena[to] + ctx->peer_roff[to]);
      F.dstoff[dir] = ctx->peer_off[(size_t)to * W2 + NDIR + 1 + dir];
      F.flag_out[dir] = ctx->peer_flags[to] + 0 * 32 + dir;
      // reverse: `to` returns that segment into my sbuf; I acknowledge to it
      R.in_mask |= 1u << dir;
      R.ack_out[dir] = ctx->peer_flags[to] + 3 * 32 + dir;
    }
    if (nrecv > 0 && from >= 0) {
      // forward: I receive segment dir from `from` and acknowledge to it
      F.in_mask |= 1u << dir;
      F.ack_out[dir] = ctx->peer_flags[from] + 2 * 32 + dir;
      // reverse: my recv segment dir goes back into `from`'s sbuf at its sendoff[dir]
      R.out_mask |= 1u << dir;
      R.dst[dir] = (double *)ctx->peer_arena[from];
      R.dstoff[dir] = ctx->peer_off[(size_t)from * W2 + dir];
      R.flag_out[dir] = ctx->peer_flags[from] + 1 * 32 + dir;
    }
  }
  return B200_OK;
}

// ------------------------------------------------------------------ per-atom storage
static int alloc_atoms(b200_ctx *ctx, int nmax) {
  // (re)allocate every per-atom array for nmax atoms; live data is copied over
  const int old = ctx->nmax, keep = ctx->nlocal + ctx->nghost, c = ctx->cur;
  auto regrow = [&](auto *&p, int n_new, int n_keep, size_t elem) -> int {
    using T = std::remove_reference_t<decltype(*p)>;
    T *q = nullptr;
    TRY(dalloc(ctx, &q, (size_t)n_new));
    if (p && n_keep > 0)
      CK(cudaMemcpyAsync(q, p, (size_t)n_keep * elem, cudaMemcpyDeviceToDevice, ctx->stream));
    if (p) {
      CK(cudaStreamSynchronize(ctx->stream));
      dfree(ctx, p, (size_t)old);
    }
    p = q;
    return B200_OK;
  };
  for (int b = 0; b < 2; b++) {
    const int k = (b == c) ? keep : 0;
    TRY(regrow(ctx->xt[b], nmax, k, sizeof(double4)));
    if (ctx->prec == B200_PREC_MIXED) {
      TRY(regrow(ctx->qrec[b], nmax, 0, sizeof(int4)));
      ctx->q_owned_valid = ctx->q_ghost_valid = false;
    }
    for (int d = 0; d < 3; d++) TRY(regrow(ctx->v[b][d], nmax, k, sizeof(double)));
    TRY(regrow(ctx->tag[b], nmax, k, sizeof(int)));
    TRY(regrow(ctx->mask[b], nmax, k, sizeof(int)));
    TRY(regrow(ctx->image[b], nmax, k, sizeof(int)));
    TRY(regrow(ctx->atombin[b], nmax, k, sizeof(int)));
  }
  for (int d = 0; d < 3; d++) {
    TRY(regrow(ctx->f[d], nmax, keep, sizeof(double)));
    TRY(regrow(ctx->xh[d], nmax, keep, sizeof(double)));
  }
  TRY(regrow(ctx->slot, nmax, keep, sizeof(int)));
  TRY(regrow(ctx->ff, nmax, 0, sizeof(float4)));
  TRY(regrow(ctx->rho, nmax, keep, sizeof(double)));
  TRY(regrow(ctx->fp, nmax, keep, sizeof(double)));
  ctx->nmax = nmax;
  return B200_OK;
}

// ------------------------------------------------------------------ geometry (host, once)
// Neighbor::init cutoffs (neighbor.cpp:337-383), CommBrick::setup slabs (comm_brick.cpp:389-420),
// NBinStandard::setup_bins (nbin_standard.cpp:82-214), NStencilBin<1,1,0>::create
// (nstencil.cpp:203-237, nstencil_bin.cpp:28-67) regrouped into x-contiguous rows.
static int setup_geometry(b200_ctx *ctx) {
  if (!ctx->have_box) return ctx->fail(B200_EARG, "b200_set_box has not been called");
  if (!ctx->pair_style) return ctx->fail(B200_EARG, "no pair style set");
  const int n = ctx->ntypes, n1 = n + 1;
  if ((int)ctx->cutsq_h.size() != n1 * n1)
    return ctx->fail(B200_EARG, "pair style was set for %d types but atoms have %d",
                     (int)std::lround(std::sqrt((double)ctx->cutsq_h.size())) - 1, n);
  ctx->triggersq = 0.25 * ctx->skin * ctx->skin;
  ctx->cutneighsq_h.assign(n1 * n1, 0.0);
  ctx->cutneighmax = 0.0;
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      double cutoff = std::sqrt(ctx->cutsq_h[i * n1 + j]);
      double delta = cutoff > 0.0 ? ctx->skin : 0.0;
      double cut = cutoff + delta;
      ctx->cutneighsq_h[i * n1 + j] = cut * cut;
      ctx->cutneighmax = std::max(ctx->cutneighmax, cut);
    }
  ctx->cutneighmaxsq = ctx->cutneighmax * ctx->cutneighmax;
  ctx->cutghost = ctx->cutneighmax;  // Comm::get_comm_cutoff, comm.cpp:683
  // neigh_modify exclude type (NPair::exclusion, npair.cpp:244-248): an excluded type pair is
  // never stored.  The build kernels test `rsq <= cutneighsq[itype][jtype]`; a negative entry
  // makes that test fail for every distance, so the exclusion costs no extra instruction.  (Bins,
  // stencil and ghost cutoff keep using the unmodified maximum, as the reference does.)
  if ((int)ctx->ex_type.size() == n1 * n1)
    for (int i = 1; i <= n; i++)
      for (int j = 1; j <= n; j++)
        if (ctx->ex_type[i * n1 + j]) ctx->cutneighsq_h[i * n1 + j] = -1.0;
  TRY(reserve(ctx, ctx->cutneighsq_d, (size_t)n1 * n1));
  CK(cudaMemcpyAsync(ctx->cutneighsq_d.p, ctx->cutneighsq_h.data(), sizeof(double) * n1 * n1,
                     cudaMemcpyHostToDevice, ctx->stream));

  Geom &g = ctx->geom;
  memset(&g, 0, sizeof g);
  // the comm frame (see Geom): the box itself, or the unit cube of lamda coordinates
  // (Domain::set_lamda_box, domain.cpp: boxlo_lamda = 0, boxhi_lamda = 1, prd_lamda = 1)
  const bool tri = ctx->tri;
  double flo[3], fhi[3], fprd[3], fcut[3];
  g.tri = tri ? 1 : 0;
  {
    // Domain::set_global_box, domain.cpp:263-290
    double *h = g.h, *hi = g.h_inv;
    h[0] = ctx->prd[0]; h[1] = ctx->prd[1]; h[2] = ctx->prd[2];
    hi[0] = 1.0 / h[0]; hi[1] = 1.0 / h[1]; hi[2] = 1.0 / h[2];
    h[3] = tri ? ctx->yz : 0.0; h[4] = tri ? ctx->xz : 0.0; h[5] = tri ? ctx->xy : 0.0;
    hi[3] = -h[3] / (h[1] * h[2]);
    hi[4] = (h[3] * h[5] - h[1] * h[4]) / (h[0] * h[1] * h[2]);
    hi[5] = -h[5] / (h[0] * h[1]);
    // CommBrick::setup, comm_brick.cpp:225-237: the ghost cutoff as a distance between lamda planes
    const double len[3] = {std::sqrt(hi[0] * hi[0] + hi[5] * hi[5] + hi[4] * hi[4]),
                           std::sqrt(hi[1] * hi[1] + hi[3] * hi[3]), hi[2]};
    for (int d = 0; d < 3; d++) {
      g.origin[d] = ctx->boxlo[d];
      flo[d] = tri ? 0.0 : ctx->boxlo[d];
      fhi[d] = tri ? 1.0 : ctx->boxhi[d];
      fprd[d] = tri ? 1.0 : ctx->prd[d];
      fcut[d] = tri ? ctx->cutghost * len[d] : ctx->cutghost;
    }
  }
  if (tri && ctx->box_changes) return ctx->fail(B200_EARG, "a changing triclinic box is not supported");
  for (int d = 0; d < 3; d++) {
    g.boxlo[d] = flo[d];
    g.boxhi[d] = fhi[d];
    g.prd[d] = fprd[d];
    g.periodic[d] = ctx->periodic[d];
    // Domain::set_local_box / set_lamda_box (domain.cpp), uniform grid: xsplit[i] = i * 1.0/procgrid
    const int P = ctx->procgrid[d], me = ctx->myloc[d];
    ctx->sublo[d] = flo[d] + fprd[d] * (me * 1.0 / P);
    ctx->subhi[d] = (me < P - 1) ? flo[d] + fprd[d] * ((me + 1) * 1.0 / P) : fhi[d];
    if (tri) {  // set_lamda_box: sublo_lamda = xsplit[myloc], subhi_lamda = xsplit[myloc+1] (= 1.0 at the end)
      ctx->sublo[d] = me * 1.0 / P;
      ctx->subhi[d] = (me < P - 1) ? (me + 1) * 1.0 / P : 1.0;
    }
    g.sublo[d] = ctx->sublo[d];
    g.subhi[d] = ctx->subhi[d];
    // comm_brick.cpp:268-270 maxneed; only one layer of neighbours is supported
    const double sub = ctx->subhi[d] - ctx->sublo[d];
    int maxneed = (int)(fcut[d] * P / fprd[d]) + 1;
    if (!ctx->periodic[d]) maxneed = std::min(maxneed, P - 1);
    if (maxneed > 1)
      return ctx->fail(B200_EARG,
                       "sub-domain edge %g in dim %d is shorter than the ghost cutoff %g "
                       "(multi-layer halos are not supported)", sub, d, fcut[d]);
    g.slab_left_hi[d] = ctx->sublo[d] + fcut[d];
    g.slab_right_lo[d] = ctx->subhi[d] - fcut[d];
    g.send_left[d] = (maxneed >= 1) && (ctx->periodic[d] || me > 0);
    g.send_right[d] = (maxneed >= 1) && (ctx->periodic[d] || me < P - 1);
  }
  for (int dir = 0; dir < NDIR; dir++) {
    const int dv[3] = {dir % 3 - 1, (dir / 3) % 3 - 1, dir / 9 - 1};
    int pbc[3];  // pbc[iswap][dim], comm_brick.cpp:389-420: set when the sender sits at the box edge
    for (int d = 0; d < 3; d++) {
      pbc[d] = 0;
      if (dv[d] < 0 && ctx->myloc[d] == 0) pbc[d] = 1;
      if (dv[d] > 0 && ctx->myloc[d] == ctx->procgrid[d] - 1) pbc[d] = -1;
      g.shift[dir][d] = pbc[d] * fprd[d];        // border time: box or lamda units (atom_vec.cpp:796-830)
      g.fshift[dir][d] = pbc[d] * ctx->prd[d];   // forward halo: box units (atom_vec.cpp:354-440)
    }
    // the y swap carries pbc[5] = pbc[1] (xy), the z swap pbc[4] = pbc[3] = pbc[2] (xz, yz)
    g.tilt[dir][0] = tri ? pbc[1] * ctx->xy : 0.0;
    g.tilt[dir][1] = tri ? pbc[2] * ctx->xz : 0.0;
    g.tilt[dir][2] = tri ? pbc[2] * ctx->yz : 0.0;
  }
  // ---- neighbour sub-domains: rank numbering as MPI_Cart (last dimension fastest,
  //      procmap.cpp:361-374), periodic wrap of the grid location, -1 beyond an open boundary
  {
    const int *P = ctx->procgrid, *me = ctx->myloc;
    int np = P[0] * P[1] * P[2];
    if (np != ctx->nranks)
      return ctx->fail(B200_EARG, "decomposition %dx%dx%d needs %d ranks but the communicator has %d",
                       P[0], P[1], P[2], np, ctx->nranks);
    b200_neighbor_ranks(P, me, ctx->periodic, ctx->nbr);
    if (!ctx->rankmap.empty()) {  // host-supplied grid -> rank map (Comm::grid2proc)
      if ((int)ctx->rankmap.size() != np)
        return ctx->fail(B200_EARG, "rank grid has %d entries, decomposition needs %d",
                         (int)ctx->rankmap.size(), np);
      for (int dir = 0; dir < NDIR; dir++)
        if (ctx->nbr[dir] >= 0) ctx->nbr[dir] = ctx->rankmap[ctx->nbr[dir]];
    }
    ctx->remote_mask = 0;
    for (int dir = 0; dir < NDIR; dir++)
      if (dir != 13 && ctx->nbr[dir] >= 0 && ctx->nbr[dir] != ctx->rank) ctx->remote_mask |= 1u << dir;
    // the neighbour table of every rank (p2p_publish_borders derives all ranks' offsets from it)
    ctx->rank_nbr.assign((size_t)np * NDIR, -1);
    for (int gidx = 0; gidx < np; gidx++) {
      const int loc[3] = {gidx / (P[1] * P[2]), (gidx / P[2]) % P[1], gidx % P[2]};
      const int r = ctx->rankmap.empty() ? gidx : ctx->rankmap[gidx];
      if (r < 0 || r >= np) return ctx->fail(B200_EARG, "rank grid entry %d out of range", r);
      int nb[NDIR];
      b200_neighbor_ranks(P, loc, ctx->periodic, nb);
      for (int dir = 0; dir < NDIR; dir++)
        ctx->rank_nbr[(size_t)r * NDIR + dir] =
            nb[dir] < 0 ? -1 : (ctx->rankmap.empty() ? nb[dir] : ctx->rankmap[nb[dir]]);
    }
    Owner &o = ctx->owner;
    memset(&o, 0, sizeof o);
    auto bounds = [&](int d, int l, double &lo, double &hi) {
      if (tri) {
        lo = l * 1.0 / P[d];
        hi = (l < P[d] - 1) ? (l + 1) * 1.0 / P[d] : 1.0;
        return;
      }
      lo = ctx->boxlo[d] + ctx->prd[d] * (l * 1.0 / P[d]);
      hi = (l < P[d] - 1) ? ctx->boxlo[d] + ctx->prd[d] * ((l + 1) * 1.0 / P[d]) : ctx->boxhi[d];
    };
    for (int d = 0; d < 3; d++) {
      bounds(d, me[d], o.lo[d][0], o.hi[d][0]);
      o.has[d][0] = 1;
      for (int side = 1; side <= 2; side++) {
        int l = me[d] + (side == 1 ? -1 : 1);
        if (l < 0 || l >= P[d]) {
          if (!ctx->periodic[d]) continue;
          l = (l + P[d]) % P[d];
        }
        if (l == me[d]) continue;  // one rank along this dimension: nobody else can own it
        bounds(d, l, o.lo[d][side], o.hi[d][side]);
        o.has[d][side] = 1;
      }
    }
  }
  // ---- bins
  double bbox[3], bsublo[3], bsubhi[3], binsize[3];
  for (int d = 0; d < 3; d++) {
    bsublo[d] = ctx->sublo[d] - ctx->cutghost;
    bsubhi[d] = ctx->subhi[d] + ctx->cutghost;
    g.binlo[d] = ctx->boxlo[d];
    g.binhi[d] = ctx->boxhi[d];
  }
  if (tri) {
    // NBin::bboxlo/hi = Domain::boxlo_bound / boxhi_bound (domain.cpp:282-290); the sub-domain's
    // extent = bounding box of its lamda brick +- cutghost (Domain::bbox, domain.cpp:2467-2530)
    g.binlo[0] = std::min(ctx->boxlo[0], ctx->boxlo[0] + ctx->xy);
    g.binlo[0] = std::min(g.binlo[0], g.binlo[0] + ctx->xz);
    g.binlo[1] = std::min(ctx->boxlo[1], ctx->boxlo[1] + ctx->yz);
    g.binlo[2] = ctx->boxlo[2];
    g.binhi[0] = std::max(ctx->boxhi[0], ctx->boxhi[0] + ctx->xy);
    g.binhi[0] = std::max(g.binhi[0], g.binhi[0] + ctx->xz);
    g.binhi[1] = std::max(ctx->boxhi[1], ctx->boxhi[1] + ctx->yz);
    g.binhi[2] = ctx->boxhi[2];
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = ctx->sublo[d] - fcut[d];
      hi[d] = ctx->subhi[d] + fcut[d];
      bsublo[d] = 1.0e20;
      bsubhi[d] = -1.0e20;
    }
    // corners in the order Domain::bbox visits them (min/max do not depend on it)
    for (int cz = 0; cz < 2; cz++)
      for (int cy = 0; cy < 2; cy++)
        for (int cx = 0; cx < 2; cx++) {
          const double l[3] = {cx ? hi[0] : lo[0], cy ? hi[1] : lo[1], cz ? hi[2] : lo[2]};
          const double x[3] = {g.h[0] * l[0] + g.h[5] * l[1] + g.h[4] * l[2] + ctx->boxlo[0],
                               g.h[1] * l[1] + g.h[3] * l[2] + ctx->boxlo[1], g.h[2] * l[2] + ctx->boxlo[2]};
          for (int d = 0; d < 3; d++) {
            bsublo[d] = std::min(bsublo[d], x[d]);
            bsubhi[d] = std::max(bsubhi[d], x[d]);
          }
        }
  }
  for (int d = 0; d < 3; d++) bbox[d] = g.binhi[d] - g.binlo[d];
  double binsize_optimal = 0.5 * ctx->cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  const double binsizeinv = 1.0 / binsize_optimal;
  int64_t mb = 1;
  for (int d = 0; d < 3; d++) {
    if (bbox[d] * binsizeinv > 2147483647.0) return ctx->fail(B200_EARG, "Domain too large for neighbor bins");
    g.nbin[d] = (int)(bbox[d] * binsizeinv);
    if (g.nbin[d] == 0) g.nbin[d] = 1;
    binsize[d] = bbox[d] / g.nbin[d];
    g.bininv[d] = 1.0 / binsize[d];
    double coord = bsublo[d] - B200_SMALL * bbox[d];
    int lo = (int)((coord - g.binlo[d]) * g.bininv[d]);
    if (coord < g.binlo[d]) lo = lo - 1;
    coord = bsubhi[d] + B200_SMALL * bbox[d];
    int hi = (int)((coord - g.binlo[d]) * g.bininv[d]);
    lo -= 1;
    hi += 1;
    g.mbinlo[d] = lo;
    g.mbin[d] = hi - lo + 1;
    mb *= g.mbin[d];
  }
  if (mb + 1 > 2147483647LL) return ctx->fail(B200_EARG, "Too many neighbor bins");
  g.mbins = (int)mb;
  // ---- stencil
  int s[3];
  for (int d = 0; d < 3; d++) {
    s[d] = (int)(ctx->cutneighmax * g.bininv[d]);
    if (s[d] * binsize[d] < ctx->cutneighmax) s[d]++;
  }
  auto bin_distance = [&](int i, int j, int k) {
    double delx, dely, delz;
    if (i > 0) delx = (i - 1) * binsize[0];
    else if (i == 0) delx = 0.0;
    else delx = (i + 1) * binsize[0];
    if (j > 0) dely = (j - 1) * binsize[1];
    else if (j == 0) dely = 0.0;
    else dely = (j + 1) * binsize[1];
    if (k > 0) delz = (k - 1) * binsize[2];
    else if (k == 0) delz = 0.0;
    else delz = (k + 1) * binsize[2];
    return delx * delx + dely * dely + delz * delz;
  };
  Stencil &st = ctx->stencil;
  memset(&st, 0, sizeof st);
  ctx->nstencil = 1;  // the central bin comes first (nstencil_bin.cpp:50)
  // row 0 is (dz=0,dy=0): own bin + bins to its right
  st.nrows = 1;
  st.rowoff[0] = 0;
  st.dxlo[0] = 0;
  st.dxhi[0] = 0;
  if (tri) {
    // NStencilBin<HALF=1,DIM_3D=1,TRI=1>::create (nstencil_bin.cpp:36-62): full in all three
    // dimensions, no separate central bin; every (dz,dy) row is a contiguous x range
    ctx->nstencil = 0;
    st.nrows = 0;
    for (int k = -s[2]; k <= s[2]; k++)
      for (int j = -s[1]; j <= s[1]; j++) {
        int lo = 1 << 30, hi = -(1 << 30), cnt = 0;
        for (int i = -s[0]; i <= s[0]; i++)
          if (bin_distance(i, j, k) < ctx->cutneighmaxsq) {
            lo = std::min(lo, i);
            hi = std::max(hi, i);
            cnt++;
          }
        if (!cnt) continue;
        if (cnt != hi - lo + 1) return ctx->fail(B200_EARG, "stencil row is not contiguous");
        if (st.nrows >= MAXROWS) return ctx->fail(B200_EARG, "too many stencil rows");
        ctx->nstencil += cnt;
        st.rowoff[st.nrows] = k * g.mbin[1] * g.mbin[0] + j * g.mbin[0];
        st.dxlo[st.nrows] = lo;
        st.dxhi[st.nrows] = hi;
        st.nrows++;
      }
  }
  for (int k = 0; !tri && k <= s[2]; k++)
    for (int j = -s[1]; j <= s[1]; j++) {
      int lo = 1 << 30, hi = -(1 << 30), cnt = 0;
      for (int i = -s[0]; i <= s[0]; i++) {
        if (k <= 0 && j <= 0 && (j != 0 || i <= 0)) continue;
        if (bin_distance(i, j, k) < ctx->cutneighmaxsq) {
          lo = std::min(lo, i);
          hi = std::max(hi, i);
          cnt++;
        }
      }
      if (!cnt) continue;
      if (cnt != hi - lo + 1) return ctx->fail(B200_EARG, "stencil row is not contiguous");
      ctx->nstencil += cnt;
      if (k == 0 && j == 0) {
        if (lo != 1) return ctx->fail(B200_EARG, "unexpected stencil row (0,0)");
        st.dxhi[0] = hi;
      } else {
        if (st.nrows >= MAXROWS) return ctx->fail(B200_EARG, "too many stencil rows");
        st.rowoff[st.nrows] = k * g.mbin[1] * g.mbin[0] + j * g.mbin[0];
        st.dxlo[st.nrows] = lo;
        st.dxhi[st.nrows] = hi;
        st.nrows++;
      }
    }
  // ---- bin tiles (kernels_tile.cuh): every (dy,dz) row of the stencil, both halves, and the
  //      range of local bins that can hold an owned atom (sublo <= x < subhi)
  {
    FullStencil &fs = ctx->fst;
    memset(&fs, 0, sizeof fs);
    bool ok = s[0] <= 3 && s[1] <= 3 && s[2] <= 3;
    for (int k = -s[2]; ok && k <= s[2]; k++)
      for (int j = -s[1]; j <= s[1]; j++) {
        int lo = 1 << 30, hi = -(1 << 30), cnt = 0;
        for (int i = -s[0]; i <= s[0]; i++)
          if (bin_distance(i, j, k) < ctx->cutneighmaxsq) {
            lo = std::min(lo, i);
            hi = std::max(hi, i);
            cnt++;
          }
        if (!cnt) continue;
        if (cnt != hi - lo + 1 || lo != -hi || fs.nrows >= FST_MAXROWS) {
          ok = false;
          break;
        }
        fs.dy[fs.nrows] = (signed char)j;
        fs.dz[fs.nrows] = (signed char)k;
        fs.dxlo[fs.nrows] = (signed char)lo;
        fs.dxhi[fs.nrows] = (signed char)hi;
        fs.nrows++;
      }
    ctx->use_tiles = ok;  // unusual stencil or triclinic box: the flat list handles it
    {
      // Order in which the build walks the stencil rows = order of the entries in a list row.  Any
      // order that does not put the rows of neighbouring z-planes next to each other makes the
      // lj/cut tile kernels ~8 % faster (687 -> 632 us at 4 M atoms, profiles/r02ad_probe_fst_order.txt):
      // the two entries a thread has in flight then come from different staged rows.  Default 9:
      // mirrored pairs (-dy,-dz),(dy,dz) from the centre row outwards; B200_FST_ORDER=0 keeps the
      // sorted order, 1..8 are the other permutations that were measured.
      const char *e = getenv("B200_FST_ORDER");
      const int mode = e ? atoi(e) : 9;
      FullStencil t = fs;
      std::vector<int> ord;
      const int n = fs.nrows, h = (n + 1) / 2;
      if (mode == 1) { for (int i = 0; i < h; i++) { ord.push_back(i); if (i + h < n) ord.push_back(i + h); } }
      else if (mode == 2) { for (int i = 0; i < n; i++) ord.push_back(n - 1 - i); }
      else if (mode == 3) { for (int st = 0; st < 5; st++) for (int i = st; i < n; i += 5) ord.push_back(i); }
      else if (mode == 4) { const int q = (n + 2) / 3; for (int i = 0; i < q; i++) for (int k = 0; k < 3; k++) if (i + k * q < n) ord.push_back(i + k * q); }
      else if (mode == 5) { for (int i = 0; i < n; i++) ord.push_back(i); unsigned x = 12345; for (int i = n - 1; i > 0; i--) { x = x * 1664525u + 1013904223u; std::swap(ord[i], ord[(x >> 8) % (i + 1)]); } }
      else if (mode == 6) { for (int i = 0; i < h; i++) { ord.push_back(i); if (n - 1 - i > i) ord.push_back(n - 1 - i); } }
      else if (mode == 7) { const int q = (n + 3) / 4; for (int i = 0; i < q; i++) for (int k = 0; k < 4; k++) if (i + k * q < n) ord.push_back(i + k * q); }
      else if (mode == 8) { for (int i = 0; i < h; i++) { if (i + h < n) ord.push_back(i + h); ord.push_back(i); } }
      else if (mode == 9) { for (int i = h - 1; i >= 0; i--) { ord.push_back(i); if (n - 1 - i > i) ord.push_back(n - 1 - i); } }
      if ((int)ord.size() == n)
        for (int i = 0; i < n; i++) {
          fs.dy[i] = t.dy[ord[i]]; fs.dz[i] = t.dz[ord[i]]; fs.dxlo[i] = t.dxlo[ord[i]]; fs.dxhi[i] = t.dxhi[ord[i]];
        }
    }
    auto host_bin = [&](double x, int d) {  // NBin::coord2bin, nbin.cpp:141-173
      int ix;
      if (x >= g.binhi[d]) ix = (int)((x - g.binhi[d]) * g.bininv[d]) + g.nbin[d];
      else if (x >= g.binlo[d]) ix = std::min((int)((x - g.binlo[d]) * g.bininv[d]), g.nbin[d] - 1);
      else ix = (int)((x - g.binlo[d]) * g.bininv[d]) - 1;
      return ix - g.mbinlo[d];
    };
    // extent of the sub-domain in box coordinates: the brick itself, or the bounding box of the
    // lamda brick (some of its bins then hold no owned atom, only ghosts: tiles with ni = 0)
    for (int d = 0; d < 3; d++) {
      ctx->osublo[d] = ctx->sublo[d];
      ctx->osubhi[d] = ctx->subhi[d];
    }
    if (tri) {
      for (int d = 0; d < 3; d++) {
        ctx->osublo[d] = 1.0e300;
        ctx->osubhi[d] = -1.0e300;
      }
      for (int cz = 0; cz < 2; cz++)
        for (int cy = 0; cy < 2; cy++)
          for (int cx = 0; cx < 2; cx++) {
            const double l[3] = {cx ? ctx->subhi[0] : ctx->sublo[0], cy ? ctx->subhi[1] : ctx->sublo[1],
                                 cz ? ctx->subhi[2] : ctx->sublo[2]};
            const double x[3] = {g.h[0] * l[0] + g.h[5] * l[1] + g.h[4] * l[2] + ctx->boxlo[0],
                                 g.h[1] * l[1] + g.h[3] * l[2] + ctx->boxlo[1], g.h[2] * l[2] + ctx->boxlo[2]};
            for (int d = 0; d < 3; d++) {
              ctx->osublo[d] = std::min(ctx->osublo[d], x[d]);
              ctx->osubhi[d] = std::max(ctx->osubhi[d], x[d]);
            }
          }
      // lamda2x of an owned atom may land an ulp outside the exact corners: one bin of slack
      for (int d = 0; d < 3; d++) {
        const double pad = 1.0e-9 * (g.binhi[d] - g.binlo[d]);
        ctx->osublo[d] -= pad;
        ctx->osubhi[d] += pad;
      }
    }
    for (int d = 0; d < 3; d++) {
      const int lo = std::max(host_bin(ctx->osublo[d], d), 0);
      const int hi = std::min(host_bin(tri ? ctx->osubhi[d] : std::nextafter(ctx->subhi[d], -1.0e300), d),
                              g.mbin[d] - 1);
      ctx->ibin_lo[d] = lo;
      ctx->ibin_n[d] = std::max(hi - lo + 1, 1);
      ctx->tg.s[d] = s[d];
    }
    ctx->tile_level = -1;
  }
  {
    // sub-domain-wide fixed point of the mixed lj/cut kernel: origin a safe margin outside the
    // ghost shell, 30 bits over the largest extent, power-of-two scale
    double ext = 0.0;
    const double margin = ctx->cutghost + 2.0 * ctx->skin;
    double qlo[3], qhi[3];
    for (int d = 0; d < 3; d++) {
      qlo[d] = ctx->osublo[d] - margin;
      qhi[d] = ctx->osubhi[d] + margin;
      if (tri) {
        // the ghost shell is cutghost wide between lamda planes: along a tilted direction its box
        // extent is larger than cutghost; bsublo/bsubhi = bounding box of the brick +- that shell
        qlo[d] = bsublo[d] - 2.0 * ctx->skin;
        qhi[d] = bsubhi[d] + 2.0 * ctx->skin;
      }
      ext = std::max(ext, qhi[d] - qlo[d]);
    }
    ctx->qgeom.ox = qlo[0];
    ctx->qgeom.oy = qlo[1];
    ctx->qgeom.oz = qlo[2];
    ctx->qgeom.scale = std::ldexp(1.0, (int)std::floor(std::log2(1073741824.0 / ext)));
    double cmin = 1.0e300;
    for (int i = 1; i <= n; i++)
      for (int j = 1; j <= n; j++)
        if (ctx->cutsq_h[i * n1 + j] > 0.0) cmin = std::min(cmin, ctx->cutsq_h[i * n1 + j]);
    // grid spacing against the 2e-6 decision band of the kernel (see k_tile_lj2f)
    ctx->gq_ok = ctx->prec == B200_PREC_MIXED && ctx->pair_style == 1 && cmin < 1.0e300 &&
                 1.0 / ctx->qgeom.scale <= 1.4e-7 * std::sqrt(cmin);
    if (const char *e = getenv("B200_GQ")) ctx->gq_ok = ctx->gq_ok && atoi(e) != 0;
    ctx->q_owned_valid = ctx->q_ghost_valid = false;
  }
  TRY(reserve(ctx, ctx->ostart, (size_t)g.mbins + 2));
  TRY(reserve(ctx, ctx->gstart, (size_t)g.mbins + 2));
  TRY(reserve(ctx, ctx->tilesum, (size_t)cdiv(g.mbins + 1, SCAN_TILE) + 2));
  ctx->geom_ready = true;
  return B200_OK;
}

// exclusive scan of counts[0..n) in place; element n receives the total
static int scan_inplace(b200_ctx *ctx, int *a, int n) {
  const int ntiles = cdiv(n, SCAN_TILE);
  int *gt = ctx->flags + 3;
  k_scan_tiles<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(a, a, n, ctx->tilesum.p);
  k_scan_sums<<<1, 1024, 0, ctx->stream>>>(ctx->tilesum.p, ntiles, gt);
  k_scan_add<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(a, n, ctx->tilesum.p, gt);
  ctx->launches += 3;
  LAUNCH_CHECK();
  return B200_OK;
}

static int check_err_flags(b200_ctx *ctx, int e) {
  if (e & 1) return ctx->fail(B200_ENONFINITE, "Non-numeric atom coords - simulation unstable");
  if (e & 2) return ctx->fail(B200_ELOST, "atom outside the local bin grid (lost atom)");
  if (e & 8) return ctx->fail(B200_ECUDA, "peer-memory halo timed out waiting for a neighbour");
  return B200_OK;
}

// ------------------------------------------------------------------ list build (bin tiles)
static const int TILE_MENU[][3] = {{8, 8, 4}, {8, 4, 4}, {4, 4, 4}, {4, 4, 2}, {4, 2, 2},
                                   {2, 2, 2}, {2, 2, 1}, {2, 1, 1}, {1, 1, 1}};
static const int TILE_NMENU = sizeof(TILE_MENU) / sizeof(TILE_MENU[0]);
static const size_t TILE_SMEM_MAX = 227 * 1024;      // opt-in limit per CTA on sm_100
static const size_t TILE_SMEM_TWO = 113 * 1024;      // two CTAs per SM fit below this
static const size_t BUILD2_SMEM_MAX = 226 * 1024;    // k_tile_build2: dynamic part next to its static arrays

static void set_tile_geom(b200_ctx *ctx, const int t[3]) {
  TileGeom &G = ctx->tg;
  G.ntiles = 1;
  for (int d = 0; d < 3; d++) {
    G.t[d] = std::max(1, std::min(t[d], ctx->ibin_n[d]));
    G.nt[d] = cdiv(ctx->ibin_n[d], G.t[d]);
    G.ilo[d] = ctx->ibin_lo[d];
    G.nib[d] = ctx->ibin_n[d];
    G.mbin[d] = ctx->geom.mbin[d];
    G.ntiles *= G.nt[d];
    G.bsize[d] = (ctx->geom.binhi[d] - ctx->geom.binlo[d]) / ctx->geom.nbin[d];
    G.bin0[d] = ctx->geom.binlo[d] + ctx->geom.mbinlo[d] * G.bsize[d];

  }
  {
    // fixed-point staging (k_tile_lj_fx): the staged bins plus a pad of one neighbour cutoff on
    // either side must fit 31 bits; power-of-two scale so that (x - origin) * scale is exact
    double ext = 0.0;
    for (int d = 0; d < 3; d++) ext = std::max(ext, (G.t[d] + 2 * G.s[d]) * G.bsize[d]);
    G.fxpad = ctx->cutneighmax;
    G.fxscale = std::ldexp(1.0, (int)std::floor(std::log2(2147483000.0 / (ext + 2.0 * G.fxpad))));
  }
  G.srow_y = G.t[1] + 2 * G.s[1];
  G.srow_z = G.t[2] + 2 * G.s[2];
  G.sbx = G.t[0] + 2 * G.s[0];
}

template <class K>
static int tile_attr(b200_ctx *ctx, K kernel) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_MAX));
  return B200_OK;
}

static int lj2_attrs(b200_ctx *ctx);
static int lj2f_attrs(b200_ctx *ctx);
static int tile_kernel_attrs(b200_ctx *ctx) {
  TRY(tile_attr(ctx, (k_tile_build<true, true>)));
  TRY(tile_attr(ctx, (k_tile_build<false, true>)));
  TRY(tile_attr(ctx, (k_tile_build<true, false>)));
  TRY(tile_attr(ctx, (k_tile_build<false, false>)));
  TRY(tile_attr(ctx, (k_tile_build<true, true, true>)));
  TRY(tile_attr(ctx, (k_tile_build<true, true, false, true, true>)));
  TRY(tile_attr(ctx, (k_tile_build<false, true, false, true, true>)));
  TRY(tile_attr(ctx, (k_tile_build<true, true, true, true, true>)));
  TRY(tile_attr(ctx, (k_tile_build<true, true, false, false, true>)));
  TRY(tile_attr(ctx, (k_tile_build<false, true, false, false, true>)));
  TRY(tile_attr(ctx, (k_tile_build<true, true, true, false, true>)));
  TRY(tile_attr(ctx, k_tile_export));
  TRY(tile_attr(ctx, k_peratom_tile<1>));
  TRY(tile_attr(ctx, k_peratom_tile<2>));
  // (k_tile_build2 also has 256 bytes of static shared memory: the opt-in limit covers both)
#define B2A(K) CK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BUILD2_SMEM_MAX))
  B2A((k_tile_build2<true, true, false>));
  B2A((k_tile_build2<false, true, false>));
  B2A((k_tile_build2<true, false, false>));
  B2A((k_tile_build2<false, false, false>));
  B2A((k_tile_build2<true, true, true>));
#undef B2A
  TRY(tile_attr(ctx, (k_tile_eam2_rho<false, 352, 2>)));
  TRY(tile_attr(ctx, (k_tile_eam2_rho<true, 352, 2>)));
  TRY(tile_attr(ctx, (k_tile_eam2_force<false, 352, 2, false>)));
  TRY(tile_attr(ctx, (k_tile_eam2_force<true, 352, 2, false>)));
  TRY(tile_attr(ctx, (k_tile_eam2_force<false, 352, 2, true>)));
#define A3(K) \
  TRY(tile_attr(ctx, K<false, false, false>)); TRY(tile_attr(ctx, K<false, false, true>)); \
  TRY(tile_attr(ctx, K<false, true, false>));  TRY(tile_attr(ctx, K<false, true, true>));  \
  TRY(tile_attr(ctx, K<true, false, false>));  TRY(tile_attr(ctx, K<true, false, true>));  \
  TRY(tile_attr(ctx, K<true, true, false>));   TRY(tile_attr(ctx, K<true, true, true>));
  A3(k_tile_lj)
#undef A3
  TRY(lj2_attrs(ctx));
  TRY(lj2f_attrs(ctx));
  TRY(tile_attr(ctx, k_tile_lj_fx<false, false>));
  TRY(tile_attr(ctx, k_tile_lj_fx<false, true>));
  TRY(tile_attr(ctx, k_tile_lj_fx<true, false>));
  TRY(tile_attr(ctx, k_tile_lj_fx<true, true>));
  TRY(tile_attr(ctx, k_tile_eam_rho<false>));
  TRY(tile_attr(ctx, k_tile_eam_rho<true>));
  TRY(tile_attr(ctx, k_tile_eam_force<false, false>));
  TRY(tile_attr(ctx, k_tile_eam_force<false, true>));
  TRY(tile_attr(ctx, k_tile_eam_force<true, false>));
  TRY(tile_attr(ctx, k_tile_eam_force<true, true>));
  return B200_OK;
}

// may eam run on the second-generation tile kernels?  (single element, one atom type, FP64)
// Auto: measured on one B200 (profiles/r02ab_probe_build2_eam_sizes.txt) the tile kernels win on
// small systems (32 k atoms: 249 vs 215 M atom-steps/s -- half the launches, no reverse halos) and
// lose on large ones (2 M atoms: pair 1.70 vs 1.33 ms, list build 2.0 vs 0.7 ms: the spline
// tables, not the scatter, are what loads the L1 data pipe, and every owned-owned pair looks them
// up twice).  The switch is the number of global bins per sub-domain, identical on every rank (all
// ranks must walk the same sequence of halos).
static bool eam2_usable(const b200_ctx *ctx) {
  if (!(ctx->eam2 && ctx->pair_style == 2 && ctx->eam_one_ok && ctx->ntypes == 1 &&
        ctx->prec == B200_PREC_DOUBLE))
    return false;
  if (ctx->eam2 == 1 || !ctx->newton) return true;  // newton off lives on the FULLGHOST tile rows
  const long long bins = (long long)ctx->geom.nbin[0] * ctx->geom.nbin[1] * ctx->geom.nbin[2];
  return bins / std::max(ctx->nranks, 1) <= ctx->eam2_max_bins;
}

// rc: B200_OK with ctx->tiles_active set, or tiles_active == false when no tile size fits
static int build_tiles(b200_ctx *ctx) {
  const int nl = ctx->nlocal, c = ctx->cur;
  cudaStream_t s = ctx->stream;
  ctx->tiles_active = false;
  const bool eam = ctx->pair_style == 2;
  const bool eam2 = eam && eam2_usable(ctx);
  ctx->eam2_active = false;
  // triclinic: the tile rows that hold every ghost partner (lj/cut, eam2) carry the tag rule; the
  // first-generation eam tile kernels (FWD ghosts + scatter) stay orthogonal-only -> flat list
  if (ctx->tri && eam && !eam2) return B200_OK;
  if (ctx->tile_level < 0) {
    // first build for this geometry: the largest tile that still gives every SM several CTAs
    int lvl = 0;
    for (; lvl < TILE_NMENU - 1; lvl++) {
      set_tile_geom(ctx, TILE_MENU[lvl]);
      if (ctx->tg.ntiles >= 2 * 148) break;
    }
    ctx->tile_level = lvl;
  }
  int h[8] = {0};
  for (;;) {
    if (ctx->tile_req[0] > 0) set_tile_geom(ctx, ctx->tile_req);
    else set_tile_geom(ctx, TILE_MENU[ctx->tile_level]);
    const TileGeom &G = ctx->tg;
    bool fits = G.srow_y * G.srow_z <= TILE_MAXROWS && G.t[1] * G.t[2] <= TILE_MAXRUNS;
    if (fits) {
      TRY(reserve(ctx, ctx->tile_ibase, (size_t)G.ntiles + 2));
      TRY(reserve(ctx, ctx->tilesum, (size_t)cdiv(std::max(G.ntiles, ctx->geom.mbins) + 1, SCAN_TILE) + 2));
      CK(cudaMemsetAsync(ctx->tflags, 0, 8 * sizeof(int), s));
      TRY(reserve(ctx, ctx->tile_bflag, (size_t)G.ntiles + 2));
      TRY(reserve(ctx, ctx->tile_bpos, (size_t)G.ntiles + 2));
      TRY(reserve(ctx, ctx->tile_ids, (size_t)G.ntiles + 2));
      TRY(reserve(ctx, ctx->tile_hdrs, (size_t)G.ntiles * TILE_HDR_BYTES));
      k_tile_count<<<G.ntiles, 128, 0, s>>>(G, ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p,
                                            ctx->tile_bflag.p, ctx->tflags, (eam && !eam2) ? 1 : 0, ctx->tile_hdrs.p);
      ctx->launches++;
      LAUNCH_CHECK();
      TRY(scan_inplace(ctx, ctx->tile_ibase.p, G.ntiles));
      CK(cudaMemcpyAsync(ctx->h_flags + 16, ctx->flags + 3, sizeof(int), cudaMemcpyDeviceToHost, s));
      // interior / boundary split of the tile ids (exclusive scan of the boundary flags)
      CK(cudaMemcpyAsync(ctx->tile_bpos.p, ctx->tile_bflag.p, sizeof(int) * G.ntiles, cudaMemcpyDeviceToDevice, s));
      TRY(scan_inplace(ctx, ctx->tile_bpos.p, G.ntiles));
      k_tile_split<<<cdiv(G.ntiles, 256), 256, 0, s>>>(G.ntiles, ctx->tile_bflag.p, ctx->tile_bpos.p,
                                                      ctx->tile_ids.p);
      ctx->launches++;
      LAUNCH_CHECK();
      CK(cudaMemcpyAsync(ctx->h_flags + 17, ctx->flags + 3, sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(ctx->h_flags + 8, ctx->tflags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      ctx->tile_nbnd = ctx->h_flags[17];
      ctx->tile_nint = G.ntiles - ctx->tile_nbnd;
      memcpy(h, ctx->h_flags + 8, sizeof h);
      if (h[4] != nl)
        return ctx->fail(B200_ELOST, "bin tiles cover %d of %d owned atoms", h[4], nl);
      ctx->tile_NI = ctx->h_flags[16];
      ctx->tile_scap = cdiv(std::max(h[0], 1) + 1, 64) * 64;  // + the dummy atom of padding entries
      const size_t bneed = ctx->build2 ? build2_smem_bytes(ctx->tile_scap, G.srow_y * G.srow_z, G.sbx, ctx->ntypes != 1,
                                                           std::max(ctx->tile_slots, 112))
                                       : tile_smem_bytes(ctx->tile_scap, G.srow_y * G.srow_z, G.sbx, true, false);
      const size_t need = std::max(bneed,
                                   eam2 ? eam2_smem_bytes(ctx->tile_scap, true)
                                        : tile_smem_bytes(ctx->tile_scap, G.srow_y * G.srow_z, G.sbx, false, eam));
      const bool last = ctx->tile_req[0] > 0 || ctx->tile_level == TILE_NMENU - 1;
      fits = h[0] <= TILE_MAXSTAGE && need <= (last ? TILE_SMEM_MAX : TILE_SMEM_TWO);
    }
    if (fits) break;
    if (ctx->tile_req[0] > 0 || ctx->tile_level == TILE_NMENU - 1) return B200_OK;  // flat list instead
    ctx->tile_level++;
  }
  const TileGeom &G = ctx->tg;
  const int rows = G.srow_y * G.srow_z;
  // one thread per owned atom of the fullest tile; very full tiles take two passes
  ctx->tile_maxown = h[1];
  int thr = cdiv(std::max(h[1], 1), 32) * 32;
  if (thr > 352) thr = cdiv(cdiv(h[1], cdiv(h[1], 352)), 32) * 32;
  ctx->tile_threads = std::min(std::max(thr, 64), 352);
  if (ctx->tile_slots == 0) ctx->tile_slots = 112;
  const int n1 = ctx->ntypes + 1;
  const size_t smem = tile_smem_bytes(ctx->tile_scap, rows, G.sbx, true, false);
  for (int attempt = 0; attempt < 4; attempt++) {
    ctx->tile_slots = cdiv(ctx->tile_slots, 8) * 8;
    const size_t NI = (size_t)std::max(ctx->tile_NI, 32);
    TRY(reserve(ctx, ctx->tl_list, NI * (ctx->tile_slots / 8)));
    TRY(reserve(ctx, ctx->tl_iloc, NI));
    TRY(reserve(ctx, ctx->tl_num, NI));
    TRY(reserve(ctx, ctx->tl_gi, NI));
    if (eam2) TRY(reserve(ctx, ctx->tl_far, NI));
    TRY(reserve(ctx, ctx->numneigh, (size_t)std::max(nl, 1)));
    CK(cudaMemsetAsync(ctx->tflags + 2, 0, 2 * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->tflags + 5, 0, sizeof(int), s));
    const int ph1 = ph_begin(ctx, B200_PH_BUILD);
    const bool one = ctx->ntypes == 1;
    const double cut1 = one ? ctx->cutneighsq_h[n1 + 1] : 0.0;
#define TB(ONE, FULL)                                                                                  \
  k_tile_build<ONE, FULL><<<G.ntiles, ctx->tile_threads, smem, s>>>(                                  \
      G, ctx->fst, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->atombin[c], ctx->tile_ibase.p,   \
      ctx->tile_NI, ctx->tile_slots, cut1, ctx->cutneighsq_d.p, ctx->ntypes, ctx->tl_iloc.p,           \
      ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->numneigh.p, ctx->tile_scap, ctx->tflags)
    // lj/cut: every ghost partner is stored (no scatter, no reverse halo); eam: FWD ghosts only
    const size_t smem2 = build2_smem_bytes(ctx->tile_scap, rows, G.sbx, !one, ctx->tile_slots);
    const double rs = eam2 ? std::max(0.0, std::sqrt(ctx->eam.cutforcesq) + ctx->eam2_margin * ctx->skin) : 0.0;
#define TB2(ONE, FULL, SPL)                                                                            \
  k_tile_build2<ONE, FULL, SPL><<<G.ntiles, B2_WARPS * 32, smem2, s>>>(                               \
      G, ctx->fst, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI,      \
      ctx->tile_slots, cut1, ctx->cutneighsq_d.p, ctx->ntypes, ctx->tl_iloc.p, ctx->tl_num.p,          \
      ctx->tl_gi.p, ctx->tl_list.p, ctx->numneigh.p, ctx->tile_scap, ctx->tflags, rs * rs,             \
      eam2 ? ctx->tl_far.p : nullptr)
    const double tdelta = 0.01 * ctx->angstrom;  // npair_bin.cpp:59
#define TBT(ONE, SPL, TRI)                                                                             \
  k_tile_build<ONE, true, SPL, TRI, true><<<G.ntiles, ctx->tile_threads, smem, s>>>(                  \
      G, ctx->fst, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->atombin[c], ctx->tile_ibase.p,   \
      ctx->tile_NI, ctx->tile_slots, cut1, ctx->cutneighsq_d.p, ctx->ntypes, ctx->tl_iloc.p,           \
      ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->numneigh.p, ctx->tile_scap, ctx->tflags,       \
      SPL ? rs * rs : 0.0, SPL ? ctx->tl_far.p : nullptr, ctx->tag[c], tdelta, ctx->newton ? 0 : 1, ctx->exg,  \
      ctx->mask[c])
    if (!ctx->newton && eam && !eam2) return B200_OK;  // no FULLGHOST rows: build_list reports it
    // triclinic boxes, newton off and group exclusions: the instantiations that carry those rules
    // (always on rows holding every ghost partner: lj/cut, eam2)
    const bool extra = !ctx->newton || ctx->exg.n > 0;
    if ((ctx->tri || extra) && eam && !eam2) return B200_OK;  // flat list instead (build_list)
    if (ctx->tri) {
      if (eam2) TBT(true, true, true);
      else if (one) TBT(true, false, true);
      else TBT(false, false, true);
    } else if (extra) {
      if (eam2) TBT(true, true, false);
      else if (one) TBT(true, false, false);
      else TBT(false, false, false);
    } else if (ctx->build2 && smem2 <= BUILD2_SMEM_MAX) {
      if (eam2) TB2(true, true, true);
      else if (!eam) { if (one) TB2(true, true, false); else TB2(false, true, false); }
      else           { if (one) TB2(true, false, false); else TB2(false, false, false); }
    } else if (eam2) {
      // NEAR = partners stored within the force cutoff + a margin of the skin (kernels_eam2.cuh)
      k_tile_build<true, true, true><<<G.ntiles, ctx->tile_threads, smem, s>>>(
          G, ctx->fst, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->atombin[c], ctx->tile_ibase.p,
          ctx->tile_NI, ctx->tile_slots, cut1, ctx->cutneighsq_d.p, ctx->ntypes, ctx->tl_iloc.p,
          ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->numneigh.p, ctx->tile_scap, ctx->tflags,
          rs * rs, ctx->tl_far.p);
    } else if (!eam) { if (one) TB(true, true); else TB(false, true); }
    else      { if (one) TB(true, false); else TB(false, false); }
#undef TB
#undef TB2
#undef TBT
    ctx->launches++;
    LAUNCH_CHECK();
    ph_end(ctx, ph1);
    CK(cudaMemcpyAsync(ctx->h_flags + 8, ctx->tflags, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    TRY(check_err_flags(ctx, ctx->h_flags[1]));
    if (ctx->h_flags[8 + 5])
      return ctx->fail(B200_ECAPACITY, "bin tile stages %d atoms, capacity %d", ctx->h_flags[8 + 5], ctx->tile_scap);
    ctx->tile_maxfull = ctx->h_flags[8 + 2];
    ctx->max_numneigh = ctx->h_flags[8 + 3];
    if (ctx->max_numneigh > ctx->one)
      return ctx->fail(B200_ECAPACITY, "Neighbor list overflow, boost neigh_modify one (%d > %d)",
                       ctx->max_numneigh, ctx->one);
    if (ctx->tile_maxfull <= ctx->tile_slots) {
      ctx->tiles_active = true;
      ctx->full_ghost = !eam || eam2;
      ctx->eam2_active = eam2;
      ctx->maxneigh = ctx->tile_slots;
      return B200_OK;
    }
    ctx->tile_slots = (ctx->tile_maxfull * 9 / 8 + 8) / 8 * 8;
  }
  return ctx->fail(B200_ECAPACITY, "neighbor list did not converge");
}

// ------------------------------------------------------------------ list build
static int build_list(b200_ctx *ctx) {
  const int nl = ctx->nlocal, c = ctx->cur;
  // eam evaluates an expensive pair function: computing every owned-owned pair from both sides
  // costs more there than the Newton scatter it removes (profiles/r01g_probe_eam_*), so the
  // flat half list stays the default for it
  const bool want_tiles = ctx->list_mode == 1 ||
                          (ctx->list_mode == 0 && (ctx->pair_style == 1 || (ctx->pair_style == 2 && eam2_usable(ctx))));
  if (ctx->use_tiles && (want_tiles || (!ctx->newton && ctx->list_mode != 2)) && nl > 0) {
    TRY(build_tiles(ctx));
    if (ctx->tiles_active && (ctx->newton || ctx->full_ghost)) return B200_OK;
  }
  // newton off (force.cpp newton_pair = 0): every boundary pair is evaluated by both owners and
  // nothing is sent back -- which is what the tile rows holding every ghost partner do anyway
  // (k_tile_build FULLGHOST).  The flat half list scatters onto ghosts and returns their forces:
  // a Newton-on scheme with no newton-off variant.
  if (!ctx->newton && nl > 0)
    return ctx->fail(B200_EARG, "newton off needs the bin-tile list: lj/cut, or single-element eam in double "
                                "precision (package b200 list flat / multi-element eam are newton-on only)");
  ctx->tiles_active = false;
  ctx->full_ghost = false;
  ctx->eam2_active = false;
  if (ctx->maxneigh == 0) ctx->maxneigh = 96;
  for (int attempt = 0; attempt < 4; attempt++) {
    ctx->nstride = cdiv(std::max(nl, 1), 32) * 32;
    ctx->maxneigh = cdiv(ctx->maxneigh, 8) * 8;  // whole slot groups for every tpa
    TRY(reserve(ctx, ctx->neigh, (size_t)ctx->maxneigh * ctx->nstride));
    TRY(reserve(ctx, ctx->numneigh, (size_t)nl));
    CK(cudaMemsetAsync(ctx->flags + 2, 0, sizeof(int), ctx->stream));
    const int ph1 = ph_begin(ctx, B200_PH_BUILD);
    if (nl > 0) {
      const int n1 = ctx->ntypes + 1;
      if (ctx->tri) {
        const double delta = 0.01 * ctx->angstrom;  // npair_bin.cpp:59
        if (ctx->ntypes == 1)
          k_build_half_tri<true><<<cdiv(nl, 128), 128, 0, ctx->stream>>>(
              nl, ctx->nstride, ctx->maxneigh, ctx->tpa, ctx->xt[c], ctx->tag[c], ctx->atombin[c], ctx->ostart.p,
              ctx->gstart.p, ctx->stencil, ctx->cutneighsq_h[n1 + 1], ctx->cutneighsq_d.p, ctx->ntypes, delta,
              ctx->numneigh.p, ctx->neigh.p, ctx->flags + 2, ctx->exg, ctx->mask[c]);
        else
          k_build_half_tri<false><<<cdiv(nl, 128), 128, 0, ctx->stream>>>(
              nl, ctx->nstride, ctx->maxneigh, ctx->tpa, ctx->xt[c], ctx->tag[c], ctx->atombin[c], ctx->ostart.p,
              ctx->gstart.p, ctx->stencil, 0.0, ctx->cutneighsq_d.p, ctx->ntypes, delta, ctx->numneigh.p,
              ctx->neigh.p, ctx->flags + 2, ctx->exg, ctx->mask[c]);
      } else {
#define BH(ONE, EXG)                                                                                   \
  k_build_half<ONE, EXG><<<cdiv(nl, 128), 128, 0, ctx->stream>>>(                                      \
      nl, ctx->nstride, ctx->maxneigh, ctx->tpa, ctx->xt[c], ctx->atombin[c], ctx->ostart.p, ctx->gstart.p, \
      ctx->stencil, ONE ? ctx->cutneighsq_h[n1 + 1] : 0.0, ctx->cutneighsq_d.p, ctx->ntypes, ctx->numneigh.p, \
      ctx->neigh.p, ctx->flags + 2, ctx->exg, ctx->mask[c])
        const bool one = ctx->ntypes == 1;
        if (ctx->exg.n) { if (one) BH(true, true); else BH(false, true); }
        else            { if (one) BH(true, false); else BH(false, false); }
#undef BH
      }
      ctx->launches++;
      LAUNCH_CHECK();
    }
    ph_end(ctx, ph1);
    CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    TRY(check_err_flags(ctx, ctx->h_flags[1]));
    ctx->max_numneigh = ctx->h_flags[2];
    if (ctx->max_numneigh <= ctx->maxneigh) return B200_OK;
    if (ctx->max_numneigh > ctx->one)
      return ctx->fail(B200_ECAPACITY, "Neighbor list overflow, boost neigh_modify one (%d > %d)",
                       ctx->max_numneigh, ctx->one);
    ctx->maxneigh = std::min(ctx->one, (ctx->max_numneigh * 9 / 8 + 8) / 8 * 8);
  }
  return ctx->fail(B200_ECAPACITY, "neighbor list did not converge");
}

// The captured timestep bakes in atom counts, buffer addresses and every kernel argument:
// anything that can change one of them drops it (it is re-captured on the next plain step).
static void drop_step_graph(b200_ctx *ctx) {
  if (ctx->step_graph) {
    cudaGraphExecDestroy(ctx->step_graph);
    ctx->step_graph = nullptr;
  }
}

// ------------------------------------------------------------------ deferred integrator operations
// b200_scale_v / _v3, b200_nve_v, b200_nve_x queue their per-atom loop (VOps, kernels_step.cuh);
// whoever next reads or writes x, v or f -- the rebuild vote, halo, rebuild, pair stage, a
// temperature sum, a download, another force term -- first runs the queue in one pass.  With
// ke_out the pass also leaves ComputeTemp's sums of the updated velocities there (7 doubles).
static int flush_vops(b200_ctx *ctx, int ke_groupbit = 0, double *ke_out = nullptr) {
  if (ctx->vops.n == 0 && !ke_out) return B200_OK;
  CK(cudaSetDevice(ctx->device));
  const int nl = ctx->nlocal, c = ctx->cur;
  if (nl > 0) {
    const int grid = ke_out ? std::min(cdiv(nl, 256), 148 * 8) : cdiv(nl, 256);
    if (ke_out)
      k_vops<true><<<grid, 256, 0, ctx->stream>>>(nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->f[0],
                                                  ctx->f[1], ctx->f[2], ctx->mask[c], ctx->mass_d.p, ctx->vops,
                                                  ctx->vops_chk, ctx->xh[0], ctx->xh[1], ctx->xh[2], ctx->triggersq,
                                                  ctx->flags, ke_groupbit, ke_out);
    else
      k_vops<false><<<grid, 256, 0, ctx->stream>>>(nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->f[0],
                                                   ctx->f[1], ctx->f[2], ctx->mask[c], ctx->mass_d.p, ctx->vops,
                                                   ctx->vops_chk, ctx->xh[0], ctx->xh[1], ctx->xh[2], ctx->triggersq,
                                                   ctx->flags, 0, nullptr);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  ctx->vops.n = 0;
  ctx->vops_chk = 0;
  return B200_OK;
}

// ------------------------------------------------------------------ reneighbor
// Verlet::run rebuild branch (verlet.cpp:268-297): pbc, exchange, borders, neighbor->build.
static int reneighbor(b200_ctx *ctx) {
  TRY(flush_vops(ctx));
  drop_step_graph(ctx);
  if (!ctx->geom_ready) TRY(setup_geometry(ctx));
  if (ctx->box_changes) {  // Neighbor::build: boxlo_hold / boxhi_hold (neighbor.cpp:2533-2541)
    for (int d = 0; d < 3; d++) {
      ctx->boxlo_hold[d] = ctx->boxlo[d];
      ctx->boxhi_hold[d] = ctx->boxhi[d];
    }
    ctx->deltasq = ctx->triggersq;
  }
  const int ph2 = ph_begin(ctx, B200_PH_NEIGH);
  const Geom &g = ctx->geom;
  const bool multi = ctx->nranks > 1;
  const int nl0 = ctx->nlocal;
  int c = ctx->cur;
  cudaStream_t s = ctx->stream;
  ctx->nghost = 0;  // the ghost region holds nothing worth keeping from here on
  CK(cudaMemsetAsync(ctx->ostart.p, 0, sizeof(int) * (g.mbins + 1), s));
  CK(cudaMemsetAsync(ctx->gstart.p, 0, sizeof(int) * (g.mbins + 1), s));
  CK(cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), s));
  CK(cudaMemsetAsync(ctx->counts, 0, 32 * sizeof(int), s));
  int *err = ctx->counts + 27;
  // ---- Domain::pbc + coord2bin (+ CommBrick::exchange classification)
  if (nl0 > 0) {
    if (multi)
      k_pbc_bin<true><<<cdiv(nl0, 256), 256, 0, s>>>(nl0, ctx->xt[c], ctx->image[c], g, ctx->owner,
                                                     ctx->atombin[c], ctx->slot, ctx->ostart.p,
                                                     ctx->counts, err);
    else
      k_pbc_bin<false><<<cdiv(nl0, 256), 256, 0, s>>>(nl0, ctx->xt[c], ctx->image[c], g, ctx->owner,
                                                      ctx->atombin[c], ctx->slot, ctx->ostart.p,
                                                      ctx->counts, err);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  int ntot = nl0, nl = nl0;
  if (multi) {
    // ---- CommBrick::exchange: ship leavers to the sub-domain that owns them now
    TRY(sync_counts(ctx));
    TRY(check_err_flags(ctx, global_err(ctx)));
    int moff[NDIR + 1], aoff[NDIR + 1];
    moff[0] = aoff[0] = 0;
    for (int dir = 0; dir < NDIR; dir++) {
      const bool rem = (ctx->remote_mask >> dir) & 1u;
      const int from = ctx->nbr[NDIR - 1 - dir];
      moff[dir + 1] = moff[dir] + (rem ? ctx->h_counts[ctx->rank * 32 + dir] : 0);
      aoff[dir + 1] = aoff[dir] + ((rem && from >= 0) ? ctx->h_counts[from * 32 + dir] : 0);
      if (!rem && ctx->h_counts[ctx->rank * 32 + dir] && dir != 13)
        return ctx->fail(B200_ELOST, "atom left the sub-domain towards a direction without neighbour");
    }
    const int nleave = moff[NDIR], narrive = aoff[NDIR];
    if (nl0 + narrive > ctx->nmax) {
      ctx->nlocal = nl0;
      TRY(alloc_atoms(ctx, (int)((nl0 + narrive) * 1.2) + 1024));
    }
    // (sized with a floor at the first rebuild: the first rebuild that really migrates atoms
    // must not be the one that allocates)
    const size_t mig_floor = std::max<size_t>(8192, (size_t)nl0 / 32);
    TRY(reserve(ctx, ctx->mig_send, std::max((size_t)nleave, mig_floor) * MIG_W));
    TRY(reserve(ctx, ctx->mig_recv, std::max((size_t)narrive, mig_floor) * MIG_W));
    CK(cudaMemcpyAsync(ctx->diroffset, moff, (NDIR + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (nleave > 0) {
      k_pack_migrate<<<cdiv(nl0, 256), 256, 0, s>>>(nl0, ctx->atombin[c], ctx->slot, ctx->diroffset,
                                                    ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2],
                                                    ctx->tag[c], ctx->mask[c], ctx->image[c],
                                                    ctx->mig_send.p);
      ctx->launches++;
    }
    TRY(halo_exchange(ctx, ctx->mig_send.p, moff, ctx->mig_recv.p, aoff, MIG_W, false));
    if (narrive > 0) {
      k_unpack_migrate<<<cdiv(narrive, 256), 256, 0, s>>>(
          narrive, nl0, ctx->mig_recv.p, g, ctx->owner, ctx->xt[c], ctx->v[c][0], ctx->v[c][1],
          ctx->v[c][2], ctx->tag[c], ctx->mask[c], ctx->image[c], ctx->atombin[c], ctx->slot,
          ctx->ostart.p, err);
      ctx->launches++;
    }
    LAUNCH_CHECK();
    ntot = nl0 + narrive;
    nl = nl0 - nleave + narrive;
    CK(cudaMemsetAsync(ctx->counts, 0, 27 * sizeof(int), s));
  }
  // ---- counting sort of the owned atoms by bin (+ xhold)
  TRY(scan_inplace(ctx, ctx->ostart.p, g.mbins));
  if (ntot > 0) {
    const int *okey = nullptr;
    if (ctx->stable_order) {
      TRY(reserve(ctx, ctx->okey, (size_t)ntot));
      k_bin_keys<<<cdiv(ntot, 256), 256, 0, s>>>(ntot, ctx->atombin[c], ctx->slot, ctx->ostart.p, ctx->tag[c],
                                                ctx->okey.p);
      ctx->launches++;
      okey = ctx->okey.p;
    }
    k_permute_owned<<<cdiv(ntot, 256), 256, 0, s>>>(
        ntot, ctx->atombin[c], ctx->slot, okey, ctx->ostart.p, ctx->xt[c], ctx->xt[c ^ 1], ctx->v[c][0],
        ctx->v[c][1], ctx->v[c][2], ctx->v[c ^ 1][0], ctx->v[c ^ 1][1], ctx->v[c ^ 1][2], ctx->tag[c],
        ctx->tag[c ^ 1], ctx->mask[c], ctx->mask[c ^ 1], ctx->image[c], ctx->image[c ^ 1],
        ctx->atombin[c ^ 1], ctx->xh[0], ctx->xh[1], ctx->xh[2], g);
    ctx->launches++;
  }
  c ^= 1;
  ctx->cur = c;
  ctx->nlocal = nl;
  // ---- CommBrick::borders: count per direction, exchange the counts, fill, ship, place
  if (nl > 0) {
    k_border<0><<<cdiv(nl, 256), 256, 0, s>>>(nl, ctx->xt[c], g, ctx->counts, ctx->diroffset, nullptr,
                                              nullptr);
    ctx->launches++;
  }
  LAUNCH_CHECK();
  TRY(sync_counts(ctx));
  TRY(check_err_flags(ctx, global_err(ctx)));
  ctx->sendoff[0] = ctx->recvoff[0] = 0;
  for (int dir = 0; dir < NDIR; dir++) {
    const bool rem = (ctx->remote_mask >> dir) & 1u;
    const int from = ctx->nbr[NDIR - 1 - dir];
    const int mine = ctx->h_counts[ctx->rank * 32 + dir];
    const int theirs = rem ? (from >= 0 ? ctx->h_counts[from * 32 + dir] : 0) : mine;
    ctx->sendoff[dir + 1] = ctx->sendoff[dir] + mine;
    ctx->recvoff[dir + 1] = ctx->recvoff[dir] + theirs;
  }
  const int nsend = ctx->sendoff[NDIR], ng = ctx->recvoff[NDIR];
  ctx->nsend = nsend;
  if (nl + ng > ctx->nmax) {
    TRY(alloc_atoms(ctx, (int)((nl + ng) * 1.1) + 1024));
    c = ctx->cur;
  }
  ctx->nghost = ng;
  TRY(reserve(ctx, ctx->sendlist, (size_t)nsend));
  TRY(reserve(ctx, ctx->senddir, (size_t)nsend));
  TRY(reserve(ctx, ctx->gsrc, (size_t)ng));
  TRY(reserve(ctx, ctx->gbin, (size_t)ng));
  TRY(reserve(ctx, ctx->gslot, (size_t)ng));
  TRY(reserve(ctx, ctx->gtag_tmp, (size_t)ng));
  TRY(reserve(ctx, ctx->gsrc_tmp, (size_t)ng));
  TRY(reserve(ctx, ctx->gdir, (size_t)ng));
  TRY(reserve(ctx, ctx->gdir_tmp, (size_t)ng));
  TRY(reserve(ctx, ctx->gtmp, (size_t)ng));
  if (multi) {
    TRY(ensure_arena(ctx, (size_t)nsend * 4, (size_t)ng * 4));
    TRY(p2p_publish_borders(ctx));
  }
  CK(cudaMemcpyAsync(ctx->diroffset, ctx->sendoff, (NDIR + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->recvoffset, ctx->recvoff, (NDIR + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(ctx->counts, 0, 27 * sizeof(int), s));
  if (nsend > 0) {
    k_border<1><<<cdiv(nl, 256), 256, 0, s>>>(nl, ctx->xt[c], g, ctx->counts, ctx->diroffset,
                                              ctx->sendlist.p, ctx->senddir.p);
    ctx->launches++;
    if (multi) {
      k_pack_border<<<cdiv(nsend, 256), 256, 0, s>>>(nsend, ctx->sendlist.p, ctx->senddir.p,
                                                     ctx->remote_mask, g, ctx->xt[c], ctx->tag[c],
                                                     reinterpret_cast<double4 *>(ctx->sbuf.p));
      ctx->launches++;
    }
  }
  LAUNCH_CHECK();
  TRY(halo_exchange(ctx, ctx->sbuf.p, ctx->sendoff, ctx->rbuf.p, ctx->recvoff, 4, false));
  if (ng > 0) {
    k_ghost_make<<<cdiv(ng, 256), 256, 0, s>>>(
        ng, ctx->recvoffset, ctx->diroffset, ctx->remote_mask, ctx->sendlist.p, g, ctx->xt[c],
        ctx->tag[c], reinterpret_cast<const double4 *>(ctx->rbuf.p), ctx->gtmp.p, ctx->gtag_tmp.p,
        ctx->gsrc_tmp.p, ctx->gbin.p, ctx->gslot.p, ctx->gdir_tmp.p, ctx->gstart.p, ctx->flags + 1);
    ctx->launches++;
  }
  // neigh_modify exclude group reads the group mask of every list candidate: ghosts owned by
  // another sub-domain get theirs through one more exchange in the order of the border records
  const double *rmask = nullptr;
  if (multi && ctx->exg.n) {
    TRY(reserve(ctx, ctx->mask_s, (size_t)std::max(nsend, 1)));
    TRY(reserve(ctx, ctx->mask_r, (size_t)std::max(ng, 1)));
    if (nsend > 0) {
      k_pack_mask<<<cdiv(nsend, 256), 256, 0, s>>>(nsend, ctx->sendlist.p, ctx->senddir.p, ctx->remote_mask,
                                                   ctx->mask[c], ctx->mask_s.p);
      ctx->launches++;
    }
    LAUNCH_CHECK();
    TRY(halo_exchange(ctx, ctx->mask_s.p, ctx->sendoff, ctx->mask_r.p, ctx->recvoff, 1, false));
    rmask = ctx->mask_r.p;
  }
  TRY(scan_inplace(ctx, ctx->gstart.p, g.mbins));
  if (ng > 0) {
    const long long *gkey = nullptr;
    if (ctx->stable_order) {
      TRY(reserve(ctx, ctx->gkey, (size_t)ng));
      k_ghost_keys<<<cdiv(ng, 256), 256, 0, s>>>(ng, ctx->gbin.p, ctx->gslot.p, ctx->gstart.p, ctx->gtag_tmp.p,
                                                 ctx->gdir_tmp.p, ctx->gkey.p);
      ctx->launches++;
      gkey = ctx->gkey.p;
    }
    k_ghost_place<<<cdiv(ng, 256), 256, 0, s>>>(ng, nl, ctx->gtmp.p, ctx->gtag_tmp.p, ctx->gsrc_tmp.p,
                                                ctx->gbin.p, ctx->gslot.p, ctx->gdir_tmp.p,
                                                ctx->gstart.p, ctx->xt[c], ctx->tag[c], ctx->mask[c],
                                                ctx->gsrc.p, ctx->gdir.p, ctx->xt[c ^ 1], gkey, rmask);
    ctx->launches++;
  }
  if (g.tri && nl > 0) {  // owned atoms back to box coordinates (the ghosts were converted as made)
    k_lamda2x<<<cdiv(nl, 256), 256, 0, s>>>(nl, ctx->xt[c], g);
    ctx->launches++;
  }
  LAUNCH_CHECK();
  TRY(build_list(ctx));
  // a rebuild invalidates ghost forces of the old ghost set
  for (int d = 0; d < 3; d++) CK(cudaMemsetAsync(ctx->f[d] + nl, 0, sizeof(double) * ng, s));
  ctx->ghost_f_clean = true;
  ctx->q_owned_valid = ctx->q_ghost_valid = false;
  ctx->ago = 0;
  ctx->nbuilds++;
  ph_end(ctx, ph2);
  return B200_OK;
}

// ------------------------------------------------------------------ step pieces
// Sub-domains that SHARE a device: the peer-memory halo kernels spin on flags another context's
// kernel raises, and streams of one device share a few hardware queues -- a spinning unpack
// kernel queued ahead of the pack kernel it waits for would never see it run.  So the contexts
// meet on the host after the pack launches and again after the unpack launches: every queue
// then holds this halo's pack kernels before its unpack kernels, and the previous halo's unpack
// kernels before this one's pack kernels.  (One device per sub-domain needs none of this.)
static inline void shared_device_rendezvous(b200_ctx *ctx) {
  if (ctx->grp && ctx->grp->shared_dev) ctx->grp->barrier();
}

// Grid of a peer-memory halo kernel.  Every block of these kernels spins on flags another
// sub-domain's kernel raises.  One device per sub-domain: one block per 256 records.  Sub-domains
// SHARING a device: the spinning blocks of all of them (pack + unpack) must be resident together
// with the kernels they wait for, or the device deadlocks until the spin limit (seen with 8
// sub-domains of 500 k atoms each); the kernels are grid-stride, so the grid is capped to a share
// of the 148 x 8 resident 256-thread blocks.
static inline int p2p_grid(const b200_ctx *ctx, int n) {
  int grid = cdiv(std::max(n, 1), 256);
  if (ctx->grp && ctx->grp->shared_dev) grid = std::min(grid, std::max(1, 148 * 8 / (4 * std::max(ctx->grp->n, 1))));
  return grid;
}

static int force_clear(b200_ctx *ctx) {
  TRY(flush_vops(ctx));  // a queued half-kick reads the forces this is about to clear
  const int ph3 = ph_begin(ctx, B200_PH_CLEAR);
  const int nall = ctx->nlocal + ctx->nghost;
  // mixed mode: the pair kernel stores f_i and k_merge_ff writes the ghosts; only the float4
  // scatter array is cleared (inside pair_compute)
  if (ctx->tiles_active && ctx->full_ghost) {
    // nothing to clear: owned forces are stored, ghost forces do not exist
  } else if (ctx->tiles_active) {
    // tile kernels store f of every owned atom; only the ghosts (Newton scatter targets) are
    // cleared -- unless the forward halo or the rebuild of this step already did
    if (ctx->nghost > 0 && !ctx->ghost_f_clean)
      for (int d = 0; d < 3; d++)
        CK(cudaMemsetAsync(ctx->f[d] + ctx->nlocal, 0, sizeof(double) * ctx->nghost, ctx->stream));
  } else if (ctx->prec != B200_PREC_MIXED)
    for (int d = 0; d < 3; d++) CK(cudaMemsetAsync(ctx->f[d], 0, sizeof(double) * nall, ctx->stream));
  ph_end(ctx, ph3);
  return B200_OK;
}

static int forward_comm(b200_ctx *ctx) {
  TRY(flush_vops(ctx));
  ctx->q_ghost_valid = false;  // ghost positions change: their fixed-point records follow in q_refresh
  const int ph4 = ph_begin(ctx, B200_PH_FORWARD);
  const int c = ctx->cur;
  cudaStream_t s = ctx->stream;
  // tile path: the unpack kernel also zeroes the ghost forces (force_clear then has nothing to do)
  Vec3Ptr fclear{{nullptr, nullptr, nullptr}};
  if (ctx->tiles_active && !ctx->full_ghost && ctx->nghost > 0) {
    fclear = Vec3Ptr{{ctx->f[0], ctx->f[1], ctx->f[2]}};
    ctx->ghost_f_clean = true;
  }
  if (ctx->p2p && ctx->remote_mask) {
    // pack + transfer in one kernel: records are stored straight into the neighbours' rbuf
    const long long seq = ++ctx->seqF;
    k_p2p_pack_forward<0><<<p2p_grid(ctx, ctx->nsend), 256, 0, s>>>(
        ctx->nsend, ctx->sendlist.p, ctx->senddir.p, ctx->diroffset, ctx->geom, ctx->xt[c], nullptr,
        ctx->fwdP, seq, ctx->p2p_counter + 0, (int *)ctx->p2p_counter + 4);
    shared_device_rendezvous(ctx);
    k_p2p_unpack_forward<0><<<p2p_grid(ctx, ctx->nghost), 256, 0, s>>>(
        ctx->nghost, ctx->nlocal, ctx->gsrc.p, ctx->gdir.p, ctx->geom, ctx->rbuf.p, ctx->xt[c], nullptr,
        ctx->fwdP, seq, ctx->p2p_counter + 1, (int *)ctx->p2p_counter + 4, fclear);
    shared_device_rendezvous(ctx);
    ctx->launches += 2;
  } else {
    if (ctx->remote_mask && ctx->nsend > 0) {
      k_pack_forward<<<cdiv(ctx->nsend, 256), 256, 0, s>>>(ctx->nsend, ctx->sendlist.p, ctx->senddir.p,
                                                         ctx->remote_mask, ctx->geom, ctx->xt[c],
                                                         ctx->sbuf.p);
      ctx->launches++;
    }
    TRY(halo_exchange(ctx, ctx->sbuf.p, ctx->sendoff, ctx->rbuf.p, ctx->recvoff, 3, false));
    if (ctx->nghost > 0) {
      k_unpack_forward<<<cdiv(ctx->nghost, 256), 256, 0, s>>>(ctx->nghost, ctx->nlocal, ctx->gsrc.p,
                                                             ctx->gdir.p, ctx->geom, ctx->rbuf.p,
                                                             ctx->xt[c], fclear);
      ctx->launches++;
    }
  }
  LAUNCH_CHECK();
  ph_end(ctx, ph4);
  return B200_OK;
}

// reverse halo of W per-atom doubles held in SoA arrays a[0..W)
template <int W>
static int reverse_halo(b200_ctx *ctx, Vec3Ptr a) {
  cudaStream_t s = ctx->stream;
  if (ctx->p2p && ctx->remote_mask) {
    const long long seq = ++ctx->seqR;
    k_p2p_pack_reverse<W><<<p2p_grid(ctx, ctx->nghost), 256, 0, s>>>(
        ctx->nghost, ctx->nlocal, ctx->gsrc.p, ctx->gdir.p, ctx->recvoffset, a, ctx->revP, seq,
        ctx->p2p_counter + 2, (int *)ctx->p2p_counter + 4);
    shared_device_rendezvous(ctx);
    k_p2p_unpack_reverse<W><<<p2p_grid(ctx, ctx->nsend), 256, 0, s>>>(
        ctx->nsend, ctx->sendlist.p, ctx->senddir.p, ctx->sbuf.p, a, ctx->revP, seq,
        ctx->p2p_counter + 3, (int *)ctx->p2p_counter + 4);
    shared_device_rendezvous(ctx);
    ctx->launches += 2;
    LAUNCH_CHECK();
    return B200_OK;
  }
  if (ctx->nghost > 0) {
    k_pack_reverse<W><<<cdiv(ctx->nghost, 256), 256, 0, s>>>(ctx->nghost, ctx->nlocal, ctx->gsrc.p, a,
                                                           ctx->rbuf.p);
    ctx->launches++;
  }
  TRY(halo_exchange(ctx, ctx->rbuf.p, ctx->recvoff, ctx->sbuf.p, ctx->sendoff, W, true));
  if (ctx->remote_mask && ctx->nsend > 0) {
    k_unpack_reverse<W><<<cdiv(ctx->nsend, 256), 256, 0, s>>>(ctx->nsend, ctx->sendlist.p,
                                                            ctx->senddir.p, ctx->remote_mask,
                                                            ctx->sbuf.p, a);
    ctx->launches++;
  }
  LAUNCH_CHECK();
  return B200_OK;
}

static int reverse_comm(b200_ctx *ctx) {
  TRY(flush_vops(ctx));
  // lj/cut on tiles evaluates boundary pairs on both sides: no ghost forces to return
  if (ctx->tiles_active && ctx->full_ghost) return B200_OK;
  const int ph5 = ph_begin(ctx, B200_PH_REVERSE);
  Vec3Ptr f{{ctx->f[0], ctx->f[1], ctx->f[2]}};
  TRY(reverse_halo<3>(ctx, f));
  ph_end(ctx, ph5);
  return B200_OK;
}

// EAM: PairEAM::pack/unpack_forward_comm of fp (pair_eam.cpp:1600-1621)
static int forward_scalar(b200_ctx *ctx, double *a) {
  cudaStream_t s = ctx->stream;
  if (ctx->p2p && ctx->remote_mask) {
    const long long seq = ++ctx->seqF;
    k_p2p_pack_forward<1><<<p2p_grid(ctx, ctx->nsend), 256, 0, s>>>(
        ctx->nsend, ctx->sendlist.p, ctx->senddir.p, ctx->diroffset, ctx->geom, nullptr, a, ctx->fwdP,
        seq, ctx->p2p_counter + 0, (int *)ctx->p2p_counter + 4);
    shared_device_rendezvous(ctx);
    k_p2p_unpack_forward<1><<<p2p_grid(ctx, ctx->nghost), 256, 0, s>>>(
        ctx->nghost, ctx->nlocal, ctx->gsrc.p, ctx->gdir.p, ctx->geom, ctx->rbuf.p, nullptr, a,
        ctx->fwdP, seq, ctx->p2p_counter + 1, (int *)ctx->p2p_counter + 4, Vec3Ptr{{nullptr, nullptr, nullptr}});
    shared_device_rendezvous(ctx);
    ctx->launches += 2;
    LAUNCH_CHECK();
    return B200_OK;
  }
  if (ctx->remote_mask && ctx->nsend > 0) {
    k_pack_forward_scalar<<<cdiv(ctx->nsend, 256), 256, 0, s>>>(ctx->nsend, ctx->sendlist.p,
                                                              ctx->senddir.p, ctx->remote_mask, a,
                                                              ctx->sbuf.p);
    ctx->launches++;
  }
  TRY(halo_exchange(ctx, ctx->sbuf.p, ctx->sendoff, ctx->rbuf.p, ctx->recvoff, 1, false));
  if (ctx->nghost > 0) {
    k_unpack_forward_scalar<<<cdiv(ctx->nghost, 256), 256, 0, s>>>(ctx->nghost, ctx->nlocal,
                                                                  ctx->gsrc.p, ctx->rbuf.p, a);
    ctx->launches++;
  }
  LAUNCH_CHECK();
  return B200_OK;
}

// FP64 lj/cut, second-generation kernel (kernels_tile2.cuh).  The launch shapes instantiated:
// {threads per CTA, CTAs per SM} in {352x2, 448x2, 320x2, 256x2, 320x3, 256x4}, 2 or 4 pair bodies in flight.
#define LJ2_SHAPES(X) X(352, 2) X(448, 2) X(320, 2) X(256, 2) X(320, 3) X(256, 4)
static int lj2_attrs(b200_ctx *ctx) {
#define X(T, B)                                                    \
  TRY(tile_attr(ctx, k_tile_lj2<false, true, 2, T, B>));           \
  TRY(tile_attr(ctx, k_tile_lj2<false, true, 4, T, B>));           \
  TRY(tile_attr(ctx, k_tile_lj2<true, true, 2, T, B>));            \
  TRY(tile_attr(ctx, k_tile_lj2<false, false, 2, T, B>));          \
  TRY(tile_attr(ctx, (k_tile_lj2<false, true, 2, T, B, true>)));   \
  TRY(tile_attr(ctx, (k_tile_lj2<false, false, 2, T, B, true>)));  \
  TRY(tile_attr(ctx, k_tile_lj2<true, false, 2, T, B>));
  LJ2_SHAPES(X)
#undef X
  return B200_OK;
}

static int launch_tile_lj2(b200_ctx *ctx, cudaStream_t s, int eflag, const int *ids, int ntiles,
                           bool fuse = false, int do_check = 0) {
  const TileGeom &G = ctx->tg;
  const int nl = ctx->nlocal, c = ctx->cur, scap = ctx->tile_scap;
  const bool one = ctx->ntypes == 1;
  const size_t sm = tile2_smem_bytes(scap, !one);
  const int T = ctx->lj2[0], B = ctx->lj2[1], ilp = (one && !eflag && !fuse) ? ctx->lj2[2] : 2;
  NveFuse nv;
  memset(&nv, 0, sizeof nv);
  if (fuse)
    nv = NveFuse{ctx->xt[c ^ 1], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->mask[c], ctx->mass_d.p,
                 ctx->dtv, ctx->dtf, ctx->groupbit, do_check, ctx->xh[0], ctx->xh[1], ctx->xh[2],
                 ctx->triggersq, ctx->flags};
  // no more threads than the fullest tile has atoms (whole warps)
  const int thr = std::max(32, std::min(T, cdiv(std::max(ctx->tile_maxown, 1), 32) * 32));
#define L2(EV, ONE, ILP, TT, BB)                                                                    \
  k_tile_lj2<EV, ONE, ILP, TT, BB><<<ntiles, thr, sm, s>>>(                                        \
      G, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI,             \
      ctx->tile_slots, ctx->tl_iloc.p, ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->f[0],      \
      ctx->f[1], ctx->f[2], ctx->lj_one, ctx->lj_tab.p, ctx->ntypes, ctx->ev, scap, ctx->tflags, ids, nv, \
      ctx->tile_hdrs.p)
#define L2N(ONE, TT, BB)                                                                            \
  k_tile_lj2<false, ONE, 2, TT, BB, true><<<ntiles, thr, sm, s>>>(                                  \
      G, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI,             \
      ctx->tile_slots, ctx->tl_iloc.p, ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->f[0],      \
      ctx->f[1], ctx->f[2], ctx->lj_one, ctx->lj_tab.p, ctx->ntypes, ctx->ev, scap, ctx->tflags, ids, nv, \
      ctx->tile_hdrs.p)
#define X(TT, BB)                                                      \
  if (T == TT && B == BB) {                                            \
    if (fuse) { if (one) L2N(true, TT, BB); else L2N(false, TT, BB); } \
    else if (one && !eflag) { if (ilp == 4) L2(false, true, 4, TT, BB); else L2(false, true, 2, TT, BB); } \
    else if (one) L2(true, true, 2, TT, BB);                           \
    else if (!eflag) L2(false, false, 2, TT, BB);                      \
    else L2(true, false, 2, TT, BB);                                   \
  } else
  LJ2_SHAPES(X)
  return ctx->fail(B200_EARG, "no k_tile_lj2 instance for %d threads x %d CTAs/SM", T, B);
#undef X
#undef L2N
#undef L2
  ctx->launches++;
  LAUNCH_CHECK();
  return B200_OK;
}

// mixed-precision lj/cut, second-generation kernel (k_tile_lj2f); B200_LJ2F=threads,minb
#define LJ2F_SHAPES(X) X(352, 2) X(352, 3) X(320, 3) X(256, 4) X(448, 2)
static int lj2f_attrs(b200_ctx *ctx) {
#define X(T, B)                                                    \
  TRY(tile_attr(ctx, k_tile_lj2f<false, true, T, B>));             \
  TRY(tile_attr(ctx, k_tile_lj2f<true, true, T, B>));              \
  TRY(tile_attr(ctx, k_tile_lj2f<false, false, T, B>));            \
  TRY(tile_attr(ctx, k_tile_lj2f<true, false, T, B>));             \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, true, T, B, true>)));     \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, false, T, B, true>)));    \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, true, T, B, true, true>)));   \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, false, T, B, true, true>)));  \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, true, T, B, false, true>)));  \
  TRY(tile_attr(ctx, (k_tile_lj2f<false, false, T, B, false, true>)));
  LJ2F_SHAPES(X)
#undef X
  return B200_OK;
}

// bring the fixed-point records up to date with xt[cur] (no-op unless the mixed kernel uses them)
static int q_refresh(b200_ctx *ctx, bool ghosts) {
  if (!(ctx->gq_ok && ctx->tiles_active && ctx->prec == B200_PREC_MIXED && ctx->mixed_fx && ctx->use_lj2))
    return B200_OK;
  const int c = ctx->cur, nl = ctx->nlocal, ng = ctx->nghost;
  if (!ctx->q_owned_valid && nl > 0) {
    k_xt_to_q<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(0, nl, ctx->xt[c], ctx->qrec[c], ctx->qgeom);
    ctx->launches++;
  }
  ctx->q_owned_valid = true;
  if (ghosts) {
    if (!ctx->q_ghost_valid && ng > 0) {
      k_xt_to_q<<<cdiv(ng, 256), 256, 0, ctx->stream>>>(nl, ng, ctx->xt[c], ctx->qrec[c], ctx->qgeom);
      ctx->launches++;
    }
    ctx->q_ghost_valid = true;
  }
  LAUNCH_CHECK();
  return B200_OK;
}

static int launch_tile_lj2f(b200_ctx *ctx, cudaStream_t s, int eflag, const int *ids, int ntiles, bool fuse,
                            int do_check) {
  const TileGeom &G = ctx->tg;
  const int nl = ctx->nlocal, c = ctx->cur, scap = ctx->tile_scap;
  const bool one = ctx->ntypes == 1;
  const size_t sm = tile2f_smem_bytes(scap);
  const int T = ctx->lj2f[0], B = ctx->lj2f[1];
  const int thr = std::max(32, std::min(T, cdiv(std::max(ctx->tile_maxown, 1), 32) * 32));
  NveFuse nv;
  memset(&nv, 0, sizeof nv);
  if (fuse)
    nv = NveFuse{ctx->xt[c ^ 1], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->mask[c], ctx->mass_d.p,
                 ctx->dtv, ctx->dtf, ctx->groupbit, do_check, ctx->xh[0], ctx->xh[1], ctx->xh[2],
                 ctx->triggersq, ctx->flags};
#define ARGS                                                                                        \
  G, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI, ctx->tile_slots, \
      ctx->tl_iloc.p, ctx->tl_num.p, ctx->tl_gi.p, ctx->tl_list.p, ctx->f[0], ctx->f[1], ctx->f[2],  \
      ctx->lj_one, ctx->lj_onef, ctx->lj_tab.p, ctx->lj_tabf.p, ctx->ntypes, ctx->ev, scap,         \
      ctx->tflags, ids, nv, ctx->tile_hdrs.p, ctx->qrec[c], ctx->qrec[c ^ 1], ctx->qgeom
#define X(TT, BB)                                                                        \
  if (T == TT && B == BB) {                                                              \
    if (fuse && ctx->gq_ok) {                                                            \
      if (one) k_tile_lj2f<false, true, TT, BB, true, true><<<ntiles, thr, sm, s>>>(ARGS);   \
      else k_tile_lj2f<false, false, TT, BB, true, true><<<ntiles, thr, sm, s>>>(ARGS);  \
    } else if (ctx->gq_ok && !eflag) {                                                   \
      if (one) k_tile_lj2f<false, true, TT, BB, false, true><<<ntiles, thr, sm, s>>>(ARGS);  \
      else k_tile_lj2f<false, false, TT, BB, false, true><<<ntiles, thr, sm, s>>>(ARGS); \
    } else if (fuse) {                                                                   \
      if (one) k_tile_lj2f<false, true, TT, BB, true><<<ntiles, thr, sm, s>>>(ARGS);     \
      else k_tile_lj2f<false, false, TT, BB, true><<<ntiles, thr, sm, s>>>(ARGS);        \
    } else if (one) {                                                                    \
      if (eflag) k_tile_lj2f<true, true, TT, BB><<<ntiles, thr, sm, s>>>(ARGS);          \
      else k_tile_lj2f<false, true, TT, BB><<<ntiles, thr, sm, s>>>(ARGS);               \
    } else {                                                                             \
      if (eflag) k_tile_lj2f<true, false, TT, BB><<<ntiles, thr, sm, s>>>(ARGS);         \
      else k_tile_lj2f<false, false, TT, BB><<<ntiles, thr, sm, s>>>(ARGS);              \
    }                                                                                    \
  } else
  LJ2F_SHAPES(X)
  return ctx->fail(B200_EARG, "no k_tile_lj2f instance for %d threads x %d CTAs/SM", T, B);
#undef X
#undef ARGS
  ctx->launches++;
  LAUNCH_CHECK();
  return B200_OK;
}

// lj/cut over `ntiles` tiles of the tile list (ids == nullptr: all tiles in order) on stream s
static int launch_tile_lj(b200_ctx *ctx, cudaStream_t s, int eflag, const int *ids, int ntiles) {
  if (ntiles <= 0) return B200_OK;
  const TileGeom &G = ctx->tg;
  const int nl = ctx->nlocal, c = ctx->cur, thr = ctx->tile_threads, scap = ctx->tile_scap;
  const size_t sm = tile_smem_bytes(scap, G.srow_y * G.srow_z, G.sbx, false, false);
 const bool one = ctx->ntypes == 1, mixed = ctx->prec == B200_PREC_MIXED;
  if (!mixed && ctx->use_lj2) return launch_tile_lj2(ctx, s, eflag, ids, ntiles, ctx->fuse_now, ctx->fuse_check);
  if (mixed && ctx->mixed_fx && ctx->use_lj2)
    return launch_tile_lj2f(ctx, s, eflag, ids, ntiles, ctx->fuse_now, ctx->fuse_check);
#define TLJ(EV, ONE, MX)                                                                            \
  k_tile_lj<EV, ONE, MX><<<ntiles, thr, sm, s>>>(                                                   \
      G, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI,             \
      ctx->tile_slots, ctx->tl_iloc.p, ctx->tl_num.p, ctx->tl_list.p, ctx->f[0], ctx->f[1],         \
      ctx->f[2], ctx->lj_one, ctx->lj_onef, ctx->lj_tab.p, ctx->lj_tabf.p, ctx->ntypes, ctx->ev,    \
      scap, ctx->tflags, ids)
#define TLJFX(EV, ONE)                                                                              \
  k_tile_lj_fx<EV, ONE><<<ntiles, thr, tile_smem_bytes_fx(scap), s>>>(                              \
      G, nl, ctx->xt[c], ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p, ctx->tile_NI,             \
      ctx->tile_slots, ctx->tl_iloc.p, ctx->tl_num.p, ctx->tl_list.p, ctx->f[0], ctx->f[1],         \
      ctx->f[2], ctx->lj_one, ctx->lj_onef, ctx->lj_tab.p, ctx->lj_tabf.p, ctx->ntypes, ctx->ev,    \
      scap, ctx->tflags, ids)
#define TLJ2(EV, ONE) if (mixed && ctx->mixed_fx) TLJFX(EV, ONE); else if (mixed) TLJ(EV, ONE, true); else TLJ(EV, ONE, false)
  if (one) { if (eflag) { TLJ2(true, true); } else { TLJ2(false, true); } }
  else     { if (eflag) { TLJ2(true, false); } else { TLJ2(false, false); } }
#undef TLJ2
#undef TLJFX
#undef TLJ
  ctx->launches++;
  LAUNCH_CHECK();
  return B200_OK;
}

// Can this step run the interior tiles beside the halo?  (per-phase profiling times one stream,
// so it keeps everything on it)
static bool overlap_active(const b200_ctx *ctx) {
  return ctx->overlap && ctx->tiles_active && ctx->pair_style == 1 && ctx->remote_mask &&
         !ctx->profiling && ctx->tile_nint > 0 && ctx->tile_nbnd > 0;
}

// fork: interior tiles on stream2, after everything enqueued on `stream` so far
static int pair_interior_async(b200_ctx *ctx, int eflag, int vflag) {
  TRY(q_refresh(ctx, false));
  if (eflag || vflag) CK(cudaMemsetAsync(ctx->ev, 0, 7 * sizeof(double), ctx->stream));
  CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  TRY(launch_tile_lj(ctx, ctx->stream2, eflag, ctx->tile_ids.p, ctx->tile_nint));
  CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  return B200_OK;
}

// part 0: the whole pair computation on `stream`; part 1: boundary tiles only (the interior
// tiles were forked by pair_interior_async, which also cleared the tallies); the join happens
// before the virial (it reads every force) or is left to the caller (*joined = false)
static int pair_compute(b200_ctx *ctx, int eflag, int vflag, int part = 0, bool *joined = nullptr) {
  TRY(flush_vops(ctx));
  const int nl = ctx->nlocal, ng = ctx->nghost, c = ctx->cur;
  cudaStream_t s = ctx->stream;
  const bool ev = eflag || vflag;
  const bool mixed = ctx->prec == B200_PREC_MIXED;
  if (joined) *joined = true;
  TRY(q_refresh(ctx, true));
  ctx->ghost_f_clean = false;  // (force_clear has run; the pair kernels write ghost forces now)
  if (part == 1) {
    const int ph6 = ph_begin(ctx, B200_PH_PAIR);
    TRY(launch_tile_lj(ctx, s, eflag, ctx->tile_ids.p + ctx->tile_nint, ctx->tile_nbnd));
    ph_end(ctx, ph6);
    if (vflag && nl + ng > 0 && !ctx->full_ghost) {
      CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
      const int gv = std::min(cdiv(nl + ng, 256), 148 * 8);
      k_virial_fdotr<<<gv, 256, 0, s>>>(nl + ng, ctx->xt[c], ctx->f[0], ctx->f[1], ctx->f[2], ctx->ev);
      ctx->launches++;
      LAUNCH_CHECK();
    } else if (joined)
      *joined = false;  // (FULLGHOST: the virial was tallied pairwise inside the tile kernels)
    return B200_OK;
  }
  if (ev) CK(cudaMemsetAsync(ctx->ev, 0, 7 * sizeof(double), s));
  const int ph6 = ph_begin(ctx, B200_PH_PAIR);
  double4 *xt = ctx->xt[c];
  const int *nn = ctx->numneigh.p, *nb = ctx->neigh.p;
  double *fx = ctx->f[0], *fy = ctx->f[1], *fz = ctx->f[2];
  const int T = ctx->tpa;
  const int grid = cdiv(std::max(nl, 1) * T, 128);  // T lanes per atom
  if (ctx->tiles_active) {
    const TileGeom &G = ctx->tg;
    const int rows = G.srow_y * G.srow_z, thr = ctx->tile_threads;
    const int *ib = ctx->tile_ibase.p;
    const unsigned short *il = ctx->tl_iloc.p, *tn = ctx->tl_num.p;
    const uint4 *tl = ctx->tl_list.p;
    const int NI = ctx->tile_NI, slots = ctx->tile_slots, scap = ctx->tile_scap;
    if (ctx->pair_style == 1) {
      TRY(launch_tile_lj(ctx, s, eflag, nullptr, G.ntiles));
    } else if (ctx->pair_style == 2 && ctx->eam2_active) {
      // kernels_eam2.cuh: density + embedding, fp forward halo, force (+ fix nve when fused)
      const int thr2 = std::max(32, std::min(352, cdiv(std::max(ctx->tile_maxown, 1), 32) * 32));
      const size_t smr = eam2_smem_bytes(scap, false), smf = eam2_smem_bytes(scap, true);
      const unsigned short *tf = ctx->tl_far.p;
      const int *tg2 = ctx->tl_gi.p;
#define ER(EV)                                                                                         \
  k_tile_eam2_rho<EV, 352, 2><<<G.ntiles, thr2, smr, s>>>(nl, xt, ib, NI, slots, il, tn, tf, tg2, tl,  \
                                                           ctx->eam, ctx->eam_one, ctx->rho, ctx->fp,   \
                                                           ctx->ev, ctx->flags + 1, scap, ctx->tflags,  \
                                                           nullptr, ctx->tile_hdrs.p)
      if (eflag) ER(true); else ER(false);
#undef ER
      TRY(forward_scalar(ctx, ctx->fp));
      NveFuse nv;
      memset(&nv, 0, sizeof nv);
      const bool fuse = ctx->fuse_now && !ev;
      if (fuse)
        nv = NveFuse{ctx->xt[c ^ 1], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->mask[c], ctx->mass_d.p,
                     ctx->dtv, ctx->dtf, ctx->groupbit, ctx->fuse_check, ctx->xh[0], ctx->xh[1], ctx->xh[2],
                     ctx->triggersq, ctx->flags};
#define EF(EV, NVE)                                                                                      \
  k_tile_eam2_force<EV, 352, 2, NVE><<<G.ntiles, thr2, smf, s>>>(nl, xt, ib, NI, slots, il, tn, tf, tg2, \
                                                                  tl, ctx->eam, ctx->eam_one, ctx->fp, fx, \
                                                                  fy, fz, ctx->ev, scap, ctx->tflags,      \
                                                                  nullptr, nv, ctx->tile_hdrs.p)
      if (fuse) EF(false, true);
      else if (ev) EF(true, false);
      else EF(false, false);
#undef EF
      ctx->launches += 2;
    } else if (ctx->pair_style == 2) {
      const size_t sm1 = tile_smem_bytes(scap, rows, G.sbx, false, false);
      const size_t sm3 = tile_smem_bytes(scap, rows, G.sbx, false, true);
      if (ng > 0) CK(cudaMemsetAsync(ctx->rho + nl, 0, sizeof(double) * ng, s));
      if (mixed)
        k_tile_eam_rho<true><<<G.ntiles, thr, sm1, s>>>(G, nl, xt, ctx->ostart.p, ctx->gstart.p, ib, NI, slots,
                                                        il, tn, tl, ctx->eam, ctx->eamf, ctx->rho, scap, ctx->tflags, nullptr);
      else
        k_tile_eam_rho<false><<<G.ntiles, thr, sm1, s>>>(G, nl, xt, ctx->ostart.p, ctx->gstart.p, ib, NI, slots,
                                                         il, tn, tl, ctx->eam, ctx->eamf, ctx->rho, scap, ctx->tflags, nullptr);
      {
        Vec3Ptr r{{ctx->rho, nullptr, nullptr}};
        TRY(reverse_halo<1>(ctx, r));
      }
      const int g2 = cdiv(std::max(nl, 1), 256);
      if (eflag)
        k_eam_embed<true><<<g2, 256, 0, s>>>(nl, xt, ctx->eam, ctx->rho, ctx->fp, ctx->ev, ctx->flags + 1);
      else
        k_eam_embed<false><<<g2, 256, 0, s>>>(nl, xt, ctx->eam, ctx->rho, ctx->fp, ctx->ev, ctx->flags + 1);
      TRY(forward_scalar(ctx, ctx->fp));
#define TEF(EV, MX)                                                                                    \
  k_tile_eam_force<EV, MX><<<G.ntiles, thr, sm3, s>>>(G, nl, xt, ctx->ostart.p, ctx->gstart.p, ib, NI, \
                                                      slots, il, tn, tl, ctx->eam, ctx->eamf, ctx->fp, \
                                                      fx, fy, fz, ctx->ev, scap, ctx->tflags, nullptr)
      if (eflag) { if (mixed) TEF(true, true); else TEF(true, false); }
      else       { if (mixed) TEF(false, true); else TEF(false, false); }
#undef TEF
      ctx->launches += 3;
    } else
      return ctx->fail(B200_EARG, "no pair style set");
    LAUNCH_CHECK();
    ph_end(ctx, ph6);
    if (vflag && nl + ng > 0 && !ctx->full_ghost) {
      const int ph7 = ph_begin(ctx, B200_PH_THERMO);
      const int gv = std::min(cdiv(nl + ng, 256), 148 * 8);
      k_virial_fdotr<<<gv, 256, 0, s>>>(nl + ng, xt, fx, fy, fz, ctx->ev);
      ctx->launches++;
      LAUNCH_CHECK();
      ph_end(ctx, ph7);
    }
    return B200_OK;
  }
  if (mixed) CK(cudaMemsetAsync(ctx->ff, 0, sizeof(float4) * (nl + ng), s));
// instantiate a kernel launch for the run-time lanes-per-atom value
#define TPA_SWITCH(LAUNCH)        \
  switch (T) {                    \
    case 1: { LAUNCH(1); } break; \
    case 4: { LAUNCH(4); } break; \
    case 8: { LAUNCH(8); } break; \
    default: { LAUNCH(2); } break; \
  }
  if (ctx->pair_style == 1) {
    const int n1 = ctx->ntypes + 1, n2 = n1 * n1;
    const bool one = ctx->ntypes == 1;
    if (!mixed) {
      const size_t sm = one ? 0 : sizeof(double) * 6 * n2;
      const double *tab = one ? nullptr : ctx->lj_tab.p;
#define LJ_K(EV, ONE, TT) \
  k_pair_lj<EV, ONE, TT><<<grid, 128, sm, s>>>(nl, ctx->nstride, xt, nn, nb, fx, fy, fz, ctx->lj_one, tab, ctx->ntypes, ctx->ev)
#define LJ_L(TT)                                                         \
  if (one) { if (eflag) LJ_K(true, true, TT); else LJ_K(false, true, TT); } \
  else     { if (eflag) LJ_K(true, false, TT); else LJ_K(false, false, TT); }
      TPA_SWITCH(LJ_L)
#undef LJ_L
#undef LJ_K
    } else {
      const size_t sm = one ? 0 : sizeof(double) * n2 + sizeof(float) * 5 * n2;
#define LJ_K(EV, ONE, TT)                                                                             \
  k_pair_lj_mixed<EV, ONE, TT><<<grid, 128, sm, s>>>(nl, ctx->nstride, xt, nn, nb, fx, fy, fz, ctx->ff, \
                                                     ctx->lj_one.cutsq, ctx->lj_onef, ctx->lj_tab.p,  \
                                                     ctx->lj_tabf.p, ctx->ntypes, ctx->ev)
#define LJ_L(TT)                                                         \
  if (one) { if (eflag) LJ_K(true, true, TT); else LJ_K(false, true, TT); } \
  else     { if (eflag) LJ_K(true, false, TT); else LJ_K(false, false, TT); }
      TPA_SWITCH(LJ_L)
#undef LJ_L
#undef LJ_K
    }
    ctx->launches++;
  } else if (ctx->pair_style == 2) {
    // every rank walks the same sequence of halo calls, even one without atoms
    CK(cudaMemsetAsync(ctx->rho, 0, sizeof(double) * (nl + ng), s));
    const bool fast1 = ctx->eam_one_ok && !mixed;
#define RHO_L(TT)                                                                                       \
  if (mixed) k_eam_rho_mixed<TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eamf, ctx->rho); \
  else if (fast1) k_eam_rho_one<TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eam_one, ctx->rho); \
  else k_eam_rho<TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->rho);
    TPA_SWITCH(RHO_L)
#undef RHO_L
    {
      Vec3Ptr r{{ctx->rho, nullptr, nullptr}};
      TRY(reverse_halo<1>(ctx, r));
    }
    const int g2 = cdiv(std::max(nl, 1), 256);
    if (eflag)
      k_eam_embed<true><<<g2, 256, 0, s>>>(nl, xt, ctx->eam, ctx->rho, ctx->fp, ctx->ev, ctx->flags + 1);
    else
      k_eam_embed<false><<<g2, 256, 0, s>>>(nl, xt, ctx->eam, ctx->rho, ctx->fp, ctx->ev, ctx->flags + 1);
    TRY(forward_scalar(ctx, ctx->fp));
#define FORCE_L(TT)                                                                                        \
  if (mixed) {                                                                                             \
    if (eflag) k_eam_force_mixed<true, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eamf, ctx->fp, fx, fy, fz, ctx->ff, ctx->ev); \
    else k_eam_force_mixed<false, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eamf, ctx->fp, fx, fy, fz, ctx->ff, ctx->ev); \
  } else if (fast1) {                                                                                      \
    if (eflag) k_eam_force_one<true, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eam_one, ctx->fp, fx, fy, fz, ctx->ev); \
    else k_eam_force_one<false, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->eam_one, ctx->fp, fx, fy, fz, ctx->ev); \
  } else {                                                                                                 \
    if (eflag) k_eam_force<true, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->fp, fx, fy, fz, ctx->ev); \
    else k_eam_force<false, TT><<<grid, 128, 0, s>>>(nl, ctx->nstride, xt, nn, nb, ctx->eam, ctx->fp, fx, fy, fz, ctx->ev); \
  }
    TPA_SWITCH(FORCE_L)
#undef FORCE_L
    ctx->launches += 3;
  } else
    return ctx->fail(B200_EARG, "no pair style set");
#undef TPA_SWITCH
  if (mixed && nl + ng > 0) {
    k_merge_ff<<<cdiv(nl + ng, 256), 256, 0, s>>>(nl + ng, nl, ctx->ff, fx, fy, fz);
    ctx->launches++;
  }
  LAUNCH_CHECK();
  ph_end(ctx, ph6);
  if (vflag && nl + ng > 0) {
    const int ph7 = ph_begin(ctx, B200_PH_THERMO);
    const int gv = std::min(cdiv(nl + ng, 256), 148 * 8);
    k_virial_fdotr<<<gv, 256, 0, s>>>(nl + ng, xt, fx, fy, fz, ctx->ev);
    ctx->launches++;
    LAUNCH_CHECK();
    ph_end(ctx, ph7);
  }
  return B200_OK;
}

static int initial_integrate(b200_ctx *ctx, int do_check) {
  TRY(flush_vops(ctx));
  if (!ctx->have_nve) return ctx->fail(B200_EARG, "b200_fix_nve has not been called");
  const int ph8 = ph_begin(ctx, B200_PH_INITIAL);
  const int nl = ctx->nlocal, c = ctx->cur;
  if (nl > 0) {
    if (ctx->pending_final)  // deferred final_integrate(n) + initial_integrate(n+1), one pass
      k_nve_final_initial<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(
          nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->f[0], ctx->f[1], ctx->f[2],
          ctx->mask[c], ctx->mass_d.p, ctx->dtv, ctx->dtf, ctx->groupbit, do_check, ctx->xh[0],
          ctx->xh[1], ctx->xh[2], ctx->triggersq, ctx->flags);
    else
      k_nve_initial<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(
          nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->f[0], ctx->f[1], ctx->f[2],
          ctx->mask[c], ctx->mass_d.p, ctx->dtv, ctx->dtf, ctx->groupbit, do_check, ctx->xh[0],
          ctx->xh[1], ctx->xh[2], ctx->triggersq, ctx->flags);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  ctx->pending_final = false;
  ctx->q_owned_valid = false;  // owned positions moved
  ph_end(ctx, ph8);
  return B200_OK;
}

static int final_integrate(b200_ctx *ctx) {
  TRY(flush_vops(ctx));
  const int ph9 = ph_begin(ctx, B200_PH_FINAL);
  const int nl = ctx->nlocal, c = ctx->cur;
  if (nl > 0) {
    k_nve_final<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1],
                                                        ctx->v[c][2], ctx->f[0], ctx->f[1], ctx->f[2],
                                                        ctx->mask[c], ctx->mass_d.p, ctx->dtf,
                                                        ctx->groupbit);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  ph_end(ctx, ph9);
  return B200_OK;
}

// A deferred final_integrate must be applied before anything reads or replaces v.
static int flush_final(b200_ctx *ctx) {
  if (ctx->ahead)
    return ctx->fail(B200_EARG, "atoms are one half-step ahead (fused integrator): a run did not end on a tallied step");
  TRY(flush_vops(ctx));  // queued operations were issued before whatever asks now
  if (!ctx->pending_final) return B200_OK;
  ctx->pending_final = false;
  return final_integrate(ctx);
}

// Neighbor::decide (neighbor.cpp:2408-2424); `moved` is the device vote of check_distance
static int decide(b200_ctx *ctx, int *rebuild) {
  TRY(flush_vops(ctx));  // the queued drift carries the displacement vote
  ctx->ago++;
  *rebuild = 0;
  if (ctx->ago >= ctx->delay && ctx->ago % ctx->every == 0) {
    if (ctx->build_once) return B200_OK;  // neighbor.cpp:2420
    if (!ctx->dist_check) {
      *rebuild = 1;
      return B200_OK;
    }
    if (ctx->box_changes) {  // the check was not fused into an integrate kernel: take it now
      CK(cudaMemsetAsync(ctx->flags, 0, sizeof(int), ctx->stream));
      if (ctx->nlocal > 0) {
        k_check_distance<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(
            ctx->nlocal, ctx->xt[ctx->cur], ctx->xh[0], ctx->xh[1], ctx->xh[2], ctx->deltasq, ctx->flags);
        ctx->launches++;
        LAUNCH_CHECK();
      }
    }
    // Neighbor::check_distance: MPI_Allreduce(MAX) of the moved flag (neighbor.cpp:2487)
    if (ctx->nranks > 1 && !ctx->grp)
      NK(g_nccl.AllReduce(ctx->flags, ctx->flags, 1, ncclInt, ncclMax, ctx->nccl, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_flags, ctx->flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->grp) {  // the vote of the whole group
      b200_group *g = ctx->grp;
      g->redi[ctx->rank] = ctx->h_flags[0];
      g->barrier();
      int m = 0;
      for (int r = 0; r < g->n; r++) m = std::max(m, g->redi[r]);
      g->barrier();
      ctx->h_flags[0] = m;
    }
    if (ctx->h_flags[0]) {
      if (ctx->ago == std::max(ctx->every, ctx->delay)) ctx->ndanger++;
      *rebuild = 1;
    }
  }
  return B200_OK;
}

static bool check_due_next(const b200_ctx *ctx) {
  const int64_t a = ctx->ago + 1;
  return ctx->dist_check && !ctx->build_once && a >= ctx->delay && a % ctx->every == 0;
}

static int ke_reduce(b200_ctx *ctx) {
  const int nl = ctx->nlocal, c = ctx->cur;
  CK(cudaMemsetAsync(ctx->ev + 7, 0, sizeof(double), ctx->stream));
  if (nl > 0) {
    const int grid = std::min(cdiv(nl, 256), 148 * 8);
    k_ke<<<grid, 256, 0, ctx->stream>>>(nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2],
                                        ctx->mask[c], ctx->mass_d.p, ctx->groupbit, ctx->ev);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  return B200_OK;
}

// sum of n <= 8 host doubles over the contexts of an in-process group (result on every context)
static int group_sum(b200_ctx *ctx, double *v, int n) {
  b200_group *g = ctx->grp;
  memcpy(&g->red[(size_t)ctx->rank * 8], v, sizeof(double) * n);
  g->barrier();
  for (int k = 0; k < n; k++) {
    double t = 0.0;
    for (int r = 0; r < g->n; r++) t += g->red[(size_t)r * 8 + k];
    v[k] = t;
  }
  g->barrier();
  return B200_OK;
}

static int fetch_ev(b200_ctx *ctx) {
  // MPI_Allreduce(SUM) of compute_pe / compute_pressure virial / compute_temp
  if (ctx->nranks > 1 && !ctx->grp && !ctx->local_tallies)
    NK(g_nccl.AllReduce(ctx->ev, ctx->ev, 8, ncclDouble, ncclSum, ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_ev, ctx->ev, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  // the device error word rides along: a peer-memory halo that timed out (bit 8), a lost atom or a
  // non-finite coordinate is reported on the next tally step, not only at the next rebuild
  CK(cudaMemcpyAsync(ctx->h_flags + 1, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h_flags[40] = 0;
  if (ctx->p2p && ctx->p2p_counter)
    CK(cudaMemcpyAsync(ctx->h_flags + 40, ctx->p2p_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->grp && !ctx->local_tallies) TRY(group_sum(ctx, ctx->h_ev, 8));
  ctx->eng_vdwl = ctx->h_ev[0];
  for (int k = 0; k < 6; k++) ctx->virial[k] = ctx->h_ev[1 + k];
  // (after the group's collective: all members reach it)
  return check_err_flags(ctx, ctx->h_flags[1] | (ctx->h_flags[40] ? 8 : 0));
}

// CUDA loads kernels lazily, on first launch.  The kernels that only run when atoms migrate
// between sub-domains are first needed inside the first rebuild of a RUN (setup has no leavers):
// load them when the communicator is created instead.
static int preload_rebuild_kernels(b200_ctx *ctx) {
  cudaFuncAttributes a;
  CK(cudaFuncGetAttributes(&a, k_pack_migrate));
  CK(cudaFuncGetAttributes(&a, k_unpack_migrate));
  CK(cudaFuncGetAttributes(&a, k_pbc_bin<true>));
  CK(cudaFuncGetAttributes(&a, k_pack_border));
  CK(cudaFuncGetAttributes(&a, k_ghost_make));
  CK(cudaFuncGetAttributes(&a, k_ghost_place));
  CK(cudaFuncGetAttributes(&a, k_permute_owned));
  CK(cudaFuncGetAttributes(&a, k_tile_split));
  CK(cudaFuncGetAttributes(&a, k_tile_count));
  return B200_OK;
}

// =====================================================================================
//                                        C ABI
// =====================================================================================
extern "C" {

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int b200_create(b200_ctx **out, int device, int precision) {
  if (!out) return B200_EARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return B200_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return B200_ECUDA;
  b200_ctx *ctx = new b200_ctx();
  ctx->device = device;
  ctx->prec = precision;
  if (const char *e = getenv("B200_TPA")) {  // tuning knob: lanes per atom in the pair kernels
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) ctx->tpa = v;
  }
  if (const char *e = getenv("B200_LIST"))  // "flat": int32 half list + RED scatter kernels
    ctx->list_mode = strcmp(e, "flat") == 0 ? 2 : (strcmp(e, "tile") == 0 ? 1 : 0);
  if (const char *e = getenv("B200_TILE")) {  // tuning knob: tile size in bins "tx,ty,tz"
    int t[3];
    if (sscanf(e, "%d,%d,%d", &t[0], &t[1], &t[2]) == 3 && t[0] > 0 && t[1] > 0 && t[2] > 0)
      for (int d = 0; d < 3; d++) ctx->tile_req[d] = t[d];
  }
  *out = ctx;
  if (precision != B200_PREC_DOUBLE && precision != B200_PREC_MIXED)
    return ctx->fail(B200_EARG, "unknown precision mode %d", precision);
  {
    // `stream` carries the step (and the halo); stream2 only ever runs interior tiles beside it
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, lo));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    if (const char *e = getenv("B200_OVERLAP")) ctx->overlap = atoi(e) != 0;
    if (const char *e = getenv("B200_GRAPH")) ctx->use_graph = atoi(e) != 0;
    if (const char *e = getenv("B200_MIXED_FX")) ctx->mixed_fx = atoi(e) != 0;
    if (const char *e = getenv("B200_FUSE")) ctx->fuse_nve = atoi(e) != 0;
    if (const char *e = getenv("B200_FUSE_MIN")) ctx->fuse_min_atoms = atoi(e);
    if (const char *e = getenv("B200_EAM2")) ctx->eam2 = strcmp(e, "auto") == 0 ? 2 : (atoi(e) != 0 ? 1 : 0);
    if (const char *e = getenv("B200_BUILD2")) ctx->build2 = atoi(e) != 0;
    if (const char *e = getenv("B200_STABLE")) ctx->stable_order = atoi(e) != 0;
    if (const char *e = getenv("B200_EAM2_MARGIN")) ctx->eam2_margin = atof(e);
    if (const char *e = getenv("B200_LJ2F")) {
      int a[2];
      if (sscanf(e, "%d,%d", &a[0], &a[1]) == 2) { ctx->lj2f[0] = a[0]; ctx->lj2f[1] = a[1]; }
    }
    if (const char *e = getenv("B200_LJ2")) {
      int a[3];
      const int k = sscanf(e, "%d,%d,%d", &a[0], &a[1], &a[2]);
      if (k == 1 && a[0] == 0) ctx->use_lj2 = false;
      else if (k == 3 && (a[2] == 2 || a[2] == 4))
        for (int d = 0; d < 3; d++) ctx->lj2[d] = a[d];
    }
  }
  TRY(dalloc(ctx, &ctx->ev, 8));
  TRY(dalloc(ctx, &ctx->ke7, 8));
  TRY(dalloc(ctx, &ctx->flags, 4));
  TRY(dalloc(ctx, &ctx->tflags, 8));
  TRY(tile_kernel_attrs(ctx));
  TRY(dalloc(ctx, &ctx->counts, 32));
  TRY(dalloc(ctx, &ctx->recvoffset, NDIR + 1));
  TRY(dalloc(ctx, &ctx->allcounts, 32));
  CK(cudaMallocHost((void **)&ctx->h_counts, 32 * sizeof(int)));
  for (int d = 0; d < NDIR; d++) ctx->nbr[d] = 0;
  memset(&ctx->owner, 0, sizeof ctx->owner);
  TRY(dalloc(ctx, &ctx->diroffset, NDIR + 1));
  TRY(dalloc(ctx, &ctx->cnt64, 1));
  CK(cudaMallocHost((void **)&ctx->h_ev, 24 * sizeof(double)));
  CK(cudaMallocHost((void **)&ctx->h_flags, 64 * sizeof(int)));
  CK(cudaMemsetAsync(ctx->ev, 0, 8 * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(ctx->flags, 0, 4 * sizeof(int), ctx->stream));
  memset(&ctx->eam, 0, sizeof ctx->eam);
  memset(&ctx->lj_one, 0, sizeof ctx->lj_one);
  return B200_OK;
}

void b200_destroy(b200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  auto F = [](auto *p) { if (p) cudaFree((void *)p); };
  for (int b = 0; b < 2; b++) {
    F(ctx->xt[b]); F(ctx->qrec[b]);
    for (int d = 0; d < 3; d++) F(ctx->v[b][d]);
    F(ctx->tag[b]); F(ctx->mask[b]); F(ctx->image[b]); F(ctx->atombin[b]);
  }
  for (int d = 0; d < 3; d++) { F(ctx->f[d]); F(ctx->xh[d]); }
  F(ctx->slot); F(ctx->rho); F(ctx->fp); F(ctx->ff); F(ctx->lj_tabf.p); F(ctx->eam_f.p);
  F(ctx->cutneighsq_d.p); F(ctx->mass_d.p); F(ctx->lang_d.p); F(ctx->lang_u.p); F(ctx->mask_s.p); F(ctx->mask_r.p); F(ctx->ostart.p); F(ctx->gstart.p); F(ctx->tilesum.p);
  F(ctx->sendlist.p); F(ctx->gsrc.p); F(ctx->gbin.p); F(ctx->gslot.p); F(ctx->gdir.p);
  F(ctx->gdir_tmp.p); F(ctx->gtmp.p); F(ctx->counts); F(ctx->diroffset); F(ctx->recvoffset); F(ctx->allcounts);
  F(ctx->senddir.p); F(ctx->gtag_tmp.p); F(ctx->gsrc_tmp.p); F(ctx->arena); F(ctx->pflags); F(ctx->p2p_counter);
  F(ctx->mig_send.p); F(ctx->mig_recv.p); F(ctx->ag_dev);
  if (ctx->ag_host) cudaFreeHost(ctx->ag_host);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->nccl && g_nccl.ok) g_nccl.CommDestroy(ctx->nccl);
  F(ctx->neigh.p);
  F(ctx->numneigh.p); F(ctx->lj_tab.p); F(ctx->eam_i.p); F(ctx->eam_d.p); F(ctx->eam_one_d.p); F(ctx->ev); F(ctx->ke7); F(ctx->flags);
  F(ctx->cnt64);
  F(ctx->tflags); F(ctx->tile_ibase.p); F(ctx->tl_iloc.p); F(ctx->tl_num.p); F(ctx->tl_list.p); F(ctx->tl_gi.p); F(ctx->tile_hdrs.p); F(ctx->tl_far.p); F(ctx->peratom.p); F(ctx->okey.p); F(ctx->gkey.p);
  if (ctx->h_ev) cudaFreeHost(ctx->h_ev);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  for (auto &r : ctx->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : ctx->evpool) cudaEventDestroy(e);
  F(ctx->tile_bflag.p); F(ctx->tile_bpos.p); F(ctx->tile_ids.p);
  if (ctx->step_graph) cudaGraphExecDestroy(ctx->step_graph);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200_set_box(b200_ctx *ctx, const double boxlo[3], const double boxhi[3], const int per[3]) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  for (int d = 0; d < 3; d++) {
    if (!(boxhi[d] > boxlo[d])) return ctx->fail(B200_EARG, "box hi <= lo in dim %d", d);
    ctx->boxlo[d] = boxlo[d];
    ctx->boxhi[d] = boxhi[d];
    ctx->prd[d] = boxhi[d] - boxlo[d];
    ctx->periodic[d] = per ? per[d] : 1;
  }
  ctx->tri = false;
  ctx->xy = ctx->xz = ctx->yz = 0.0;
  ctx->have_box = true;
  ctx->geom_ready = false;
  return B200_OK;
}

// Triclinic box (Domain::set_global_box, domain.cpp:263-290): boxlo/boxhi as above plus the tilt
// factors; `angstrom` = Force::angstrom of the unit style (the list rule's tolerance is
// 0.01 angstrom, npair_bin.cpp:59).  Runs on the flat half list (NPairBin<1,1,1,0,1>).
int b200_set_box_triclinic(b200_ctx *ctx, const double boxlo[3], const double boxhi[3], double xy, double xz,
                           double yz, const int per[3], double angstrom) {
  if (!ctx) return B200_EARG;
  TRY(b200_set_box(ctx, boxlo, boxhi, per));
  if (!std::isfinite(xy) || !std::isfinite(xz) || !std::isfinite(yz) || !(angstrom > 0.0))
    return ctx->fail(B200_EARG, "b200_set_box_triclinic: bad tilt factors");
  if (per) {  // Domain::set_initial_box, domain.cpp:203-206
    if ((xy != 0.0 && !per[0]) || (xz != 0.0 && !per[0]) || (yz != 0.0 && !per[1]))
      return ctx->fail(B200_EARG, "Triclinic box must be periodic in skewed dimensions");
  }
  ctx->tri = true;
  ctx->xy = xy;
  ctx->xz = xz;
  ctx->yz = yz;
  ctx->angstrom = angstrom;
  return B200_OK;
}

// Force::newton_pair (force.cpp, `newton` command).  off: the neighbour list holds every
// owned-ghost pair on both of its owners (npair_bin.cpp:126-131) and no force returns from ghosts.
int b200_set_newton(b200_ctx *ctx, int newton_pair) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  ctx->newton = newton_pair ? 1 : 0;
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_set_decomposition(b200_ctx *ctx, const int procgrid[3], const int myloc[3]) {
  if (!ctx) return B200_EARG;
  for (int d = 0; d < 3; d++) {
    if (procgrid[d] < 1 || myloc[d] < 0 || myloc[d] >= procgrid[d])
      return ctx->fail(B200_EARG, "bad decomposition in dim %d", d);
    ctx->procgrid[d] = procgrid[d];
    ctx->myloc[d] = myloc[d];
  }
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_set_rank_grid(b200_ctx *ctx, const int *grid2rank, int n) {
  if (!ctx || n < 0 || (n > 0 && !grid2rank)) return B200_EARG;
  ctx->rankmap.assign(grid2rank, grid2rank + n);
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_set_neighbor(b200_ctx *ctx, double skin, int every, int delay, int dist_check, int one) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  if (skin < 0 || every < 1 || delay < 0) return ctx->fail(B200_EARG, "Illegal neighbor settings");
  ctx->skin = skin;
  ctx->every = every;
  ctx->delay = delay;
  ctx->dist_check = dist_check ? 1 : 0;
  if (one > 0) ctx->one = one;
  ctx->geom_ready = false;
  return B200_OK;
}

// neigh_modify once yes|no and exclude type (neighbor.cpp:2727-2790): ex_type is Neighbor's
// symmetric table of excluded type pairs, [(ntypes+1)^2] flags, or NULL for none
int b200_neigh_modify(b200_ctx *ctx, int build_once, int ntypes, const int *ex_type) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  ctx->build_once = build_once != 0;
  ctx->ex_type.clear();
  if (ex_type) {
    if (ntypes < 1) return ctx->fail(B200_EARG, "b200_neigh_modify: bad type count");
    const int n1 = ntypes + 1;
    ctx->ex_type.assign(ex_type, ex_type + n1 * n1);
    for (int i = 1; i <= ntypes; i++)
      for (int j = 1; j <= ntypes; j++)
        if (ctx->ex_type[i * n1 + j] != ctx->ex_type[j * n1 + i])
          return ctx->fail(B200_EARG, "b200_neigh_modify: the exclusion table must be symmetric");
  }
  ctx->geom_ready = false;
  return B200_OK;
}

// neigh_modify exclude group g1 g2 (NPair::exclusion, npair.cpp:249-254): n pairs of group bits
// (Neighbor::ex1_bit / ex2_bit); n = 0 clears them
int b200_neigh_modify_groups(b200_ctx *ctx, int n, const int *bit1, const int *bit2) {
  if (!ctx || n < 0 || (n && (!bit1 || !bit2))) return B200_EARG;
  if (n > MAXEXGROUP) return ctx->fail(B200_EARG, "neigh_modify exclude group: at most %d group pairs", MAXEXGROUP);
  drop_step_graph(ctx);
  ctx->exg = ExGroups{0, {0}, {0}};
  ctx->exg.n = n;
  for (int m = 0; m < n; m++) {
    ctx->exg.bit1[m] = bit1[m];
    ctx->exg.bit2[m] = bit2[m];
  }
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_set_atoms(b200_ctx *ctx, int nlocal, int ntypes, const double *mass, const double *x,
                   const double *v, const int *type, const int *tag, const int *mask,
                   const int *image) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  if (nlocal < 0 || ntypes < 1 || !mass || (nlocal && (!x || !v || !type || !tag)))
    return ctx->fail(B200_EARG, "b200_set_atoms: bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  ctx->ntypes = ntypes;
  ctx->vops.n = 0;  // queued integrator operations belonged to the previous atoms
  ctx->vops_chk = 0;
  ctx->mass_h.assign(mass, mass + ntypes + 1);
  TRY(reserve(ctx, ctx->mass_d, (size_t)ntypes + 1));
  CK(cudaMemcpyAsync(ctx->mass_d.p, mass, sizeof(double) * (ntypes + 1), cudaMemcpyHostToDevice, s));
  ctx->nlocal = 0;
  ctx->nghost = 0;
  const int want = (int)(nlocal * 1.3) + 4096;
  if (want > ctx->nmax) TRY(alloc_atoms(ctx, want));
  ctx->cur = 0;
  const int n = nlocal;
  if (n > 0) {
    // stage AoS host arrays through device scratch (the alternate ping-pong buffers)
    double *stage = reinterpret_cast<double *>(ctx->xt[1]);  // 4*nmax doubles >= 3*n
    int *itmp = ctx->atombin[1];
    CK(cudaMemcpyAsync(stage, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(itmp, type, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    k_pack_xt<<<cdiv(n, 256), 256, 0, s>>>(n, stage, itmp, ctx->xt[0]);
    CK(cudaMemcpyAsync(stage, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
    k_aos_to_soa<<<cdiv(n, 256), 256, 0, s>>>(n, stage, ctx->v[0][0], ctx->v[0][1], ctx->v[0][2]);
    ctx->launches += 2;
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(ctx->tag[0], tag, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    if (mask)
      CK(cudaMemcpyAsync(ctx->mask[0], mask, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    else {  // group `all` (bit 0) for every atom
      k_fill_int<<<cdiv(n, 256), 256, 0, s>>>(n, 1, ctx->mask[0]);
      ctx->launches++;
    }
    if (image)
      CK(cudaMemcpyAsync(ctx->image[0], image, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    else {  // image flags 0 0 0 (biased by IMGMAX = 512, lmptype.h)
      const int img0 = (512 << IMG2BITS) | (512 << IMGBITS) | 512;
      k_fill_int<<<cdiv(n, 256), 256, 0, s>>>(n, img0, ctx->image[0]);
      ctx->launches++;
    }
    LAUNCH_CHECK();
    for (int d = 0; d < 3; d++) CK(cudaMemsetAsync(ctx->f[d], 0, sizeof(double) * n, s));
  }
  CK(cudaStreamSynchronize(s));
  ctx->nlocal = n;
  ctx->pending_final = false;
  ctx->setup_done = false;
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_get_counts(const b200_ctx *ctx, int *nlocal, int *nghost) {
  if (!ctx) return B200_EARG;
  if (nlocal) *nlocal = ctx->nlocal;
  if (nghost) *nghost = ctx->nghost;
  return B200_OK;
}

int b200_get_atoms(b200_ctx *ctx, int with_ghosts, double *x, double *v, double *f, int *type,
                   int *tag, int *mask, int *image) {
  if (!ctx) return B200_EARG;
  if (ctx->ahead && (x || v || f))
    return ctx->fail(B200_EARG, "atoms requested while the integrator runs ahead (b200_step_ahead with more != 0)");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  TRY(flush_final(ctx));
  const int c = ctx->cur, nl = ctx->nlocal, n = nl + (with_ghosts ? ctx->nghost : 0);
  if (n == 0) return B200_OK;
  double *stage = reinterpret_cast<double *>(ctx->xt[c ^ 1]);
  int *itmp = ctx->atombin[c ^ 1];
  if (x || type) {
    k_unpack_xt<<<cdiv(n, 256), 256, 0, s>>>(n, ctx->xt[c], x ? stage : nullptr, type ? itmp : nullptr);
    ctx->launches++;
    LAUNCH_CHECK();
    if (x) CK(cudaMemcpyAsync(x, stage, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    if (type) CK(cudaMemcpyAsync(type, itmp, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (v && nl > 0) {  // velocities exist for owned atoms only
    k_soa_to_aos<<<cdiv(nl, 256), 256, 0, s>>>(nl, ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], stage);
    ctx->launches++;
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(v, stage, sizeof(double) * 3 * nl, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (f) {
    k_soa_to_aos<<<cdiv(n, 256), 256, 0, s>>>(n, ctx->f[0], ctx->f[1], ctx->f[2], stage);
    ctx->launches++;
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(f, stage, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (tag) CK(cudaMemcpyAsync(tag, ctx->tag[c], sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  if (mask) CK(cudaMemcpyAsync(mask, ctx->mask[c], sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  if (image) CK(cudaMemcpyAsync(image, ctx->image[c], sizeof(int) * nl, cudaMemcpyDeviceToHost, s));
  // The staging area was the other position buffer.  Its owned records are rewritten by whoever
  // uses it next (the sort, the fused integrator), but its ghost records must keep the ghost
  // TYPES: the fused integrator makes it the live buffer and the forward halo only writes x, y, z
  // (k_ghost_place fills both buffers at a rebuild for the same reason).
  if ((x || v || f) && ctx->nghost > 0)
    CK(cudaMemcpyAsync(ctx->xt[c ^ 1] + nl, ctx->xt[c] + nl, sizeof(double4) * ctx->nghost,
                       cudaMemcpyDeviceToDevice, s));
  CK(cudaStreamSynchronize(s));
  return B200_OK;
}

int b200_pair_lj_cut(b200_ctx *ctx, int ntypes, const double *cutsq, const double *lj1,
                     const double *lj2, const double *lj3, const double *lj4,
                     const double *offset, const double special_lj[4]) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  if (ntypes < 1 || !cutsq || !lj1 || !lj2 || !lj3 || !lj4 || !offset)
    return ctx->fail(B200_EARG, "b200_pair_lj_cut: bad arguments");
  if (ntypes > 15) return ctx->fail(B200_EARG, "lj/cut/b200 supports at most 15 atom types");
  if (special_lj && (special_lj[1] != 1.0 || special_lj[2] != 1.0 || special_lj[3] != 1.0))
    ; /* atom_style atomic never sets special bits; factors other than 1 are never applied */
  CK(cudaSetDevice(ctx->device));
  const int n1 = ntypes + 1, n2 = n1 * n1;
  ctx->pair_style = 1;
  ctx->cutsq_h.assign(cutsq, cutsq + n2);
  std::vector<double> tab(6 * n2);
  const double *src[6] = {cutsq, lj1, lj2, lj3, lj4, offset};
  for (int t = 0; t < 6; t++) memcpy(&tab[t * n2], src[t], sizeof(double) * n2);
  TRY(reserve(ctx, ctx->lj_tab, tab.size()));
  CK(cudaMemcpyAsync(ctx->lj_tab.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int k = n1 + 1;  // [1][1]
  ctx->lj_one = LJOne{cutsq[k], lj1[k], lj2[k], lj3[k], lj4[k], offset[k]};
  ctx->lj_onef = LJOneF{(float)lj1[k], (float)lj2[k], (float)lj3[k], (float)lj4[k], (float)offset[k]};
  {
    std::vector<float> tf(5 * (size_t)n2);
    const double *srcf[5] = {lj1, lj2, lj3, lj4, offset};
    for (int t = 0; t < 5; t++)
      for (int q = 0; q < n2; q++) tf[(size_t)t * n2 + q] = (float)srcf[t][q];
    TRY(reserve(ctx, ctx->lj_tabf, tf.size()));
    CK(cudaMemcpy(ctx->lj_tabf.p, tf.data(), sizeof(float) * tf.size(), cudaMemcpyHostToDevice));
  }
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_pair_eam(b200_ctx *ctx, int ntypes, int nr, int nrho, double rdr, double rdrho,
                  double rhomax, double cutforcesq, const int *type2frho, const int *type2rhor,
                  const int *type2z2r, const double *scale, int nfrho, const double *frho_spline,
                  int nrhor, const double *rhor_spline, int nz2r, const double *z2r_spline) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  if (ntypes < 1 || nr < 2 || nrho < 2 || !type2frho || !type2rhor || !type2z2r || !scale ||
      !frho_spline || !rhor_spline || !z2r_spline)
    return ctx->fail(B200_EARG, "b200_pair_eam: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int n1 = ntypes + 1, n2 = n1 * n1;
  ctx->pair_style = 2;
  ctx->cutsq_h.assign(n2, cutforcesq);  // Pair::init: cutsq[i][j] = init_one()^2 = cutmax^2
  std::vector<int> ih(n1 + 2 * n2);
  memcpy(&ih[0], type2frho, sizeof(int) * n1);
  memcpy(&ih[n1], type2rhor, sizeof(int) * n2);
  memcpy(&ih[n1 + n2], type2z2r, sizeof(int) * n2);
  // spline knots keep the reference's 7-double (56-byte) stride: a 64-byte stride was measured
  // 30 % slower (r01d) -- the 32 lanes of a gather then hit only two L1 bank groups
  const size_t nf = (size_t)nfrho * (nrho + 1) * 7;
  const size_t kh = (size_t)nrhor * (nr + 1), kz = (size_t)nz2r * (nr + 1);
  const size_t o_frho = n2, o_rhor = o_frho + nf, o_z2r = o_rhor + kh * 7;
  std::vector<double> dh(o_z2r + kz * 7, 0.0);
  memcpy(&dh[0], scale, sizeof(double) * n2);
  memcpy(&dh[o_frho], frho_spline, sizeof(double) * nf);
  memcpy(&dh[o_rhor], rhor_spline, sizeof(double) * kh * 7);
  memcpy(&dh[o_z2r], z2r_spline, sizeof(double) * kz * 7);
  TRY(reserve(ctx, ctx->eam_i, ih.size()));
  TRY(reserve(ctx, ctx->eam_d, dh.size()));
  CK(cudaMemcpyAsync(ctx->eam_i.p, ih.data(), sizeof(int) * ih.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->eam_d.p, dh.data(), sizeof(double) * dh.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  EAMParams &P = ctx->eam;
  P.nr = nr; P.nrho = nrho; P.ntypes = ntypes;
  P.rdr = rdr; P.rdrho = rdrho; P.rhomax = rhomax; P.cutforcesq = cutforcesq;
  P.type2frho = ctx->eam_i.p;
  P.type2rhor = ctx->eam_i.p + n1;
  P.type2z2r = ctx->eam_i.p + n1 + n2;
  P.scale = ctx->eam_d.p;
  P.frho = ctx->eam_d.p + o_frho;
  P.rhor = ctx->eam_d.p + o_rhor;
  P.z2r = ctx->eam_d.p + o_z2r;
  ctx->eam_one_ok = false;
  {
    const char *e = getenv("B200_EAM_ONE");
    if (ntypes == 1 && !(e && atoi(e) == 0)) {
      // single element: pack {rhor value cubic} and {rhor cubic | z2r cubic} per knot
      const int tr = type2rhor[n1 + 1], tz = type2z2r[n1 + 1];
      const size_t K = (size_t)nr + 1;
      std::vector<double> pk(K * 12, 0.0);
      for (size_t m = 0; m < K; m++) {
        const double *a = rhor_spline + ((size_t)tr * K + m) * 7, *z = z2r_spline + ((size_t)tz * K + m) * 7;
        for (int c = 0; c < 4; c++) pk[m * 4 + c] = a[3 + c];
        double *f8 = &pk[K * 4 + m * 8];
        f8[0] = a[3]; f8[1] = a[4]; f8[2] = a[5]; f8[3] = 0.0;
        f8[4] = z[3]; f8[5] = z[4]; f8[6] = z[5]; f8[7] = z[6];
      }
      TRY(reserve(ctx, ctx->eam_one_d, pk.size()));
      CK(cudaMemcpy(ctx->eam_one_d.p, pk.data(), sizeof(double) * pk.size(), cudaMemcpyHostToDevice));
      ctx->eam_one.rho4 = reinterpret_cast<const double2 *>(ctx->eam_one_d.p);
      ctx->eam_one.frc8 = reinterpret_cast<const double2 *>(ctx->eam_one_d.p + K * 4);
      ctx->eam_one.scale = scale[n1 + 1];
      ctx->eam_one_ok = true;
    }
  }
  {  // mixed mode: float copies of the r-space splines, one 32-byte sector per knot
    std::vector<float> tf((kh + kz) * 8, 0.0f);
    for (size_t q = 0; q < kh; q++)
      for (int t = 0; t < 7; t++) tf[q * 8 + t] = (float)rhor_spline[q * 7 + t];
    for (size_t q = 0; q < kz; q++)
      for (int t = 0; t < 7; t++) tf[(kh + q) * 8 + t] = (float)z2r_spline[q * 7 + t];
    TRY(reserve(ctx, ctx->eam_f, tf.size()));
    CK(cudaMemcpy(ctx->eam_f.p, tf.data(), sizeof(float) * tf.size(), cudaMemcpyHostToDevice));
    ctx->eamf.rhor = ctx->eam_f.p;
    ctx->eamf.z2r = ctx->eam_f.p + kh * 8;
    ctx->eamf.rdr = (float)rdr;
  }
  ctx->geom_ready = false;
  return B200_OK;
}

int b200_fix_nve(b200_ctx *ctx, double dtv, double dtf, int groupbit) {
  if (!ctx) return B200_EARG;
  drop_step_graph(ctx);
  TRY(flush_final(ctx));  // a pending half-kick belongs to the old dtf
  ctx->dtv = dtv;
  ctx->dtf = dtf;
  ctx->groupbit = groupbit;
  ctx->have_nve = true;
  return B200_OK;
}

int b200_setup(b200_ctx *ctx, int eflag, int vflag) {
  if (!ctx) return B200_EARG;
  CK(cudaSetDevice(ctx->device));
  TRY(flush_final(ctx));
  TRY(setup_geometry(ctx));
  TRY(reneighbor(ctx));
  ctx->nbuilds = 0;  // Verlet::setup: neighbor->ncalls = 0 (verlet.cpp:131)
  ctx->ndanger = 0;
  TRY(force_clear(ctx));
  TRY(pair_compute(ctx, eflag, vflag));
  TRY(reverse_comm(ctx));
  if (eflag || vflag) TRY(fetch_ev(ctx));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->setup_done = true;
  return B200_OK;
}

int b200_initial_integrate(b200_ctx *ctx) {
  if (!ctx) return B200_EARG;
  TRY(flush_final(ctx));
  CK(cudaMemsetAsync(ctx->flags, 0, sizeof(int), ctx->stream));
  return initial_integrate(ctx, check_due_next(ctx) ? 1 : 0);
}
int b200_final_integrate(b200_ctx *ctx) {
  if (!ctx) return B200_EARG;
  TRY(flush_final(ctx));
  return final_integrate(ctx);
}
// ---- the per-atom loops of FixNH (fix nvt), driven stage by stage by the host's own FixNH code
// (lammps_pkg/B200/fix_nvt_b200.cpp): see kernels_step.cuh
static int staged_guard(b200_ctx *ctx) {
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "integrator stage before b200_setup");
  if (ctx->ahead) return ctx->fail(B200_EARG, "integrator stage inside a fused run");
  TRY(flush_final(ctx));
  CK(cudaSetDevice(ctx->device));
  return B200_OK;
}

// queue one per-atom integrator operation (or run it at once with `package b200 lazy no`)
static int push_vop(b200_ctx *ctx, int kind, int groupbit, double a0, double a1 = 0.0, double a2 = 0.0,
                    const double *more = nullptr, int nmore = 0) {
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "integrator stage before b200_setup");
  if (ctx->ahead) return ctx->fail(B200_EARG, "integrator stage inside a fused run");
  if (ctx->pending_final) TRY(flush_final(ctx));  // a deferred half-kick of b200_step comes first
  if (ctx->vops.n == VOP_MAX) TRY(flush_vops(ctx));
  VOps &q = ctx->vops;
  q.kind[q.n] = kind;
  q.groupbit[q.n] = groupbit;
  q.a[q.n][0] = a0;
  q.a[q.n][1] = a1;
  q.a[q.n][2] = a2;
  for (int k = 0; k < nmore && k < 9; k++) q.a[q.n][3 + k] = more[k];
  q.n++;
  if (!ctx->lazy_ops) TRY(flush_vops(ctx));
  return B200_OK;
}

int b200_nve_v(b200_ctx *ctx, double dtf, int groupbit) {
  if (!ctx) return B200_EARG;
  return push_vop(ctx, VOP_KICK, groupbit, dtf);
}

int b200_nve_x(b200_ctx *ctx, double dtv, int groupbit) {
  if (!ctx) return B200_EARG;
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "integrator stage before b200_setup");
  if (ctx->ahead) return ctx->fail(B200_EARG, "integrator stage inside a fused run");
  if (ctx->pending_final) TRY(flush_final(ctx));
  if (ctx->vops.n == VOP_MAX || ctx->vops_chk) TRY(flush_vops(ctx));  // room for it; one vote per queue
  // the vote decide() reads after this stage (with a changing box decide() takes it itself)
  const int chk = (check_due_next(ctx) && !ctx->box_changes) ? 1 : 0;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemsetAsync(ctx->flags, 0, sizeof(int), ctx->stream));
  ctx->vops_chk = chk;
  ctx->q_owned_valid = false;  // owned positions move
  return push_vop(ctx, VOP_DRIFT, groupbit, dtv);
}

int b200_scale_v(b200_ctx *ctx, double factor, int groupbit) {
  if (!ctx) return B200_EARG;
  return push_vop(ctx, VOP_SCALE, groupbit, factor);
}

int b200_scale_v3(b200_ctx *ctx, const double factor[3], int groupbit) {
  if (!ctx || !factor) return B200_EARG;
  return push_vop(ctx, VOP_SCALE3, groupbit, factor[0], factor[1], factor[2]);
}

// FixLangevin::post_force (fix_langevin.cpp:383-507) on the stored forces of this step: see
// k_langevin.  gfactor1 / gfactor2_tsqrt: per-type prefactors [ntypes+1] as FixLangevin::init
// computes them, the second already multiplied by sqrt(t_target) of this step.  fsum (nullable):
// the group's summed random force comes back (zero yes).
int b200_langevin(b200_ctx *ctx, int ntypes, const double *gfactor1, const double *gfactor2_tsqrt, int groupbit,
                  uint64_t seed, int64_t step, const double *uniforms_by_tag, int64_t nuniform, double *fsum) {
  if (!ctx || !gfactor1 || !gfactor2_tsqrt) return B200_EARG;
  TRY(staged_guard(ctx));
  if (ntypes != ctx->ntypes) return ctx->fail(B200_EARG, "b200_langevin: ntypes %d != %d", ntypes, ctx->ntypes);
  const int nl = ctx->nlocal, c = ctx->cur;
  cudaStream_t s = ctx->stream;
  const size_t nt = (size_t)ntypes + 1;
  TRY(reserve(ctx, ctx->lang_d, 2 * nt + 4));
  ctx->lang_h.assign(2 * nt, 0.0);
  memcpy(ctx->lang_h.data(), gfactor1, nt * sizeof(double));
  memcpy(ctx->lang_h.data() + nt, gfactor2_tsqrt, nt * sizeof(double));
  CK(cudaMemcpyAsync(ctx->lang_d.p, ctx->lang_h.data(), 2 * nt * sizeof(double), cudaMemcpyHostToDevice, s));
  double *dsum = fsum ? ctx->lang_d.p + 2 * nt : nullptr;
  if (dsum) CK(cudaMemsetAsync(dsum, 0, 3 * sizeof(double), s));
  const double *duni = nullptr;
  if (uniforms_by_tag && nuniform > 0) {
    TRY(reserve(ctx, ctx->lang_u, (size_t)nuniform));
    CK(cudaMemcpyAsync(ctx->lang_u.p, uniforms_by_tag, sizeof(double) * nuniform, cudaMemcpyHostToDevice, s));
    duni = ctx->lang_u.p;
  }
  if (nl > 0) {
    k_langevin<<<std::min(cdiv(nl, 256), 148 * 8), 256, 0, s>>>(
        nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->f[0], ctx->f[1], ctx->f[2], ctx->tag[c],
        ctx->mask[c], groupbit, ntypes, ctx->lang_d.p, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step,
        (uint32_t)((uint64_t)step >> 32), duni, (long long)nuniform, dsum);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  if (fsum) {
    CK(cudaMemcpyAsync(ctx->h_ev + 16, dsum, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(fsum, ctx->h_ev + 16, 3 * sizeof(double));
  } else if (duni)
    CK(cudaStreamSynchronize(s));  // the caller's uniforms may be reused
  return B200_OK;
}

// f_i += df for the owned atoms of a group (FixLangevin zero yes, fix_langevin.cpp:481-497)
int b200_add_force(b200_ctx *ctx, const double df[3], int groupbit) {
  if (!ctx || !df) return B200_EARG;
  TRY(staged_guard(ctx));
  const int nl = ctx->nlocal, c = ctx->cur;
  if (nl > 0) {
    k_add_force<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(nl, ctx->f[0], ctx->f[1], ctx->f[2], ctx->mask[c], groupbit,
                                                        df[0], df[1], df[2]);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  return B200_OK;
}

// The box follows the barostat.  Owned atoms of `groupbit` are dilated with it (FixNH::remap,
// fix_nh.cpp:1156-1300: x2lamda in the old box, lamda2x in the new one); then the new box is
// adopted: periodic shifts of the halo at once (ghosts follow at the next forward halo), bins,
// stencil, slabs and sub-domain bounds at the next rebuild (Verlet::run, verlet.cpp:300-304).
int b200_remap(b200_ctx *ctx, const double oldlo[3], const double oldhi[3], const double newlo[3],
               const double newhi[3], int groupbit) {
  if (!ctx || !oldlo || !oldhi || !newlo || !newhi) return B200_EARG;
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "integrator stage before b200_setup");
  if (!ctx->geom_ready && !ctx->box_changes) return ctx->fail(B200_EARG, "b200_remap before b200_setup");
  if (ctx->tri) return ctx->fail(B200_EARG, "b200_remap: a changing triclinic box is not supported");
  RemapBox B;
  for (int d = 0; d < 3; d++) {
    if (!(newhi[d] > newlo[d])) return ctx->fail(B200_EARG, "box hi <= lo in dim %d", d);
    B.oldlo[d] = oldlo[d];
    B.oldhinv[d] = 1.0 / (oldhi[d] - oldlo[d]);  // Domain::set_global_box: h_inv = 1/prd
    B.newlo[d] = newlo[d];
    B.newh[d] = newhi[d] - newlo[d];
  }
  {
    // the dilation of the atoms joins the queued integrator operations (it sits between the
    // half-kick and the drift of FixNH::initial_integrate); the box itself is adopted at once
    const double rest[9] = {B.oldhinv[0], B.oldhinv[1], B.oldhinv[2], B.newlo[0], B.newlo[1], B.newlo[2],
                            B.newh[0],    B.newh[1],    B.newh[2]};
    TRY(push_vop(ctx, VOP_REMAP, groupbit, B.oldlo[0], B.oldlo[1], B.oldlo[2], rest, 9));
  }
  ctx->q_owned_valid = false;
  drop_step_graph(ctx);
  if (!ctx->box_changes) {  // first call: the box the list in force was built in
    for (int d = 0; d < 3; d++) {
      ctx->boxlo_hold[d] = ctx->boxlo[d];
      ctx->boxhi_hold[d] = ctx->boxhi[d];
    }
    ctx->box_changes = true;
  }
  Geom &g = ctx->geom;
  for (int d = 0; d < 3; d++) {
    ctx->boxlo[d] = newlo[d];
    ctx->boxhi[d] = newhi[d];
    ctx->prd[d] = newhi[d] - newlo[d];
    g.boxlo[d] = ctx->boxlo[d];
    g.boxhi[d] = ctx->boxhi[d];
    g.prd[d] = ctx->prd[d];
    for (int dir = 0; dir < NDIR; dir++) {
      if (g.shift[dir][d] != 0.0) g.shift[dir][d] = g.shift[dir][d] > 0.0 ? ctx->prd[d] : -ctx->prd[d];
      g.fshift[dir][d] = g.shift[dir][d];
    }
  }
  ctx->geom_ready = false;  // bins, stencil, slabs, sub-domain bounds: at the next rebuild
  // Neighbor::check_distance, neighbor.cpp:2443-2455: the trigger shrinks by the corner motion
  double d1 = 0.0, d2 = 0.0;
  for (int d = 0; d < 3; d++) {
    const double a = ctx->boxlo[d] - ctx->boxlo_hold[d], b = ctx->boxhi[d] - ctx->boxhi_hold[d];
    d1 += a * a;
    d2 += b * b;
  }
  double delta = 0.5 * (ctx->skin - (std::sqrt(d1) + std::sqrt(d2)));
  if (delta < 0.0) delta = 0.0;
  ctx->deltasq = delta * delta;
  return B200_OK;
}

int b200_decide(b200_ctx *ctx, int *rebuild) { return (ctx && rebuild) ? decide(ctx, rebuild) : B200_EARG; }
int b200_forward_comm(b200_ctx *ctx) { return ctx ? forward_comm(ctx) : B200_EARG; }
int b200_reverse_comm(b200_ctx *ctx) { return ctx ? reverse_comm(ctx) : B200_EARG; }
int b200_reneighbor(b200_ctx *ctx) { return ctx ? reneighbor(ctx) : B200_EARG; }
int b200_force_clear(b200_ctx *ctx) { return ctx ? force_clear(ctx) : B200_EARG; }
int b200_pair_compute(b200_ctx *ctx, int eflag, int vflag) {
  if (!ctx) return B200_EARG;
  TRY(pair_compute(ctx, eflag, vflag));
  if (eflag || vflag) TRY(fetch_ev(ctx));
  return B200_OK;
}

// one iteration of Verlet::run (verlet.cpp:246-355) without the output stage
// A plain timestep: no tallies, no displacement check due, no rebuild due (known on the host
// when the schedule is `check no`), one sub-domain, final_integrate of the previous step
// pending so that the fused integrate kernel opens the step.
static bool plain_step(const b200_ctx *ctx, int eflag, int vflag) {
  // (`check yes` schedules rebuild irregularly and often; re-instantiating the graph after every
  // rebuild then costs more than it saves -- measured on bench/in.eam)
  if (!ctx->use_graph || eflag || vflag || ctx->profiling || ctx->nranks > 1 || ctx->remote_mask ||
      !ctx->pending_final || ctx->nlocal <= 0 || ctx->dist_check || ctx->ahead)
    return false;
  const int64_t a = ctx->ago + 1;
  const bool due = !ctx->build_once && a >= ctx->delay && a % ctx->every == 0;  // Neighbor::decide
  return !due;  // with `check yes` a due step needs the device vote; with `check no` it rebuilds
}

static int one_step(b200_ctx *ctx, int eflag, int vflag, int *rebuilt, bool allow_fuse = false);

// may this step's pair kernel carry fix nve (k_tile_lj2<..., NVE>)?  Static part of the answer;
// the list kind is known only after the step's rebuild decision
static bool fuse_candidate(const b200_ctx *ctx, int eflag, int vflag, bool allow_fuse) {
  const bool lj = ctx->pair_style == 1 && (ctx->prec == B200_PREC_DOUBLE || ctx->mixed_fx) && ctx->use_lj2;
  return allow_fuse && !eflag && !vflag && ctx->fuse_nve && ctx->have_nve && (lj || eam2_usable(ctx)) &&
         ctx->nlocal > 0 && (ctx->nranks > 1 || ctx->nlocal >= ctx->fuse_min_atoms);
}

static int graph_step(b200_ctx *ctx) {
  if (!ctx->step_graph) {
    // capture this very step: the same host code path, recorded instead of launched
    cudaGraph_t g = nullptr;
    const int64_t l0 = ctx->launches;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = initial_integrate(ctx, 0);
    if (rc == B200_OK) rc = forward_comm(ctx);
    if (rc == B200_OK) rc = force_clear(ctx);
    if (rc == B200_OK) rc = pair_compute(ctx, 0, 0);
    if (rc == B200_OK) rc = reverse_comm(ctx);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
    if (rc != B200_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    CK(ce);
    ctx->graph_launches = (int)(ctx->launches - l0);
    ctx->launches = l0;
    const cudaError_t ie = cudaGraphInstantiate(&ctx->step_graph, g, 0);
    cudaGraphDestroy(g);
    CK(ie);
  }
  CK(cudaGraphLaunch(ctx->step_graph, ctx->stream));
  ctx->launches += ctx->graph_launches;
  ctx->ago++;
  ctx->pending_final = true;
  return B200_OK;
}

static int one_step(b200_ctx *ctx, int eflag, int vflag, int *rebuilt, bool allow_fuse) {
  TRY(flush_vops(ctx));
  const bool fcand = fuse_candidate(ctx, eflag, vflag, allow_fuse);
  if (!fcand && plain_step(ctx, eflag, vflag)) {
    if (rebuilt) *rebuilt = 0;
    return graph_step(ctx);
  }
  if (ctx->ahead) {
    // the previous step's pair kernel already applied this step's initial_integrate (and its
    // displacement check)
    ctx->ahead = false;
  } else {
    const int chk = check_due_next(ctx) ? 1 : 0;
    if (chk) CK(cudaMemsetAsync(ctx->flags, 0, sizeof(int), ctx->stream));
    TRY(initial_integrate(ctx, chk));
  }
  int nflag = 0;
  TRY(decide(ctx, &nflag));
  if (nflag) TRY(reneighbor(ctx));
  const bool fuse = fcand && ctx->tiles_active && ctx->full_ghost && (ctx->pair_style == 1 || ctx->eam2_active);
  if (fuse) {
    ctx->fuse_check = check_due_next(ctx) ? 1 : 0;  // for the NEXT step's decide()
    if (ctx->fuse_check) CK(cudaMemsetAsync(ctx->flags, 0, sizeof(int), ctx->stream));
    ctx->fuse_now = true;
  }
  if (overlap_active(ctx)) {
    // interior tiles start now; halo, boundary tiles and the reverse halo overlap with them
    bool joined = false;
    TRY(pair_interior_async(ctx, eflag, vflag));
    if (!nflag) TRY(forward_comm(ctx));
    TRY(force_clear(ctx));
    TRY(pair_compute(ctx, eflag, vflag, 1, &joined));
    TRY(reverse_comm(ctx));
    if (!joined) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  } else {
    if (!nflag) TRY(forward_comm(ctx));
    TRY(force_clear(ctx));
    TRY(pair_compute(ctx, eflag, vflag));
    TRY(reverse_comm(ctx));
  }
  if (fuse) {
    // the pair kernels wrote x(n+1) into the other position buffer: it is the live one now
    ctx->fuse_now = false;
    std::swap(ctx->xt[ctx->cur], ctx->xt[ctx->cur ^ 1]);
    // the fused epilogue of the mixed kernel wrote the owned atoms' fixed-point records as well
    std::swap(ctx->qrec[ctx->cur], ctx->qrec[ctx->cur ^ 1]);
    ctx->q_owned_valid = ctx->gq_ok && ctx->prec == B200_PREC_MIXED;
    ctx->q_ghost_valid = false;
    ctx->ahead = true;
    ctx->pending_final = false;
    ctx->ghost_f_clean = true;
    if (rebuilt) *rebuilt = nflag;
    return B200_OK;
  }
  // FixNVE::final_integrate: on steps whose velocities nobody reads (no tallies) it is deferred
  // and fused with the next step's initial_integrate (k_nve_final_initial)
  if (eflag || vflag)
    TRY(final_integrate(ctx));
  else
    ctx->pending_final = true;
  if (rebuilt) *rebuilt = nflag;
  return B200_OK;
}

int b200_step(b200_ctx *ctx, int eflag, int vflag, int *rebuilt) {
  return b200_step_ahead(ctx, eflag, vflag, 0, rebuilt);
}

// b200_step for a host that knows what comes next: `more` != 0 promises that another step follows
// before anything reads atoms, velocities or forces.  The engine may then apply the next step's
// initial_integrate inside this step's pair kernel (NveFuse): the state it leaves is x(n+1),
// v(n+1/2), no forces -- exactly what the next step starts from, and nothing a host may look at
// (b200_get_atoms and the kinetic-energy sums refuse while the integrator runs ahead).
int b200_step_ahead(b200_ctx *ctx, int eflag, int vflag, int more, int *rebuilt) {
  if (!ctx) return B200_EARG;
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "b200_step before b200_setup");
  TRY(one_step(ctx, eflag, vflag, rebuilt, /*allow_fuse=*/more != 0));
  if (eflag || vflag) TRY(fetch_ev(ctx));  // the host is about to read eng_vdwl / virial
  return B200_OK;
}

int b200_run(b200_ctx *ctx, int nsteps, int64_t first_step, int thermo_every, double *thermo_out,
             int max_thermo, int *n_thermo) {
  if (!ctx) return B200_EARG;
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "b200_run before b200_setup");
  CK(cudaSetDevice(ctx->device));
  int nout = 0;
  if (!ctx->run_a) {
    CK(cudaEventCreate(&ctx->run_a));
    CK(cudaEventCreate(&ctx->run_b));
  }
  CK(cudaEventRecord(ctx->run_a, ctx->stream));
  for (int sidx = 1; sidx <= nsteps; sidx++) {
    const int64_t step = first_step + sidx;
    const int ev = (thermo_every > 0 && step % thermo_every == 0) || sidx == nsteps;
    int nflag = 0;
    TRY(one_step(ctx, ev, ev, &nflag, /*allow_fuse=*/sidx < nsteps));
    if (ev) {
      const int ph10 = ph_begin(ctx, B200_PH_THERMO);
      TRY(ke_reduce(ctx));
      ph_end(ctx, ph10);
      TRY(fetch_ev(ctx));
      if (thermo_out && nout < max_thermo) {
        double *t = thermo_out + 10 * (size_t)nout;
        t[0] = (double)step;
        t[1] = ctx->h_ev[7];
        t[2] = ctx->eng_vdwl;
        for (int k = 0; k < 6; k++) t[3 + k] = ctx->virial[k];
        t[9] = 0.0;
        nout++;
      }
    }
  }
  TRY(flush_final(ctx));  // normally a no-op: the last step of a run tallies
  CK(cudaEventRecord(ctx->run_b, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->run_a, ctx->run_b));
    ctx->last_run_ms = ms;
  }
  if (n_thermo) *n_thermo = nout;
  return B200_OK;
}

int b200_last_run_ms(b200_ctx *ctx, double *ms) {
  if (!ctx || !ms) return B200_EARG;
  *ms = ctx->last_run_ms;
  return B200_OK;
}

int b200_get_tallies(b200_ctx *ctx, double *eng_vdwl, double virial[6]) {
  if (!ctx) return B200_EARG;
  if (eng_vdwl) *eng_vdwl = ctx->eng_vdwl;
  if (virial) memcpy(virial, ctx->virial, sizeof(double) * 6);
  return B200_OK;
}

int b200_ke_sum(b200_ctx *ctx, double *mv2) {
  if (!ctx || !mv2) return B200_EARG;
  if (ctx->ahead) return ctx->fail(B200_EARG, "kinetic energy requested while the integrator runs ahead");
  TRY(flush_final(ctx));
  TRY(ke_reduce(ctx));
  if (ctx->nranks > 1 && !ctx->grp && !ctx->local_tallies)
    NK(g_nccl.AllReduce(ctx->ev + 7, ctx->ev + 7, 1, ncclDouble, ncclSum, ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_ev + 7, ctx->ev + 7, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->grp && !ctx->local_tallies) TRY(group_sum(ctx, ctx->h_ev + 7, 1));
  *mv2 = ctx->h_ev[7];
  return B200_OK;
}

// sum m v^2 and the kinetic tensor sums of the atoms in `groupbit` (ComputeTemp,
// compute_temp.cpp:73-140); summed over all sub-domains unless tallies are local
int b200_ke_group(b200_ctx *ctx, int groupbit, double *mv2, double tensor[6]) {
  if (!ctx) return B200_EARG;
  if (ctx->ahead) return ctx->fail(B200_EARG, "kinetic energy requested while the integrator runs ahead");
  if (ctx->pending_final) TRY(flush_final(ctx));
  CK(cudaSetDevice(ctx->device));
  const int nl = ctx->nlocal, c = ctx->cur;
  double *acc = ctx->ke7;
  CK(cudaMemsetAsync(acc, 0, 7 * sizeof(double), ctx->stream));
  if (ctx->vops.n > 0) {
    // queued integrator operations (the half-kick a Nose-Hoover chain takes before it reads the
    // temperature): one pass applies them and sums the kinetic energy of the result
    TRY(flush_vops(ctx, groupbit, acc));
  } else if (nl > 0) {
    k_ke_group<<<std::min(cdiv(nl, 256), 148 * 8), 256, 0, ctx->stream>>>(
        nl, ctx->xt[c], ctx->v[c][0], ctx->v[c][1], ctx->v[c][2], ctx->mask[c], ctx->mass_d.p, groupbit, acc);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  if (ctx->nranks > 1 && !ctx->grp && !ctx->local_tallies)
    NK(g_nccl.AllReduce(acc, acc, 7, ncclDouble, ncclSum, ctx->nccl, ctx->stream));
  double h[8];
  CK(cudaMemcpyAsync(ctx->h_ev + 8, acc, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(h, ctx->h_ev + 8, 7 * sizeof(double));
  if (ctx->grp && !ctx->local_tallies) TRY(group_sum(ctx, h, 7));
  if (mv2) *mv2 = h[0];
  if (tensor) memcpy(tensor, h + 1, 6 * sizeof(double));
  return B200_OK;
}

int b200_sync(b200_ctx *ctx) {
  if (!ctx) return B200_EARG;
  CK(cudaSetDevice(ctx->device));
  TRY(flush_vops(ctx));
  if (ctx->stream2) CK(cudaStreamSynchronize(ctx->stream2));
  // the device error word comes back with the sync: a peer-memory halo that gave up waiting for a
  // neighbour (P2P_SPIN_LIMIT), a lost atom, a non-finite coordinate
  if (!(ctx->flags && ctx->h_flags)) {
    CK(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
  }
  CK(cudaMemcpyAsync(ctx->h_flags + 1, ctx->flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->h_flags[40] = 0;
  if (ctx->p2p && ctx->p2p_counter)  // the peer-memory halo kernels keep their own error word
    CK(cudaMemcpyAsync(ctx->h_flags + 40, ctx->p2p_counter + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return check_err_flags(ctx, ctx->h_flags[1] | (ctx->h_flags[40] ? 8 : 0));
}

// run-time knobs by name (the `package b200` keywords; the B200_* environment variables set the
// same fields at b200_create).  Takes effect at the next rebuild / setup.
int b200_set_option(b200_ctx *ctx, const char *key, const char *value) {
  if (!ctx || !key || !value) return B200_EARG;
  drop_step_graph(ctx);
  const std::string k = key, v = value;
  auto yes = [&]() { return v == "yes" || v == "on" || v == "1" || v == "true"; };
  if (k == "list") {
    if (v == "tile") ctx->list_mode = 1;
    else if (v == "flat") ctx->list_mode = 2;
    else if (v == "auto") ctx->list_mode = 0;
    else return ctx->fail(B200_EARG, "package b200 list: tile, flat or auto (got %s)", value);
    ctx->geom_ready = false;
  } else if (k == "tile") {
    int t[3];
    if (sscanf(value, "%d,%d,%d", &t[0], &t[1], &t[2]) != 3 || t[0] < 1 || t[1] < 1 || t[2] < 1)
      return ctx->fail(B200_EARG, "package b200 tile: three positive bin counts (got %s)", value);
    for (int d = 0; d < 3; d++) ctx->tile_req[d] = t[d];
    ctx->geom_ready = false;
  } else if (k == "overlap") ctx->overlap = yes();
  else if (k == "graph") ctx->use_graph = yes();
  else if (k == "mixed_fx") ctx->mixed_fx = yes();
  else if (k == "fuse") ctx->fuse_nve = yes();
  else if (k == "eam2") { ctx->eam2 = v == "auto" ? 2 : (yes() ? 1 : 0); ctx->geom_ready = false; }
  else if (k == "build2") ctx->build2 = yes();
  else if (k == "lazy") { TRY(flush_vops(ctx)); ctx->lazy_ops = yes(); }
  else if (k == "tpa") {
    const int t = atoi(value);
    if (t != 1 && t != 2 && t != 4 && t != 8) return ctx->fail(B200_EARG, "package b200 tpa: 1, 2, 4 or 8");
    ctx->tpa = t;
    ctx->geom_ready = false;
  } else if (k == "tallies") {
    if (v == "local") ctx->local_tallies = true;
    else if (v == "global") ctx->local_tallies = false;
    else return ctx->fail(B200_EARG, "tallies: local or global");
  } else
    return ctx->fail(B200_EARG, "unknown option %s", key);
  return B200_OK;
}

int b200_get_stats(b200_ctx *ctx, b200_stats *out) {
  if (!ctx || !out) return B200_EARG;
  memset(out, 0, sizeof *out);
  out->nbuilds = ctx->nbuilds;
  out->ndanger = ctx->ndanger;
  out->ago = ctx->ago;
  out->maxneigh = ctx->maxneigh;
  out->max_numneigh = ctx->max_numneigh;
  for (int d = 0; d < 3; d++) out->nbins[d] = ctx->geom_ready ? ctx->geom.nbin[d] : 0;
  out->mbins = ctx->geom_ready ? ctx->geom.mbins : 0;
  out->nstencil = ctx->nstencil;
  if (ctx->numneigh.p && ctx->nlocal > 0) {
    CK(cudaMemsetAsync(ctx->cnt64, 0, sizeof(unsigned long long), ctx->stream));
    k_sum_int<<<std::min(cdiv(ctx->nlocal, 256), 1184), 256, 0, ctx->stream>>>(ctx->nlocal, ctx->numneigh.p,
                                                                             ctx->cnt64);
    ctx->launches++;
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, ctx->cnt64, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->npairs = (int64_t)h;
  }
  out->launches = ctx->launches;
  out->device_bytes = (double)ctx->dev_bytes;
  out->halo_transport = ctx->nranks > 1 ? (ctx->p2p ? 2 : 1) : 0;
  out->lanes_per_atom = ctx->tpa;
  out->list_kind = ctx->tiles_active ? 1 : 0;
  out->list_entries = out->npairs;
  if (ctx->tiles_active) {
    for (int d = 0; d < 3; d++) out->tile[d] = ctx->tg.t[d];
    out->tile_stage_max = ctx->tile_scap;
    out->tiles_interior = ctx->tile_nint;
    out->tiles_boundary = ctx->tile_nbnd;
    out->halo_overlap = ctx->overlap && ctx->pair_style == 1 && ctx->remote_mask && ctx->tile_nint > 0 &&
                        ctx->tile_nbnd > 0;
    // entries = sum of the list-row lengths (FWD members + transposed copies)
    CK(cudaMemsetAsync(ctx->cnt64, 0, sizeof(unsigned long long), ctx->stream));
    k_sum_u16<<<std::min(cdiv(ctx->tile_NI, 256), 1184), 256, 0, ctx->stream>>>(ctx->tile_NI, ctx->tl_num.p,
                                                                            ctx->cnt64);
    ctx->launches++;
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, ctx->cnt64, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->list_entries = (int64_t)h;
  }
  return B200_OK;
}

int b200_get_neighbor_list(b200_ctx *ctx, int *numneigh, int *neigh, int64_t cap, int64_t *npairs) {
  if (!ctx) return B200_EARG;
  const int nl = ctx->nlocal;
  std::vector<int> nn(nl);
  if (nl) CK(cudaMemcpy(nn.data(), ctx->numneigh.p, sizeof(int) * nl, cudaMemcpyDeviceToHost));
  std::vector<long long> first(nl + 1, 0);
  for (int i = 0; i < nl; i++) first[i + 1] = first[i] + nn[i];
  if (npairs) *npairs = first[nl];
  if (numneigh && nl) memcpy(numneigh, nn.data(), sizeof(int) * nl);
  if (!neigh || first[nl] == 0) return B200_OK;
  if (cap < first[nl]) return ctx->fail(B200_EARG, "neighbor buffer too small");
  long long *dfirst = nullptr;
  int *dflat = nullptr;
  CK(cudaMalloc((void **)&dfirst, sizeof(long long) * (nl + 1)));
  CK(cudaMalloc((void **)&dflat, sizeof(int) * first[nl]));
  CK(cudaMemcpy(dfirst, first.data(), sizeof(long long) * (nl + 1), cudaMemcpyHostToDevice));
  if (ctx->tiles_active) {
    const TileGeom &G = ctx->tg;
    const size_t sm = TILE_HDR_BYTES + (size_t)ctx->tile_scap * sizeof(int);
    k_tile_export<<<G.ntiles, 256, sm, ctx->stream>>>(G, nl, ctx->ostart.p, ctx->gstart.p, ctx->tile_ibase.p,
                                                     ctx->tile_NI, ctx->tile_slots, ctx->tl_num.p,
                                                     ctx->tl_list.p, dfirst, dflat, ctx->tile_scap,
                                                     ctx->eam2_active ? ctx->tl_far.p : nullptr,
                                                     ctx->newton ? 0 : 1);
  } else
    k_export_csr<<<cdiv(nl, 256), 256, 0, ctx->stream>>>(nl, ctx->nstride, ctx->tpa, ctx->numneigh.p,
                                                         ctx->neigh.p, dfirst, dflat);
  ctx->launches++;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(neigh, dflat, sizeof(int) * first[nl], cudaMemcpyDeviceToHost));
  cudaFree(dfirst);
  cudaFree(dflat);
  return B200_OK;
}

int b200_get_eam_rho_fp(b200_ctx *ctx, int with_ghosts, double *rho, double *fp) {
  if (!ctx) return B200_EARG;
  const int n = ctx->nlocal + (with_ghosts ? ctx->nghost : 0);
  CK(cudaStreamSynchronize(ctx->stream));
  if (rho && n) CK(cudaMemcpy(rho, ctx->rho, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (fp && n) CK(cudaMemcpy(fp, ctx->fp, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return B200_OK;
}

// Per-atom energy and virial of the pair style for the positions, list and (eam) fp in force:
// Pair::ev_tally's eatom / vatom (pair.cpp:1087-1182).  Call after a step or setup that tallied
// (eflag/vflag set: then final_integrate has run and nothing is pending).  Collective over the
// sub-domains when the list is a half list (the ghost shares travel back by the reverse halo).
int b200_pair_peratom(b200_ctx *ctx, double *eatom, double *vatom) {
  if (!ctx) return B200_EARG;
  if (!ctx->setup_done) return ctx->fail(B200_EARG, "b200_pair_peratom before b200_setup");
  TRY(flush_vops(ctx));
  if (ctx->ahead || ctx->pending_final)
    return ctx->fail(B200_EARG, "per-atom tallies need a step that tallied (eflag/vflag) just before");
  if (ctx->tiles_active && !ctx->full_ghost)
    return ctx->fail(B200_EARG, "per-atom tallies are not available on the first-generation eam tile list");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int nl = ctx->nlocal, ng = ctx->nghost, c = ctx->cur;
  const bool flat = !ctx->tiles_active;
  const size_t na = (size_t)std::max(flat ? nl + ng : nl, 1);
  TRY(reserve(ctx, ctx->peratom, 7 * na));
  PerAtomOut out;
  out.e = ctx->peratom.p;
  for (int k = 0; k < 6; k++) out.v[k] = ctx->peratom.p + (size_t)(1 + k) * na;
  if (flat) CK(cudaMemsetAsync(ctx->peratom.p, 0, sizeof(double) * 7 * na, s));
  const int st = ctx->pair_style;
  if (st != 1 && st != 2) return ctx->fail(B200_EARG, "no pair style set");
  if (nl > 0) {
    if (flat) {
      if (st == 1)
        k_peratom_flat<1><<<cdiv(nl, 128), 128, 0, s>>>(nl, ctx->nstride, ctx->tpa, ctx->xt[c], nullptr,
                                                        ctx->numneigh.p, ctx->neigh.p, ctx->lj_tab.p, ctx->eam,
                                                        ctx->ntypes, out);
      else
        k_peratom_flat<2><<<cdiv(nl, 128), 128, 0, s>>>(nl, ctx->nstride, ctx->tpa, ctx->xt[c], ctx->fp,
                                                        ctx->numneigh.p, ctx->neigh.p, nullptr, ctx->eam,
                                                        ctx->ntypes, out);
    } else {
      const TileGeom &G = ctx->tg;
      const size_t sm = tile_smem_bytes(ctx->tile_scap, G.srow_y * G.srow_z, G.sbx, false, true);
      const int thr = std::max(32, std::min(352, cdiv(std::max(ctx->tile_maxown, 1), 32) * 32));
      const unsigned short *tf = ctx->eam2_active ? ctx->tl_far.p : nullptr;
      if (st == 1)
        k_peratom_tile<1><<<G.ntiles, thr, sm, s>>>(G, nl, ctx->xt[c], nullptr, ctx->ostart.p, ctx->gstart.p,
                                                    ctx->tile_ibase.p, ctx->tile_NI, ctx->tile_slots,
                                                    ctx->tl_iloc.p, ctx->tl_num.p, tf, ctx->tl_list.p,
                                                    ctx->lj_tab.p, ctx->eam, ctx->ntypes, out, ctx->tile_scap);
      else
        k_peratom_tile<2><<<G.ntiles, thr, sm, s>>>(G, nl, ctx->xt[c], ctx->fp, ctx->ostart.p, ctx->gstart.p,
                                                    ctx->tile_ibase.p, ctx->tile_NI, ctx->tile_slots,
                                                    ctx->tl_iloc.p, ctx->tl_num.p, tf, ctx->tl_list.p,
                                                    nullptr, ctx->eam, ctx->ntypes, out, ctx->tile_scap);
    }
    ctx->launches++;
    LAUNCH_CHECK();
  }
  if (flat)  // every sub-domain walks the same halos, even one without atoms
    for (int k = 0; k < 7; k++) {
      Vec3Ptr a{{ctx->peratom.p + (size_t)k * na, nullptr, nullptr}};
      TRY(reverse_halo<1>(ctx, a));
    }
  if (st == 2 && nl > 0) {
    k_peratom_embed<<<cdiv(nl, 256), 256, 0, s>>>(nl, ctx->xt[c], ctx->eam, ctx->rho, out.e);
    ctx->launches++;
    LAUNCH_CHECK();
  }
  CK(cudaStreamSynchronize(s));
  if (nl > 0) {
    if (eatom) CK(cudaMemcpy(eatom, out.e, sizeof(double) * nl, cudaMemcpyDeviceToHost));
    if (vatom) {
      std::vector<double> h((size_t)6 * nl);
      for (int k = 0; k < 6; k++)
        CK(cudaMemcpy(h.data() + (size_t)k * nl, out.v[k], sizeof(double) * nl, cudaMemcpyDeviceToHost));
      for (int i = 0; i < nl; i++)
        for (int k = 0; k < 6; k++) vatom[(size_t)i * 6 + k] = h[(size_t)k * nl + i];
    }
  }
  return B200_OK;
}

int b200_set_profiling(b200_ctx *ctx, int on) {
  if (!ctx) return B200_EARG;
  ph_collect(ctx);
  ctx->profiling = on != 0;
  return B200_OK;
}

int b200_get_phase_times(b200_ctx *ctx, double ms[B200_NPHASE], int64_t calls[B200_NPHASE]) {
  if (!ctx) return B200_EARG;
  ph_collect(ctx);
  for (int k = 0; k < B200_NPHASE; k++) {
    if (ms) ms[k] = ctx->ph_ms[k];
    if (calls) calls[k] = ctx->ph_calls[k];
    ctx->ph_ms[k] = 0;
    ctx->ph_calls[k] = 0;
  }
  return B200_OK;
}

int b200_neighbor_ranks(const int procgrid[3], const int myloc[3], const int periodicity[3],
                        int nbr[27]) {
  if (!procgrid || !myloc || !periodicity || !nbr) return B200_EARG;
  for (int dir = 0; dir < NDIR; dir++) {
    const int dv[3] = {dir % 3 - 1, (dir / 3) % 3 - 1, dir / 9 - 1};
    int loc[3];
    bool exists = true;
    for (int d = 0; d < 3; d++) {
      loc[d] = myloc[d] + dv[d];
      if (loc[d] < 0 || loc[d] >= procgrid[d]) {
        if (periodicity[d]) loc[d] = (loc[d] + procgrid[d]) % procgrid[d];
        else exists = false;
      }
    }
    nbr[dir] = exists ? (loc[0] * procgrid[1] + loc[1]) * procgrid[2] + loc[2] : -1;
  }
  return B200_OK;
}

int b200_comm_unique_id(void *id128) {
  if (!id128) return B200_EARG;
  std::string why;
  if (!load_nccl(why)) return B200_ECUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return B200_ECUDA;
  memcpy(id128, &id, sizeof id);
  return B200_OK;
}

int b200_comm_init(b200_ctx *ctx, int nranks, int rank, const void *id128) {
  if (!ctx) return B200_EARG;
  if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128))
    return ctx->fail(B200_EARG, "b200_comm_init: bad arguments");
  if (ctx->nccl) return ctx->fail(B200_EARG, "b200_comm_init called twice");
  ctx->nranks = nranks;
  ctx->rank = rank;
  ctx->geom_ready = false;
  if (nranks == 1) return B200_OK;
  CK(cudaSetDevice(ctx->device));
  std::string why;
  if (!load_nccl(why)) return ctx->fail(B200_ECUDA, "%s", why.c_str());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NK(g_nccl.CommInitRank(&ctx->nccl, nranks, id, rank));
  cudaFree(ctx->allcounts);
  cudaFreeHost(ctx->h_counts);
  ctx->allcounts = nullptr;
  ctx->h_counts = nullptr;
  TRY(dalloc(ctx, &ctx->allcounts, (size_t)32 * nranks));
  CK(cudaMallocHost((void **)&ctx->h_counts, sizeof(int) * 32 * nranks));
  CK(cudaMalloc((void **)&ctx->ag_dev, (size_t)AG_BYTES * (nranks + 1)));
  CK(cudaMallocHost((void **)&ctx->ag_host, (size_t)AG_BYTES * (nranks + 1)));
  TRY(p2p_init(ctx));
  TRY(preload_rebuild_kernels(ctx));
  return B200_OK;
}


// =====================================================================================
//                 in-process group: N sub-domains driven by one host process
// =====================================================================================
static int group_fail(b200_group *g, int code, const std::string &msg) {
  g->err = msg;
  return code;
}

// run fn(i) on the worker thread of every context and wait; returns the first failure
static int group_run(b200_group *g, std::function<int(int)> fn) {
  {
    std::unique_lock<std::mutex> lk(g->jmu);
    g->job = std::move(fn);
    g->job_finished = 0;
    g->job_seq++;
  }
  g->jcv.notify_all();
  std::unique_lock<std::mutex> lk(g->jmu);
  g->dcv.wait(lk, [&] { return g->job_finished == g->n; });
  for (int i = 0; i < g->n; i++)
    if (g->rc[i] != B200_OK) {
      g->err = "sub-domain " + std::to_string(i) + ": " + g->ctx[i]->err;
      return g->rc[i];
    }
  return B200_OK;
}

static void group_worker(b200_group *g, int i) {
  cudaSetDevice(g->ctx[i]->device);
  unsigned long long seen = 0;
  for (;;) {
    std::function<int(int)> fn;
    {
      std::unique_lock<std::mutex> lk(g->jmu);
      g->jcv.wait(lk, [&] { return g->quit || g->job_seq != seen; });
      if (g->quit) return;
      seen = g->job_seq;
      fn = g->job;
    }
    const int rc = fn(i);
    {
      std::unique_lock<std::mutex> lk(g->jmu);
      g->rc[i] = rc;
      g->job_finished++;
    }
    g->dcv.notify_all();
  }
}

// what b200_comm_init does for processes, for the contexts of a group
static int group_comm_init(b200_ctx *ctx, b200_group *g, int rank) {
  ctx->grp = g;
  ctx->nranks = g->n;
  ctx->rank = rank;
  ctx->geom_ready = false;
  CK(cudaSetDevice(ctx->device));
  cudaFreeHost(ctx->h_counts);
  ctx->h_counts = nullptr;
  CK(cudaMallocHost((void **)&ctx->h_counts, sizeof(int) * 32 * (g->n + 1)));
  const size_t fbytes = 2u << 20;
  CK(cudaMalloc((void **)&ctx->pflags, fbytes));
  CK(cudaMemset(ctx->pflags, 0, fbytes));
  CK(cudaMalloc((void **)&ctx->p2p_counter, 8 * sizeof(unsigned)));
  CK(cudaMemset(ctx->p2p_counter, 0, 8 * sizeof(unsigned)));
  CK(cudaEventCreateWithFlags(&g->pub[rank].ready, cudaEventDisableTiming));
  ctx->peer_flags.assign(g->n, nullptr);
  ctx->peer_arena.assign(g->n, nullptr);
  ctx->peer_base.assign(g->n, nullptr);
  ctx->peer_roff.assign(g->n, 0);
  ctx->peer_gen.assign(g->n, 0);
  ctx->peer_off.assign((size_t)g->n * 2 * (NDIR + 1), 0);
  ctx->peer_cap_s.assign(g->n, 0);
  ctx->peer_cap_r.assign(g->n, 0);
  ctx->p2p = true;
  TRY(preload_rebuild_kernels(ctx));
  return B200_OK;
}

int b200_group_create(b200_group **out, int nsub, const int *devices, int precision) {
  if (!out || nsub < 1 || nsub > 64) return B200_EARG;
  *out = nullptr;
  b200_group *g = new b200_group();
  g->n = nsub;
  g->ctx.assign(nsub, nullptr);
  g->rc.assign(nsub, B200_OK);
  g->counts.assign((size_t)nsub * 32, 0);
  g->red.assign((size_t)nsub * 8, 0.0);
  g->redi.assign(nsub, 0);
  g->pub.resize(nsub);
  *out = g;
  for (int i = 0; i < nsub; i++) {
    const int rc = b200_create(&g->ctx[i], devices ? devices[i] : 0, precision);
    if (rc != B200_OK) return group_fail(g, rc, g->ctx[i] ? g->ctx[i]->err : "cannot create a device context");
  }
  // peer access between the distinct devices of the group (the per-step halo stores straight
  // into the neighbour's staging buffer)
  for (int i = 0; i < nsub; i++)
    for (int j = 0; j < nsub; j++) {
      const int a = g->ctx[i]->device, b = g->ctx[j]->device;
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, a, b);
      if (!can) return group_fail(g, B200_ECUDA, "devices of the group cannot access each other's memory");
      cudaSetDevice(a);
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return group_fail(g, B200_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  for (int i = 0; i < nsub; i++)
    for (int j = 0; j < i; j++) g->shared_dev |= g->ctx[i]->device == g->ctx[j]->device;
  for (int i = 0; i < nsub; i++) {
    const int rc = group_comm_init(g->ctx[i], g, i);
    if (rc != B200_OK) return group_fail(g, rc, g->ctx[i]->err);
  }
  for (int i = 0; i < nsub; i++)
    for (int r = 0; r < nsub; r++) g->ctx[i]->peer_flags[r] = g->ctx[r]->pflags;
  for (int i = 0; i < nsub; i++) g->workers.emplace_back(group_worker, g, i);
  return B200_OK;
}

void b200_group_destroy(b200_group *g) {
  if (!g) return;
  {
    std::unique_lock<std::mutex> lk(g->jmu);
    g->quit = true;
  }
  g->jcv.notify_all();
  for (auto &t : g->workers) t.join();
  for (int i = 0; i < g->n; i++) {
    if (g->pub[i].ready) cudaEventDestroy(g->pub[i].ready);
    if (g->ctx[i]) {
      g->ctx[i]->grp = nullptr;
      b200_destroy(g->ctx[i]);
    }
  }
  delete g;
}

const char *b200_group_last_error(const b200_group *g) { return g ? g->err.c_str() : "null group"; }
int b200_group_size(const b200_group *g) { return g ? g->n : 0; }
b200_ctx *b200_group_context(b200_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }

// brick grid for n sub-domains of a box with edge lengths prd: the factorisation with the
// smallest sub-domain surface (ProcMap::onelevel_grid / best_factors, procmap.cpp:48,690-750)
int b200_group_auto_grid(int n, const double prd[3], int grid[3]) {
  if (n < 1 || !prd || !grid) return B200_EARG;
  double best = 1.0e300;
  for (int px = 1; px <= n; px++) {
    if (n % px) continue;
    for (int py = 1; py <= n / px; py++) {
      if ((n / px) % py) continue;
      const int pz = n / px / py;
      const double lx = prd[0] / px, ly = prd[1] / py, lz = prd[2] / pz;
      const double surf = lx * ly + ly * lz + lx * lz;
      if (surf < best) {
        best = surf;
        grid[0] = px; grid[1] = py; grid[2] = pz;
      }
    }
  }
  return B200_OK;
}

int b200_group_set_grid(b200_group *g, const int grid[3]) {
  if (!g || !grid || grid[0] * grid[1] * grid[2] != g->n)
    return g ? group_fail(g, B200_EARG, "sub-domain grid does not match the number of sub-domains") : B200_EARG;
  for (int d = 0; d < 3; d++) g->grid[d] = grid[d];
  for (int i = 0; i < g->n; i++) {
    const int loc[3] = {i / (grid[1] * grid[2]), (i / grid[2]) % grid[1], i % grid[2]};
    const int rc = b200_set_decomposition(g->ctx[i], grid, loc);
    if (rc != B200_OK) return group_fail(g, rc, g->ctx[i]->err);
  }
  return B200_OK;
}

// all owned atoms of the whole box (already wrapped into it, Domain::pbc): split by owning
// sub-domain [sublo, subhi) with the bounds b200_set_decomposition gives every context
int b200_group_set_atoms(b200_group *g, int n, int ntypes, const double *mass, const double *x,
                         const double *v, const int *type, const int *tag, const int *mask,
                         const int *image) {
  if (!g || n < 0 || !mass || (n && (!x || !v || !type || !tag))) return B200_EARG;
  b200_ctx *c0 = g->ctx[0];
  if (!c0->have_box) return group_fail(g, B200_EARG, "b200_set_box before b200_group_set_atoms");
  std::vector<std::vector<int>> idx(g->n);
  const int *P = g->grid;
  // triclinic: sub-domains are bricks in lamda coordinates (Domain::x2lamda, domain.cpp:2376-2390)
  const bool tri = c0->tri;
  double hinv[6] = {0, 0, 0, 0, 0, 0};
  if (tri) {
    const double h0 = c0->prd[0], h1 = c0->prd[1], h2 = c0->prd[2];
    hinv[0] = 1.0 / h0; hinv[1] = 1.0 / h1; hinv[2] = 1.0 / h2;
    hinv[3] = -c0->yz / (h1 * h2);
    hinv[4] = (c0->yz * c0->xy - h1 * c0->xz) / (h0 * h1 * h2);
    hinv[5] = -c0->xy / (h0 * h1);
  }
  for (int i = 0; i < n; i++) {
    int loc[3];
    double lam[3] = {0, 0, 0};
    if (tri) {
      const double d0 = x[3 * (size_t)i] - c0->boxlo[0], d1 = x[3 * (size_t)i + 1] - c0->boxlo[1],
                   d2 = x[3 * (size_t)i + 2] - c0->boxlo[2];
      lam[0] = hinv[0] * d0 + hinv[5] * d1 + hinv[4] * d2;
      lam[1] = hinv[1] * d1 + hinv[3] * d2;
      lam[2] = hinv[2] * d2;
    }
    for (int d = 0; d < 3; d++) {
      const double c = tri ? lam[d] : x[3 * (size_t)i + d];
      const double blo = tri ? 0.0 : c0->boxlo[d], bprd = tri ? 1.0 : c0->prd[d];
      int l = (int)((c - blo) / bprd * P[d]);
      l = std::min(std::max(l, 0), P[d] - 1);
      // the exact bounds of setup_geometry (boxlo + prd * l / P, last one closed at boxhi)
      auto lo = [&](int q) { return tri ? q * 1.0 / P[d] : c0->boxlo[d] + c0->prd[d] * (q * 1.0 / P[d]); };
      while (l > 0 && c < lo(l)) l--;
      while (l < P[d] - 1 && c >= lo(l + 1)) l++;
      loc[d] = l;
    }
    idx[(loc[0] * P[1] + loc[1]) * P[2] + loc[2]].push_back(i);
  }
  for (int r = 0; r < g->n; r++) {
    const std::vector<int> &id = idx[r];
    const size_t m = id.size();
    std::vector<double> xs(3 * m), vs(3 * m);
    std::vector<int> ts(m), gs(m), ms(mask ? m : 0), is(image ? m : 0);
    for (size_t k = 0; k < m; k++) {
      const size_t i = id[k];
      for (int d = 0; d < 3; d++) { xs[3 * k + d] = x[3 * i + d]; vs[3 * k + d] = v[3 * i + d]; }
      ts[k] = type[i]; gs[k] = tag[i];
      if (mask) ms[k] = mask[i];
      if (image) is[k] = image[i];
    }
    const int rc = b200_set_atoms(g->ctx[r], (int)m, ntypes, mass, xs.data(), vs.data(), ts.data(), gs.data(),
                                  mask ? ms.data() : nullptr, image ? is.data() : nullptr);
    if (rc != B200_OK) return group_fail(g, rc, g->ctx[r]->err);
  }
  return B200_OK;
}

int b200_group_count(b200_group *g, int *nlocal_total, int *nghost_total) {
  if (!g) return B200_EARG;
  int a = 0, b = 0;
  for (int i = 0; i < g->n; i++) { a += g->ctx[i]->nlocal; b += g->ctx[i]->nghost; }
  if (nlocal_total) *nlocal_total = a;
  if (nghost_total) *nghost_total = b;
  return B200_OK;
}

// owned atoms of all sub-domains, concatenated in sub-domain order
int b200_group_get_atoms(b200_group *g, double *x, double *v, double *f, int *type, int *tag, int *mask,
                         int *image) {
  if (!g) return B200_EARG;
  std::vector<size_t> off(g->n + 1, 0);
  for (int i = 0; i < g->n; i++) off[i + 1] = off[i] + g->ctx[i]->nlocal;
  return group_run(g, [&](int i) {
    const size_t o = off[i];
    return b200_get_atoms(g->ctx[i], 0, x ? x + 3 * o : nullptr, v ? v + 3 * o : nullptr, f ? f + 3 * o : nullptr,
                          type ? type + o : nullptr, tag ? tag + o : nullptr, mask ? mask + o : nullptr,
                          image ? image + o : nullptr);
  });
}

// ---- the stages of a timestep, one by one, for a host whose integrator is its own fix
// (fix nvt/b200 under `package b200 gpus N`): every sub-domain runs the stage, in step
#define GROUP_STAGE(NAME, ARGS, CALL)                                        \
  int b200_group_##NAME ARGS {                                               \
    if (!g) return B200_EARG;                                                \
    return group_run(g, [&](int i) { b200_ctx *c = g->ctx[i]; return CALL; }); \
  }
GROUP_STAGE(nve_v, (b200_group *g, double dtf, int groupbit), b200_nve_v(c, dtf, groupbit))
GROUP_STAGE(nve_x, (b200_group *g, double dtv, int groupbit), b200_nve_x(c, dtv, groupbit))
GROUP_STAGE(scale_v, (b200_group *g, double factor, int groupbit), b200_scale_v(c, factor, groupbit))
GROUP_STAGE(scale_v3, (b200_group *g, const double factor[3], int groupbit), b200_scale_v3(c, factor, groupbit))
GROUP_STAGE(remap, (b200_group *g, const double oldlo[3], const double oldhi[3], const double newlo[3],
                    const double newhi[3], int groupbit), b200_remap(c, oldlo, oldhi, newlo, newhi, groupbit))
GROUP_STAGE(add_force, (b200_group *g, const double df[3], int groupbit), b200_add_force(c, df, groupbit))
GROUP_STAGE(reneighbor, (b200_group *g), b200_reneighbor(c))
GROUP_STAGE(forward_comm, (b200_group *g), b200_forward_comm(c))
GROUP_STAGE(force_clear, (b200_group *g), b200_force_clear(c))
GROUP_STAGE(pair_compute, (b200_group *g, int eflag, int vflag), b200_pair_compute(c, eflag, vflag))
GROUP_STAGE(reverse_comm, (b200_group *g), b200_reverse_comm(c))
#undef GROUP_STAGE

// fix langevin over all sub-domains; fsum (nullable) = the random force summed over the whole group
int b200_group_langevin(b200_group *g, int ntypes, const double *gfactor1, const double *gfactor2_tsqrt,
                        int groupbit, uint64_t seed, int64_t step, const double *uniforms_by_tag,
                        int64_t nuniform, double *fsum) {
  if (!g) return B200_EARG;
  std::vector<double> part(3 * (size_t)g->n, 0.0);
  int rc = group_run(g, [&](int i) {
    return b200_langevin(g->ctx[i], ntypes, gfactor1, gfactor2_tsqrt, groupbit, seed, step, uniforms_by_tag,
                         nuniform, fsum ? &part[3 * (size_t)i] : nullptr);
  });
  if (rc != B200_OK) return rc;
  if (fsum)
    for (int d = 0; d < 3; d++) {
      fsum[d] = 0.0;
      for (int i = 0; i < g->n; i++) fsum[d] += part[3 * (size_t)i + d];
    }
  return B200_OK;
}

int b200_group_decide(b200_group *g, int *rebuild) {
  if (!g || !rebuild) return B200_EARG;
  std::vector<int> rb(g->n, 0);
  const int rc = group_run(g, [&](int i) { return b200_decide(g->ctx[i], &rb[i]); });
  *rebuild = rb[0];  // (the vote is the group's: every member holds the same answer)
  return rc;
}

int b200_group_pair_peratom(b200_group *g, double *eatom, double *vatom) {
  if (!g) return B200_EARG;
  std::vector<size_t> off(g->n + 1, 0);
  for (int i = 0; i < g->n; i++) off[i + 1] = off[i] + g->ctx[i]->nlocal;
  return group_run(g, [&](int i) {
    return b200_pair_peratom(g->ctx[i], eatom ? eatom + off[i] : nullptr, vatom ? vatom + 6 * off[i] : nullptr);
  });
}

int b200_group_setup(b200_group *g, int eflag, int vflag) {
  if (!g) return B200_EARG;
  return group_run(g, [&](int i) { return b200_setup(g->ctx[i], eflag, vflag); });
}

int b200_group_step(b200_group *g, int eflag, int vflag, int *rebuilt) {
  return b200_group_step_ahead(g, eflag, vflag, 0, rebuilt);
}

int b200_group_step_ahead(b200_group *g, int eflag, int vflag, int more, int *rebuilt) {
  if (!g) return B200_EARG;
  std::vector<int> rb(g->n, 0);
  const int rc = group_run(g, [&](int i) { return b200_step_ahead(g->ctx[i], eflag, vflag, more, &rb[i]); });
  if (rebuilt) *rebuilt = rb[0];
  return rc;
}

int b200_group_run(b200_group *g, int nsteps, int64_t first_step, int thermo_every, double *thermo_out,
                   int max_thermo, int *n_thermo) {
  if (!g) return B200_EARG;
  std::vector<int> nt(g->n, 0);
  std::vector<std::vector<double>> scratch(g->n);
  for (int i = 1; i < g->n; i++) scratch[i].assign((size_t)10 * std::max(max_thermo, 1), 0.0);
  const int rc = group_run(g, [&](int i) {
    return b200_run(g->ctx[i], nsteps, first_step, thermo_every, i == 0 ? thermo_out : scratch[i].data(),
                    max_thermo, &nt[i]);
  });
  if (n_thermo) *n_thermo = nt[0];
  return rc;
}

int b200_group_get_tallies(b200_group *g, double *eng_vdwl, double virial[6]) {
  return g ? b200_get_tallies(g->ctx[0], eng_vdwl, virial) : B200_EARG;  // group sums on every context
}

int b200_group_ke_sum(b200_group *g, double *mv2) {
  if (!g || !mv2) return B200_EARG;
  std::vector<double> r(g->n, 0.0);
  const int rc = group_run(g, [&](int i) { return b200_ke_sum(g->ctx[i], &r[i]); });
  *mv2 = r[0];
  return rc;
}

int b200_group_ke_group(b200_group *g, int groupbit, double *mv2, double tensor[6]) {
  if (!g) return B200_EARG;
  std::vector<double> r((size_t)g->n * 7, 0.0);
  const int rc = group_run(g, [&](int i) { return b200_ke_group(g->ctx[i], groupbit, &r[7 * (size_t)i], &r[7 * (size_t)i + 1]); });
  if (mv2) *mv2 = r[0];
  if (tensor) memcpy(tensor, &r[1], 6 * sizeof(double));
  return rc;
}

int b200_group_last_run_ms(b200_group *g, double *ms) {
  if (!g || !ms) return B200_EARG;
  double m = 0.0;
  for (int i = 0; i < g->n; i++) m = std::max(m, g->ctx[i]->last_run_ms);
  *ms = m;
  return B200_OK;
}

// counters summed over the sub-domains (builds, ago, bins, ... are those of sub-domain 0)
int b200_group_get_stats(b200_group *g, b200_stats *out) {
  if (!g || !out) return B200_EARG;
  b200_stats t;
  for (int i = 0; i < g->n; i++) {
    const int rc = b200_get_stats(g->ctx[i], i == 0 ? out : &t);
    if (rc != B200_OK) return group_fail(g, rc, g->ctx[i]->err);
    if (i > 0) {
      out->npairs += t.npairs;
      out->list_entries += t.list_entries;
      out->launches += t.launches;
      out->device_bytes += t.device_bytes;
      out->max_numneigh = std::max(out->max_numneigh, t.max_numneigh);
      out->tiles_interior += t.tiles_interior;
      out->tiles_boundary += t.tiles_boundary;
    }
  }
  return B200_OK;
}

}  // extern "C"
