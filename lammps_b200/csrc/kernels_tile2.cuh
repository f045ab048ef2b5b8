// kernels_tile2.cuh -- second-generation FP64 lj/cut kernel over the bin-tile list.
//
// Same list and staged-tile idea as kernels_tile.cuh (one CTA per tile of bins, the tile's
// neighbourhood staged in shared memory, 16-bit entries, every pair evaluated from both sides so
// that f_i is STORED: no atomics, no force clear, and -- with the FULLGHOST list -- no Newton
// scatter onto ghosts and no reverse halo).  What changed against k_tile_lj<EV,ONETYPE,false>,
// each item from an ncu capture or an ablation run on 4 M atoms (profiles/r02_*):
//   * ghost scatter gone (FULLGHOST list, k_tile_build): the per-step search for FWD|GHOST
//     entries and their RED.F64 path cost 22 % of the kernel on a single GPU, where 30 % of the
//     tiles touch the periodic boundary.
//   * FP64 instruction diet, 22 -> 18 per entry.  rsq is formed with two FMAs (3 instead of 5
//     instructions); it can differ from the reference's separately rounded sum by a few ulp, so
//     an entry whose rsq shares its upper 32 bits (+-1) with the cutoff -- about 2e-4 entries per
//     atom and step -- is left out of the fast path and re-decided with the reference's
//     operation order (rsq_ref): the set of interacting pairs is still exactly the CPU path's.
//     The reciprocal is MUFU.RCP64H (20 bits) + one cubic Newton step (error e^3 = 2^-60, below
//     1 ulp; 3 FMAs instead of 5).  The pair function is r2inv^4 * (lj1*r6inv - lj2): the same
//     five multiplies, one level less of dependency.
//   * no per-entry "valid" predicate: the build pads the last word of a row with the index of a
//     dummy staged atom parked far away (rsq = 3e20, never inside a cutoff).
//   * positions are staged as one 24-byte record per atom: one address per entry, immediate
//     offsets for y and z; a 24-byte stride spreads over the 16 bank pairs like SoA does.
//   * software pipeline: the positions of the next step's entries are loaded (volatile
//     ld.shared) while the current step computes, and list words are fetched two words ahead
//     -- the first use of a list word was 28 % of all stall samples with one word of lookahead.
//   * the row header (iloc, tnum, global index) and the first two list words of a chunk are
//     requested together: one DRAM round trip per chunk instead of two.
//   * warps take 32-atom chunks of the tile from a shared counter instead of fixed strides.
//   * shared memory per staged atom 24 B (32 B before: no staged->global map, no type array for
//     a single type).
// Launch shape (threads per CTA, CTAs per SM) and the number of pair bodies interleaved per
// thread are template parameters; engine.cu picks one (B200_LJ2=threads,minb,ilp overrides).
// Measured FP peaks of this GPU (tools/microbench/fp_peak.cu): DFMA 16.8 T/s = 33.7 TFLOP/s,
// latency 8.2 cycles, 2 cycles per warp instruction and SM sub-partition; MUFU.RCP64H 17.7
// cycles, 8 cycles per warp; dependent LDS.64 39.6 cycles.
// No tensor cores: nothing here is a dense contraction.
#pragma once
#include "kernels_tile.cuh"

__host__ __device__ __forceinline__ size_t tile2_smem_bytes(int scap, bool with_type) {
  size_t b = TILE_HDR_BYTES + (size_t)scap * 3 * sizeof(double);
  if (with_type) b += (size_t)scap * sizeof(int);
  return (b + 127) / 128 * 128;
}

// reciprocal of a normal positive double, <= ~1 ulp: MUFU seed + one cubic Newton step
__device__ __forceinline__ double rcp_cubic(double a) {
  double x;
#ifdef TILE2_X_NOMUFU  // timing experiment only (wrong arithmetic)
  x = 0.2;
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
#endif
  double e = fma(-a, x, 1.0);
  e = fma(e, e, e);
  return fma(x, e, x);
}

#define TILE2_FAR 1.0e10

// Time integration fused into the pair kernel's epilogue (NVE = true).  With the FULLGHOST list a
// thread holds the COMPLETE force on its atom when its row is done, so FixNVE::final_integrate
// of this step and initial_integrate of the next (fix_nve.cpp:68-145) can be applied at once,
// with the arithmetic of k_nve_final_initial (each operation rounded separately, like the two
// reference loops): v += dtfm f; v += dtfm f; x += dtv v.  The new positions go to the OTHER
// position buffer (all CTAs still stage from the current one); the engine swaps the two after
// the launch.  Forces are not stored at all on such a step, and the separate integrate kernel
// (164 B/atom of HBM traffic at 6.6 TB/s = 9 % of a 32 M-atom step) disappears.
struct NveFuse {
  double4 *xt_out;
  double *vx, *vy, *vz;
  const int *mask;
  const double *mass;
  double dtv, dtf;
  int groupbit, do_check;
  const double *xhx, *xhy, *xhz;
  double triggersq;
  int *moved;
};

template <bool EV, bool ONETYPE, int ILP, int MAXT, int MINB, bool NVE = false>
__global__ void __launch_bounds__(MAXT, MINB) k_tile_lj2(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const int *__restrict__ tgi, const uint4 *__restrict__ list, double *__restrict__ fx,
    double *__restrict__ fy, double *__restrict__ fz, LJOne one, const double *__restrict__ tab,
    int ntypes, double *__restrict__ ev, int scap, int *__restrict__ tflags,
    const int *__restrict__ tile_ids, NveFuse nv, const unsigned char *__restrict__ hdrs) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  // staged positions: {x,y} as one 16-byte pair per atom, z in its own array -- an entry costs one
  // LDS.128 and one LDS.64 (16 wavefronts per warp on scattered partners; three LDS.64 cost 18.5)
  double *pos = reinterpret_cast<double *>(tsm + TILE_HDR_BYTES);
  double *posz = pos + (size_t)2 * scap;
  int *stype = reinterpret_cast<int *>(pos + (size_t)3 * scap);
  int *chunk_ctr = reinterpret_cast<int *>(&H->pad0);
  const unsigned pos_s = (unsigned)__cvta_generic_to_shared(pos);
  const unsigned posz_s = (unsigned)__cvta_generic_to_shared(posz);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows_cached(hdrs, tile, H);  // (the stored header carries a zero chunk counter)
  if (S + 1 > scap) {  // cannot happen between rebuilds (the rows are those the build staged)
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {  // stage: one warp per run of records, cp.async of {x,y}, z (and type) of each record
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
        if (!ONETYPE) cp_async4(stype + s, p + 3);
      }
    }
    if (tid < 2) pos[2 * S + tid] = TILE2_FAR;  // the dummy atom padding entries point at
    if (tid == 2) posz[S] = TILE2_FAR;
    if (!ONETYPE && tid == 3) stype[S] = 1;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  const int ni = H->ni, ibase = tile_ibase[tile];
  const int n1 = ntypes + 1, n2 = n1 * n1;
  const int chi1 = __double2hiint(one.cutsq) - 1;
  double evdwl = 0.0, vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(chunk_ctr, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk * 32 >= ni) break;
    const int ti = chunk * 32 + lane;
    if (ti >= ni) continue;
    const int g = ibase + ti;
    // row header and the first two list words: independent loads, one DRAM round trip
    // (every row owns maxslots >= 16 slots, so word 1 is always readable)
    const uint4 *lp = list + g;
    uint4 q0 = __ldg(lp), q1 = __ldg(lp + NI);
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    if (n == 0) q0 = make_uint4(S, S, S, S);  // nothing was written: name the dummy atom
    const int gi = tgi[g];
    const double pix = pos[2 * li], piy = pos[2 * li + 1], piz = posz[li];
    const int itype = ONETYPE ? 1 : stype[li];
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;
    // fused integrator: the atom's velocity and group mask are requested now (volatile asm keeps
    // the loads here) and used after the row, one DRAM latency earlier than the epilogue would
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int imask = 0;
    if (NVE) {
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v0) : "l"(nv.vx + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v1) : "l"(nv.vy + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v2) : "l"(nv.vz + gi));
      asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(imask) : "l"(nv.mask + gi));
    }

    // staged position of the atom an entry names.  The loads are volatile asm so that they stay
    // where the software pipeline below puts them: one step ahead of their use.
    struct P3 { double x, y, z; };
    auto ldpos = [&](unsigned e) -> P3 {
      P3 p;
      const unsigned j = e & TILE_IDX;
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.x), "=d"(p.y) : "r"(pos_s + j * 16u));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(p.z) : "r"(posz_s + j * 8u));
      return p;
    };
    // geometry of entry e: del, rsq (FMA form), pair-type index, cutoff
    auto geom = [&](unsigned e, const P3 &pj, double &dx, double &dy, double &dz, double &rsq,
                    int &tij, double &cutsq, int &chi) {
      dx = pix - pj.x;
      dy = piy - pj.y;
      dz = piz - pj.z;
      rsq = fma(dz, dz, fma(dy, dy, dx * dx));
      tij = 0;
      cutsq = one.cutsq;
      chi = chi1;
      if (!ONETYPE) {
        tij = itype * n1 + stype[e & TILE_IDX];
        cutsq = __ldg(tab + tij);
        chi = __double2hiint(cutsq) - 1;
      }
    };
    // pair function: fpair (times del = force on i) and, when EV, the pair energy
    auto ljf = [&](double rsq, int tij, double &fpair, double &epair) {
      const double lj1 = ONETYPE ? one.lj1 : __ldg(tab + n2 + tij);
      const double lj2 = ONETYPE ? one.lj2 : __ldg(tab + 2 * n2 + tij);
      const double r2inv = rcp_cubic(rsq);
      const double u = r2inv * r2inv;
      const double r6inv = u * r2inv;
      fpair = (u * u) * fma(lj1, r6inv, -lj2);
      if (EV) {
        const double lj3 = ONETYPE ? one.lj3 : __ldg(tab + 3 * n2 + tij);
        const double lj4 = ONETYPE ? one.lj4 : __ldg(tab + 4 * n2 + tij);
        const double off = ONETYPE ? one.offset : __ldg(tab + 5 * n2 + tij);
        epair = r6inv * fma(lj3, r6inv, -lj4) - off;
      }
    };
    // energy and virial of one pair: tallied once, on the half-list (FWD) copy
    // (Pair::ev_tally, pair.cpp:1087-1182: v = del (x) del * fpair)
    auto tally = [&](unsigned e, double dx, double dy, double dz, double f, double ep) {
      const bool fwd = (e & TILE_FWD) != 0;
      evdwl += fwd ? ep : 0.0;
      const double w = fwd ? f : 0.0;
      vir[0] = fma(dx * dx, w, vir[0]); vir[1] = fma(dy * dy, w, vir[1]);
      vir[2] = fma(dz * dz, w, vir[2]); vir[3] = fma(dx * dy, w, vir[3]);
      vir[4] = fma(dx * dz, w, vir[4]); vir[5] = fma(dy * dz, w, vir[5]);
    };
    // fast path of one entry (branch-free); returns true when the cutoff decision is too close
    // to call: the entry then contributes nothing here and `exact` decides it
    auto body = [&](unsigned e, const P3 &pj) -> bool {
      double dx, dy, dz, rsq, cutsq, fp, ep = 0.0;
      int tij, chi;
      geom(e, pj, dx, dy, dz, rsq, tij, cutsq, chi);
#ifdef TILE2_X_NOAMB
      const bool amb = false;
#else
      const bool amb = (unsigned)(__double2hiint(rsq) - chi) <= 2u;
#endif
      const bool in = rsq < cutsq && !amb;
      ljf(rsq, tij, fp, ep);
      const double f = in ? fp : 0.0;
      fxi = fma(dx, f, fxi);
      fyi = fma(dy, f, fyi);
      fzi = fma(dz, f, fzi);
      if (EV) tally(e, dx, dy, dz, f, in ? ep : 0.0);
      return amb;
    };
    // an entry the fast path left out: the reference's rsq decides, then the same function
    auto exact = [&](unsigned e) {
      double dx, dy, dz, rsq, cutsq, fp, ep = 0.0;
      int tij, chi;
      geom(e, ldpos(e), dx, dy, dz, rsq, tij, cutsq, chi);
      if ((unsigned)(__double2hiint(rsq) - chi) <= 2u && rsq_ref(dx, dy, dz) < cutsq) {
        ljf(rsq, tij, fp, ep);
        fxi = fma(dx, fp, fxi);
        fyi = fma(dy, fp, fyi);
        fzi = fma(dz, fp, fzi);
        if (EV) tally(e, dx, dy, dz, fp, ep);
      }
    };

    // The row, ILP entries per step.  cur[] holds the positions for the step about to run (loaded
    // during the previous step), nxt[] is filled while it runs.
    P3 cur[ILP], nxt[ILP];
    auto entry = [&](const uint4 &w, int k) -> unsigned {  // entry k (0..7) of a list word
      const unsigned v = k < 2 ? w.x : (k < 4 ? w.y : (k < 6 ? w.z : w.w));
      return (k & 1) ? (v >> 16) : (v & 0xffffu);
    };
#pragma unroll
    for (int k = 0; k < ILP; k++) cur[k] = ldpos(entry(q0, k));
    for (int k0 = 0; k0 < n; k0 += 8) {
      const uint4 c = q0;
      q0 = q1;
      if (k0 + 16 < n) q1 = __ldg(lp + (size_t)((k0 >> 3) + 2) * NI);
      bool again = false;
      // where the positions loaded during this word's last step come from: the next word, or
      // (after the last word, whose successor may never have been written) this word again
      const uint4 cn = (k0 + 8 < n) ? q0 : c;
#pragma unroll
      for (int st = 0; st < 8 / ILP; st++) {
#pragma unroll
        for (int k = 0; k < ILP; k++)
          nxt[k] = (st + 1 < 8 / ILP) ? ldpos(entry(c, (st + 1) * ILP + k)) : ldpos(entry(cn, k));
#pragma unroll
        for (int k = 0; k < ILP; k++) again |= body(entry(c, st * ILP + k), cur[k]);
#pragma unroll
        for (int k = 0; k < ILP; k++) cur[k] = nxt[k];
      }
      if (again) {
#pragma unroll 1
        for (int k = 0; k < 8; k++) {
          const unsigned v = k < 2 ? c.x : (k < 4 ? c.y : (k < 6 ? c.z : c.w));
          exact((k & 1) ? (v >> 16) : (v & 0xffffu));
        }
      }
    }
    if (NVE) {
      double px = pix, py = piy, pz = piz;
      if (imask & nv.groupbit) {
        const double dtfm = nv.dtf / nv.mass[itype];
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
      nv.xt_out[gi] = make_double4(px, py, pz, type2d(itype));
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
    block_sum<7>(v, pos);
    if (tid == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ev[k], v[k]);
  }
}


// global index of staged atom s (row tables; used on rare paths only)
__device__ __forceinline__ int tile_global_index(const TileHdr *H, int nlocal, int s) {
  int lo = 0, hi = H->nrows;  // rowbase[lo] <= s < rowbase[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (H->rowbase[mid] <= s) lo = mid; else hi = mid;
  }
  const int k = s - H->rowbase[lo], no = H->row_no[lo];
  return k < no ? H->row_o0[lo] + k : nlocal + H->row_g0[lo] + (k - no);
}

// ---------------------------------------------------------------------------------------
// lj/cut over a tile, mixed precision, second generation (replaces k_tile_lj_fx as the default
// of `prec mixed`).  FP32 pair math on fixed-point staged positions, FP64 accumulation:
//   * a staged atom is ONE 16-byte record {qx, qy, qz, type}: q = rn((x - origin) * fxscale) as
//     int32 (resolution ~1.5e-8 sigma, differences of two staged coordinates are exact): one
//     LDS.128 and one address per list entry (k_tile_lj_fx: four LDS.32 from four arrays), and
//     16 instead of 20 bytes of shared memory per staged atom;
//   * differences, rsq and the running force sums stay in fixed-point units (the scale is folded
//     into the reciprocal and applied to the sums once per list word);
//   * a cutoff decision within 2e-6 (relative) of the cutoff is re-taken in FP64 from the global
//     positions with the reference's operation order, so the set of interacting pairs is the CPU
//     path's; everything else of k_tile_lj2: no ghost scatter (FULLGHOST list), dummy-padded rows,
//     software-pipelined shared-memory and list loads, chunk scheduling, fused fix nve;
//   * FP32 partial sums of one list word (8 entries) are added to FP64 accumulators.
// Tolerances (BASELINE.json): forces <= 1e-5 relative, energy/pressure <= 1e-6.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t tile2f_smem_bytes(int scap) {
  return (TILE_HDR_BYTES + (size_t)scap * sizeof(int4) + 127) / 128 * 128;
}

// Sub-domain-wide fixed point (GQ = true): every atom's record {qx,qy,qz,type} relative to ONE
// origin (the corner of the local bin grid) lives in a global int4 array that the integrators
// keep current, so staging a tile is one 16-byte cp.async per atom -- no conversion, half the
// staging bytes of the double4 records.  30 bits per coordinate (the dummy atom of padding
// entries sits at -2^30, so differences cannot overflow); engine.cu uses it when the grid
// spacing extent / 2^30 is fine enough for the 2e-6 decision band (up to ~370 sigma per
// sub-domain edge), else the tile-relative conversion below (GQ = false).
struct QGeom {
  double ox, oy, oz, scale;
};
__device__ __forceinline__ int4 q_record(const double4 &p, const QGeom &Q) {
  return make_int4(__double2int_rn((p.x - Q.ox) * Q.scale), __double2int_rn((p.y - Q.oy) * Q.scale),
                   __double2int_rn((p.z - Q.oz) * Q.scale), d2type(p.w));
}
__global__ void __launch_bounds__(256) k_xt_to_q(int first, int n, const double4 *__restrict__ xt,
                                                 int4 *__restrict__ q, QGeom Q) {
  const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < first + n) q[i] = q_record(xt[i], Q);
}

template <bool EV, bool ONETYPE, int MAXT, int MINB, bool NVE = false, bool GQ = false>
__global__ void __launch_bounds__(MAXT, MINB) k_tile_lj2f(
    TileGeom G, int nlocal, const double4 *__restrict__ xt, const int *__restrict__ ostart,
    const int *__restrict__ gstart, const int *__restrict__ tile_ibase, int NI, int maxslots,
    const unsigned short *__restrict__ iloc, const unsigned short *__restrict__ tnum,
    const int *__restrict__ tgi, const uint4 *__restrict__ list, double *__restrict__ fx,
    double *__restrict__ fy, double *__restrict__ fz, LJOne one, LJOneF onef,
    const double *__restrict__ tab, const float *__restrict__ tabf, int ntypes,
    double *__restrict__ ev, int scap, int *__restrict__ tflags, const int *__restrict__ tile_ids,
    NveFuse nv, const unsigned char *__restrict__ hdrs, const int4 *__restrict__ qin,
    int4 *__restrict__ qout, QGeom Q) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  int4 *rec = reinterpret_cast<int4 *>(tsm + TILE_HDR_BYTES);
  int *chunk_ctr = reinterpret_cast<int *>(&H->pad0);
  const unsigned rec_s = (unsigned)__cvta_generic_to_shared(rec);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const TilePos P = tile_pos(G, tile);
  const int S = tile_rows_cached(hdrs, tile, H);  // (the stored header carries a zero chunk counter)
  if (S + 1 > scap) {
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  if (GQ) {  // stage: one warp per run, one 16-byte cp.async per atom from the global fixed-point records
    const int nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rec_s + (unsigned)(base + k) * 16u),
                     "l"(qin + src)
                     : "memory");
      }
    }
    if (tid == 0) rec[S] = make_int4(-(1 << 30), -(1 << 30), -(1 << 30), 1);  // the dummy atom
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  } else {  // stage: one warp per run of records, LDG.256 of the record, fixed-point conversion, STS.128
    const double ox = G.bin0[0] + (P.tx0 - G.s[0]) * G.bsize[0] - G.fxpad,
                 oy = G.bin0[1] + (P.ty0 - G.s[1]) * G.bsize[1] - G.fxpad,
                 oz = G.bin0[2] + (P.tz0 - G.s[2]) * G.bsize[2] - G.fxpad;
    const int nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k;
        const double4 p = ld_xt(xt + src);
        rec[base + k] = make_int4(__double2int_rn((p.x - ox) * G.fxscale), __double2int_rn((p.y - oy) * G.fxscale),
                                  __double2int_rn((p.z - oz) * G.fxscale), d2type(p.w));
      }
    }
    // the dummy atom of padding entries: a corner of the fixed-point range, > one cutoff (the
    // pad) away from every staged atom in each coordinate
    if (tid == 0) rec[S] = make_int4(0, 0, 0, 1);
    __syncthreads();
  }
  const int ni = H->ni, ibase = tile_ibase[tile];
  const int n1 = ntypes + 1, n2 = n1 * n1;
  const double dscale = GQ ? Q.scale : G.fxscale;
  const float scale = (float)dscale, s2 = scale * scale, inv = (float)(1.0 / dscale);
  // the single-type constants in fixed-point units: rsq_q = rsq * scale^2
  const float cutq1 = (float)(one.cutsq * dscale * dscale);
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
    uint4 q0 = __ldg(lp), q1 = __ldg(lp + NI);
    const int li = iloc[g];
    const int n = min((int)tnum[g], maxslots);
    if (n == 0) q0 = make_uint4(S, S, S, S);
    const int gi = tgi[g];
    const int4 me = rec[li];
    const int itype = ONETYPE ? 1 : me.w;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    int imask = 0;
    if (NVE) {
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v0) : "l"(nv.vx + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v1) : "l"(nv.vy + gi));
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v2) : "l"(nv.vz + gi));
      asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(imask) : "l"(nv.mask + gi));
    }
    double fxi = 0.0, fyi = 0.0, fzi = 0.0;          // FP64 accumulators (real units)
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;           // FP32 partial sums of one list word (q units)
    float ei = 0.0f, w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f, w4 = 0.0f, w5 = 0.0f;

    auto ldrec = [&](unsigned e) -> int4 {
      int4 r;
      const unsigned a = rec_s + (e & TILE_IDX) * 16u;
      asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
      return r;
    };
    // FP32 pair function in fixed-point units: rsq_q -> fpair (force on i = dq * inv * fpair)
    auto ljf = [&](float rsq_q, int tij, float &fpair, float &epair) {
      const float lj1 = ONETYPE ? onef.lj1 : __ldg(tabf + tij);
      const float lj2 = ONETYPE ? onef.lj2 : __ldg(tabf + n2 + tij);
      const float r2inv = rcp_f(rsq_q) * s2;
      const float u = r2inv * r2inv;
      const float r6inv = u * r2inv;
      fpair = (u * u) * fmaf(lj1, r6inv, -lj2);
      if (EV) {
        const float lj3 = ONETYPE ? onef.lj3 : __ldg(tabf + 2 * n2 + tij);
        const float lj4 = ONETYPE ? onef.lj4 : __ldg(tabf + 3 * n2 + tij);
        const float off = ONETYPE ? onef.offset : __ldg(tabf + 4 * n2 + tij);
        epair = r6inv * fmaf(lj3, r6inv, -lj4) - off;
      }
    };
    auto geom = [&](const int4 &pj, float &dx, float &dy, float &dz, float &rsq, int &tij, float &cutq) {
      dx = (float)(me.x - pj.x);
      dy = (float)(me.y - pj.y);
      dz = (float)(me.z - pj.z);
      rsq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      tij = 0;
      cutq = cutq1;
      if (!ONETYPE) {
        tij = itype * n1 + pj.w;
        cutq = (float)__ldg(tab + tij) * s2;
      }
    };
    auto tally = [&](unsigned e, float dx, float dy, float dz, float f, float ep) {
      const bool fwd = (e & TILE_FWD) != 0;
      ei += fwd ? ep : 0.0f;
      const float w = fwd ? f : 0.0f;
      w0 = fmaf(dx * dx, w, w0); w1 = fmaf(dy * dy, w, w1); w2 = fmaf(dz * dz, w, w2);
      w3 = fmaf(dx * dy, w, w3); w4 = fmaf(dx * dz, w, w4); w5 = fmaf(dy * dz, w, w5);
    };
    // fast path, branch-free; returns true when the decision is too close to call in FP32
    auto body = [&](unsigned e, const int4 &pj) -> bool {
      float dx, dy, dz, rsq, cutq, fp, ep = 0.0f;
      int tij;
      geom(pj, dx, dy, dz, rsq, tij, cutq);
      const bool amb = fabsf(rsq - cutq) < 2.0e-6f * cutq;
      const bool in = rsq < cutq && !amb;
      ljf(rsq, tij, fp, ep);
      const float f = in ? fp : 0.0f;
      gx = fmaf(dx, f, gx);
      gy = fmaf(dy, f, gy);
      gz = fmaf(dz, f, gz);
      if (EV) tally(e, dx, dy, dz, f, in ? ep : 0.0f);
      return amb;
    };
    // an ambiguous entry: the reference's FP64 test on the global positions decides
    auto exact = [&](unsigned e) {
      const int4 pj = ldrec(e);
      float dx, dy, dz, rsq, cutq, fp, ep = 0.0f;
      int tij;
      geom(pj, dx, dy, dz, rsq, tij, cutq);
      if (!(fabsf(rsq - cutq) < 2.0e-6f * cutq)) return;
      const double4 a = ld_xt(xt + gi), b = ld_xt(xt + tile_global_index(H, nlocal, e & TILE_IDX));
      if (rsq_ref(a.x - b.x, a.y - b.y, a.z - b.z) < (ONETYPE ? one.cutsq : __ldg(tab + tij))) {
        ljf(rsq, tij, fp, ep);
        gx = fmaf(dx, fp, gx);
        gy = fmaf(dy, fp, gy);
        gz = fmaf(dz, fp, gz);
        if (EV) tally(e, dx, dy, dz, fp, ep);
      }
    };
    auto entry = [&](const uint4 &w, int k) -> unsigned {
      const unsigned v = k < 2 ? w.x : (k < 4 ? w.y : (k < 6 ? w.z : w.w));
      return (k & 1) ? (v >> 16) : (v & 0xffffu);
    };

    // four entries per step; their records are loaded one step ahead
    int4 cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = ldrec(entry(q0, k));
    for (int k0 = 0; k0 < n; k0 += 8) {
      const uint4 c = q0;
      q0 = q1;
      if (k0 + 16 < n) q1 = __ldg(lp + (size_t)((k0 >> 3) + 2) * NI);
      const uint4 cn = (k0 + 8 < n) ? q0 : c;
      bool again = false;
#pragma unroll
      for (int st = 0; st < 2; st++) {
#pragma unroll
        for (int k = 0; k < 4; k++) nxt[k] = st == 0 ? ldrec(entry(c, 4 + k)) : ldrec(entry(cn, k));
#pragma unroll
        for (int k = 0; k < 4; k++) again |= body(entry(c, st * 4 + k), cur[k]);
#pragma unroll
        for (int k = 0; k < 4; k++) cur[k] = nxt[k];
      }
      if (again) {
#pragma unroll 1
        for (int k = 0; k < 8; k++) {
          const unsigned v = k < 2 ? c.x : (k < 4 ? c.y : (k < 6 ? c.z : c.w));
          exact((k & 1) ? (v >> 16) : (v & 0xffffu));
        }
      }
      // end of a list word: fold the FP32 partial sums into the FP64 accumulators
      fxi += (double)gx; fyi += (double)gy; fzi += (double)gz;
      gx = gy = gz = 0.0f;
    }
    // fixed-point units -> real units: del = dq / scale
    const double dinv = 1.0 / dscale;
    fxi *= dinv; fyi *= dinv; fzi *= dinv;
    if (EV) {
      evdwl += (double)ei;
      const double d2 = dinv * dinv;
      vir[0] += (double)w0 * d2; vir[1] += (double)w1 * d2; vir[2] += (double)w2 * d2;
      vir[3] += (double)w3 * d2; vir[4] += (double)w4 * d2; vir[5] += (double)w5 * d2;
    }
    if (NVE) {
      const double4 p0 = ld_xt(xt + gi);
      double px = p0.x, py = p0.y, pz = p0.z;
      if (imask & nv.groupbit) {
        const double dtfm = nv.dtf / nv.mass[itype];
        const double ka = __dmul_rn(dtfm, fxi), kb = __dmul_rn(dtfm, fyi), kc = __dmul_rn(dtfm, fzi);
        double a = v0, b = v1, cc = v2;
        a = __dadd_rn(__dadd_rn(a, ka), ka);
        b = __dadd_rn(__dadd_rn(b, kb), kb);
        cc = __dadd_rn(__dadd_rn(cc, kc), kc);
        nv.vx[gi] = a; nv.vy[gi] = b; nv.vz[gi] = cc;
        px = __dadd_rn(px, __dmul_rn(nv.dtv, a));
        py = __dadd_rn(py, __dmul_rn(nv.dtv, b));
        pz = __dadd_rn(pz, __dmul_rn(nv.dtv, cc));
      }
      nv.xt_out[gi] = make_double4(px, py, pz, p0.w);
      if (GQ) qout[gi] = q_record(make_double4(px, py, pz, p0.w), Q);  // next step's staged record
      if (nv.do_check) {
        const double dx = px - nv.xhx[gi], dy = py - nv.xhy[gi], dz = pz - nv.xhz[gi];
        if (rsq_ref(dx, dy, dz) > nv.triggersq) *nv.moved = 1;
      }
    } else {
      fx[gi] = fxi;
      fy[gi] = fyi;
      fz[gi] = fzi;
    }
  }
  (void)inv;
  if (EV) {
    double v[7] = {evdwl, vir[0], vir[1], vir[2], vir[3], vir[4], vir[5]};
    __syncthreads();
    block_sum<7>(v, reinterpret_cast<double *>(rec));
    if (tid == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ev[k], v[k]);
  }
}
