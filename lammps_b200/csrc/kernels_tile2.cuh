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
    const int *__restrict__ tile_ids, NveFuse nv) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TileHdr *H = reinterpret_cast<TileHdr *>(tsm);
  double *pos = reinterpret_cast<double *>(tsm + TILE_HDR_BYTES);
  int *stype = reinterpret_cast<int *>(pos + (size_t)3 * scap);
  int *chunk_ctr = reinterpret_cast<int *>(&H->pad0);
  const unsigned pos_s = (unsigned)__cvta_generic_to_shared(pos);
  const int tile = tile_ids ? tile_ids[blockIdx.x] : blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const TilePos P = tile_pos(G, tile);
  if (tid == 0) *chunk_ctr = 0;
  const int S = tile_rows(G, P, ostart, gstart, H);
  if (S + 1 > scap) {  // cannot happen between rebuilds (the rows are those the build staged)
    if (tid == 0) atomicMax(&tflags[5], S + 1);
    return;
  }
  {  // stage: one warp per run of records, cp.async of x,y,z (and type) of each record
    const int nrows = H->nrows;
    for (int r = warp; r < nrows; r += nwarp) {
      const int base = H->rowbase[r], no = H->row_no[r], n = no + H->row_ng[r];
      const int o0 = H->row_o0[r], g0 = nlocal + H->row_g0[r] - no;
      for (int k = lane; k < n; k += 32) {
        const int src = k < no ? o0 + k : g0 + k, s = base + k;
        const double *p = reinterpret_cast<const double *>(xt + src);
        double *d = pos + 3 * s;
        cp_async8(d, p);
        cp_async8(d + 1, p + 1);
        cp_async8(d + 2, p + 2);
        if (!ONETYPE) cp_async4(stype + s, p + 3);
      }
    }
    if (tid < 3) pos[3 * S + tid] = TILE2_FAR;  // the dummy atom padding entries point at
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
    const double pix = pos[3 * li], piy = pos[3 * li + 1], piz = pos[3 * li + 2];
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
#ifdef TILE2_X_NOCONFLICT  // timing experiment only (wrong physics)
      const unsigned a = pos_s + (((e & TILE_IDX) & ~15u) | (threadIdx.x & 15u)) * 24u;
#else
      const unsigned a = pos_s + (e & TILE_IDX) * 24u;
#endif
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(p.x) : "r"(a));
      asm volatile("ld.shared.f64 %0, [%1+8];" : "=d"(p.y) : "r"(a));
      asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(p.z) : "r"(a));
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

