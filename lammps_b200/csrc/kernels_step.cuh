// kernels_step.cuh -- per-timestep streaming kernels: fix nve, displacement check, ghost halo
// (single-device periodic images; the multi-GPU pack/unpack variants live next to them).
// All HBM-bound: one coalesced pass over SoA arrays, 8/16-byte accesses per lane.
#pragma once
#include "common.cuh"

// FixNVE::initial_integrate (fix_nve.cpp:68-108) fused with Neighbor::check_distance
// (neighbor.cpp:2438-2490).  Arithmetic is kept as separate multiply and add (no FMA
// contraction) so that trajectories follow the reference's x86 build bit-for-bit as long as
// forces do.  Algorithmic traffic per atom: x 32r+32w, v 24r+24w, f 24r, mask 4r,
// (+ xhold 24r when checking) = 140 (164) B.
__global__ void __launch_bounds__(256) k_nve_initial(
    int nlocal, double4 *__restrict__ xt, double *__restrict__ vx, double *__restrict__ vy,
    double *__restrict__ vz, const double *__restrict__ fx, const double *__restrict__ fy,
    const double *__restrict__ fz, const int *__restrict__ mask, const double *__restrict__ mass,
    double dtv, double dtf, int groupbit, int do_check, const double *__restrict__ xhx,
    const double *__restrict__ xhy, const double *__restrict__ xhz, double triggersq,
    int *__restrict__ moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  if (mask[i] & groupbit) {
    const double dtfm = dtf / mass[d2type(p.w)];
    double a = vx[i], b = vy[i], c = vz[i];
    a = __dadd_rn(a, __dmul_rn(dtfm, fx[i]));
    b = __dadd_rn(b, __dmul_rn(dtfm, fy[i]));
    c = __dadd_rn(c, __dmul_rn(dtfm, fz[i]));
    vx[i] = a; vy[i] = b; vz[i] = c;
    p.x = __dadd_rn(p.x, __dmul_rn(dtv, a));
    p.y = __dadd_rn(p.y, __dmul_rn(dtv, b));
    p.z = __dadd_rn(p.z, __dmul_rn(dtv, c));
    xt[i] = p;
  }
  if (do_check) {
    const double dx = p.x - xhx[i], dy = p.y - xhy[i], dz = p.z - xhz[i];
    const double rsq = rsq_ref(dx, dy, dz);
    if (rsq > triggersq) *moved = 1;
  }
}

// FixNVE::final_integrate (fix_nve.cpp:112-145).  v 24r+24w, f 24r, mask 4r, type 8r = 84 B.
__global__ void __launch_bounds__(256) k_nve_final(
    int nlocal, const double4 *__restrict__ xt, double *__restrict__ vx, double *__restrict__ vy,
    double *__restrict__ vz, const double *__restrict__ fx, const double *__restrict__ fy,
    const double *__restrict__ fz, const int *__restrict__ mask, const double *__restrict__ mass,
    double dtf, int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (mask[i] & groupbit) {
    const double w = reinterpret_cast<const double *>(xt)[4 * (size_t)i + 3];
    const double dtfm = dtf / mass[d2type(w)];
    vx[i] = __dadd_rn(vx[i], __dmul_rn(dtfm, fx[i]));
    vy[i] = __dadd_rn(vy[i], __dmul_rn(dtfm, fy[i]));
    vz[i] = __dadd_rn(vz[i], __dmul_rn(dtfm, fz[i]));
  }
}

// The pieces of FixNH's integrator (fix nvt: fix_nh.cpp:916-1014) that loop over atoms.  The
// Nose-Hoover chain itself is scalar host arithmetic and stays in the reference's own FixNH code
// (the package's fix nvt/b200 inherits it); what runs here is
//   nve_v     v += dtf/m * f           = k_nve_final above with the fix's dtf and group
//   nve_x     x += dtv * v             (fix_nh.cpp:2278-2298) + Neighbor::check_distance when due
//   nh_v_temp v *= factor_eta          (fix_nh.cpp:2338-2352, no bias)
// each operation rounded separately like the reference loops.
__global__ void __launch_bounds__(256) k_nve_x(
    int nlocal, double4 *__restrict__ xt, const double *__restrict__ vx, const double *__restrict__ vy,
    const double *__restrict__ vz, const int *__restrict__ mask, double dtv, int groupbit, int do_check,
    const double *__restrict__ xhx, const double *__restrict__ xhy, const double *__restrict__ xhz,
    double triggersq, int *__restrict__ moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  if (mask[i] & groupbit) {
    p.x = __dadd_rn(p.x, __dmul_rn(dtv, vx[i]));
    p.y = __dadd_rn(p.y, __dmul_rn(dtv, vy[i]));
    p.z = __dadd_rn(p.z, __dmul_rn(dtv, vz[i]));
    xt[i] = p;
  }
  if (do_check) {
    const double dx = p.x - xhx[i], dy = p.y - xhy[i], dz = p.z - xhz[i];
    if (rsq_ref(dx, dy, dz) > triggersq) *moved = 1;
  }
}

__global__ void __launch_bounds__(256) k_scale_v(int nlocal, double *__restrict__ vx,
                                                 double *__restrict__ vy, double *__restrict__ vz,
                                                 const int *__restrict__ mask, double factor,
                                                 int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (mask[i] & groupbit) {
    vx[i] = __dmul_rn(vx[i], factor);
    vy[i] = __dmul_rn(vy[i], factor);
    vz[i] = __dmul_rn(vz[i], factor);
  }
}

// fix npt / nph (FixNH with a barostat, orthogonal box): nh_v_press scales the velocity components
// twice by exp(-dt4 (omega_dot + mtk_term2)) (fix_nh.cpp:2227-2252), remap() dilates the atoms
// with the box: x -> lamda in the old box (Domain::x2lamda, domain.cpp) -> x in the new box
// (Domain::lamda2x), same operations, each rounded separately.
__global__ void __launch_bounds__(256) k_scale_v3(int nlocal, double *__restrict__ vx,
                                                  double *__restrict__ vy, double *__restrict__ vz,
                                                  const int *__restrict__ mask, double f0, double f1,
                                                  double f2, int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (mask[i] & groupbit) {
    vx[i] = __dmul_rn(__dmul_rn(vx[i], f0), f0);
    vy[i] = __dmul_rn(__dmul_rn(vy[i], f1), f1);
    vz[i] = __dmul_rn(__dmul_rn(vz[i], f2), f2);
  }
}

// ---------------------------------------------------------------------------------------
// Deferred per-atom integrator operations.  A host-driven integrator (fix nvt / npt / nph, fix nve
// beside fix langevin) issues its per-atom loops one by one: scale, half-kick, drift, each a full
// pass over v (and x, f).  None of them is needed before the next consumer of x or v (the rebuild
// vote, the halo, the pair stage, a temperature sum, a download), so they are queued and executed
// in ONE pass, operation after operation in registers, each rounded exactly as the separate
// kernels round it (k_scale_v, k_scale_v3, k_nve_final, k_nve_x).  KE: the same pass ends with
// ComputeTemp's sums (k_ke_group) -- the temperature a Nose-Hoover chain reads after its half-kick.
// ---------------------------------------------------------------------------------------
#define VOP_MAX 8
enum { VOP_SCALE = 1, VOP_SCALE3 = 2, VOP_KICK = 3, VOP_DRIFT = 4, VOP_REMAP = 5 };
struct VOps {
  int n;
  int kind[VOP_MAX], groupbit[VOP_MAX];
  // SCALE: a[0]; SCALE3: per-component factor applied twice; KICK: dtf; DRIFT: dtv;
  // REMAP (FixNH::remap, k_remap): oldlo[3], 1/oldprd[3], newlo[3], newprd[3]
  double a[VOP_MAX][12];
};

template <bool KE>
__global__ void __launch_bounds__(256) k_vops(
    int nlocal, double4 *__restrict__ xt, double *__restrict__ vx, double *__restrict__ vy,
    double *__restrict__ vz, const double *__restrict__ fx, const double *__restrict__ fy,
    const double *__restrict__ fz, const int *__restrict__ mask, const double *__restrict__ mass, VOps ops,
    int do_check, const double *__restrict__ xhx, const double *__restrict__ xhy,
    const double *__restrict__ xhz, double triggersq, int *__restrict__ moved, int ke_groupbit,
    double *__restrict__ ke_out) {
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  bool has_kick = false, has_drift = false;
  for (int k = 0; k < ops.n; k++) {
    has_kick |= ops.kind[k] == VOP_KICK;
    has_drift |= ops.kind[k] == VOP_DRIFT;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
    double4 p = xt[i];
    const int m = mask[i];
    const double mi = mass[d2type(p.w)];
    double a = vx[i], b = vy[i], c = vz[i];
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    if (has_kick) { f0 = fx[i]; f1 = fy[i]; f2 = fz[i]; }
    bool vdirty = false, xdirty = false;
    for (int k = 0; k < ops.n; k++) {
      if (!(m & ops.groupbit[k])) continue;
      const int kind = ops.kind[k];
      if (kind == VOP_SCALE) {
        const double q = ops.a[k][0];
        a = __dmul_rn(a, q); b = __dmul_rn(b, q); c = __dmul_rn(c, q);
        vdirty = true;
      } else if (kind == VOP_SCALE3) {
        a = __dmul_rn(__dmul_rn(a, ops.a[k][0]), ops.a[k][0]);
        b = __dmul_rn(__dmul_rn(b, ops.a[k][1]), ops.a[k][1]);
        c = __dmul_rn(__dmul_rn(c, ops.a[k][2]), ops.a[k][2]);
        vdirty = true;
      } else if (kind == VOP_KICK) {
        const double dtfm = ops.a[k][0] / mi;
        a = __dadd_rn(a, __dmul_rn(dtfm, f0));
        b = __dadd_rn(b, __dmul_rn(dtfm, f1));
        c = __dadd_rn(c, __dmul_rn(dtfm, f2));
        vdirty = true;
      } else if (kind == VOP_DRIFT) {
        const double dtv = ops.a[k][0];
        p.x = __dadd_rn(p.x, __dmul_rn(dtv, a));
        p.y = __dadd_rn(p.y, __dmul_rn(dtv, b));
        p.z = __dadd_rn(p.z, __dmul_rn(dtv, c));
        xdirty = true;
      } else {  // VOP_REMAP: x -> lamda in the old box -> x in the new box, as k_remap
        const double *q = ops.a[k];
        const double l0 = __dmul_rn(q[3], __dadd_rn(p.x, -q[0]));
        const double l1 = __dmul_rn(q[4], __dadd_rn(p.y, -q[1]));
        const double l2 = __dmul_rn(q[5], __dadd_rn(p.z, -q[2]));
        p.x = __dadd_rn(__dmul_rn(q[9], l0), q[6]);
        p.y = __dadd_rn(__dmul_rn(q[10], l1), q[7]);
        p.z = __dadd_rn(__dmul_rn(q[11], l2), q[8]);
        xdirty = true;
      }
    }
    if (vdirty) { vx[i] = a; vy[i] = b; vz[i] = c; }
    if (xdirty) xt[i] = p;
    if (has_drift && do_check) {
      const double dx = p.x - xhx[i], dy = p.y - xhy[i], dz = p.z - xhz[i];
      if (rsq_ref(dx, dy, dz) > triggersq) *moved = 1;
    }
    if (KE && (m & ke_groupbit)) {
      s[0] += (a * a + b * b + c * c) * mi;
      s[1] += mi * a * a; s[2] += mi * b * b; s[3] += mi * c * c;
      s[4] += mi * a * b; s[5] += mi * a * c; s[6] += mi * b * c;
    }
  }
  if (KE) {
    __shared__ double red[7 * 32];
    block_sum<7>(s, red);
    if (threadIdx.x == 0)
      for (int k = 0; k < 7; k++) atomicAdd(&ke_out[k], s[k]);
  }
}

struct RemapBox {
  double oldlo[3], oldhinv[3], newlo[3], newh[3];
};

__global__ void __launch_bounds__(256) k_remap(int nlocal, double4 *__restrict__ xt,
                                               const int *__restrict__ mask, int groupbit, RemapBox B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (!(mask[i] & groupbit)) return;
  double4 p = xt[i];
  const double l0 = __dmul_rn(B.oldhinv[0], __dadd_rn(p.x, -B.oldlo[0]));
  const double l1 = __dmul_rn(B.oldhinv[1], __dadd_rn(p.y, -B.oldlo[1]));
  const double l2 = __dmul_rn(B.oldhinv[2], __dadd_rn(p.z, -B.oldlo[2]));
  p.x = __dadd_rn(__dmul_rn(B.newh[0], l0), B.newlo[0]);
  p.y = __dadd_rn(__dmul_rn(B.newh[1], l1), B.newlo[1]);
  p.z = __dadd_rn(__dmul_rn(B.newh[2], l2), B.newlo[2]);
  xt[i] = p;
}

// Neighbor::check_distance on its own (neighbor.cpp:2438-2490): with a changing box the trigger
// distance shrinks with the box corners and the test must see the positions decide() sees
__global__ void __launch_bounds__(256) k_check_distance(int nlocal, const double4 *__restrict__ xt,
                                                        const double *__restrict__ xhx,
                                                        const double *__restrict__ xhy,
                                                        const double *__restrict__ xhz, double deltasq,
                                                        int *__restrict__ moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const double4 p = xt[i];
  if (rsq_ref(p.x - xhx[i], p.y - xhy[i], p.z - xhz[i]) > deltasq) *moved = 1;
}

// final_integrate of step n immediately followed by initial_integrate of step n+1 in ONE pass
// over the atoms (both use the same forces f(n)): v += dtfm*f ; v += dtfm*f ; x += dtv*v, each
// operation rounded separately exactly like the two reference loops.  Used whenever nothing
// on the host needs the full-step velocities in between; saves the 84 B/atom of a separate
// final_integrate pass (x 32r+32w, v 24r+24w, f 24r, mask 4r = 140 B for both half-kicks).
__global__ void __launch_bounds__(256) k_nve_final_initial(
    int nlocal, double4 *__restrict__ xt, double *__restrict__ vx, double *__restrict__ vy,
    double *__restrict__ vz, const double *__restrict__ fx, const double *__restrict__ fy,
    const double *__restrict__ fz, const int *__restrict__ mask, const double *__restrict__ mass,
    double dtv, double dtf, int groupbit, int do_check, const double *__restrict__ xhx,
    const double *__restrict__ xhy, const double *__restrict__ xhz, double triggersq,
    int *__restrict__ moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  double4 p = xt[i];
  if (mask[i] & groupbit) {
    const double dtfm = dtf / mass[d2type(p.w)];
    const double ka = __dmul_rn(dtfm, fx[i]), kb = __dmul_rn(dtfm, fy[i]), kc = __dmul_rn(dtfm, fz[i]);
    double a = vx[i], b = vy[i], c = vz[i];
    a = __dadd_rn(__dadd_rn(a, ka), ka);  // final_integrate(n), then the half-kick of n+1
    b = __dadd_rn(__dadd_rn(b, kb), kb);
    c = __dadd_rn(__dadd_rn(c, kc), kc);
    vx[i] = a; vy[i] = b; vz[i] = c;
    p.x = __dadd_rn(p.x, __dmul_rn(dtv, a));
    p.y = __dadd_rn(p.y, __dmul_rn(dtv, b));
    p.z = __dadd_rn(p.z, __dmul_rn(dtv, c));
    xt[i] = p;
  }
  if (do_check) {
    const double dx = p.x - xhx[i], dy = p.y - xhy[i], dz = p.z - xhz[i];
    const double rsq = rsq_ref(dx, dy, dz);
    if (rsq > triggersq) *moved = 1;
  }
}

__global__ void __launch_bounds__(256) k_fill_int(int n, int value, int *__restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = value;
}

// ComputeTemp::compute_scalar numerator (compute_temp.cpp:73-97): ev[7] += sum m v^2
__global__ void __launch_bounds__(256) k_ke(int nlocal, const double4 *__restrict__ xt,
                                            const double *__restrict__ vx,
                                            const double *__restrict__ vy,
                                            const double *__restrict__ vz,
                                            const int *__restrict__ mask,
                                            const double *__restrict__ mass, int groupbit,
                                            double *__restrict__ ev) {
  double v[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
    if (mask[i] & groupbit) {
      const double w = reinterpret_cast<const double *>(xt)[4 * (size_t)i + 3];
      v[0] += (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]) * mass[d2type(w)];
    }
  }
  __shared__ double red[32];
  block_sum<1>(v, red);
  if (threadIdx.x == 0) atomicAdd(&ev[7], v[0]);
}

// ComputeTemp::compute_scalar + compute_vector (compute_temp.cpp:73-140) for any group bit:
// out[0] += sum m v^2, out[1..6] += sum m (vx vx, vy vy, vz vz, vx vy, vx vz, vy vz)
__global__ void __launch_bounds__(256) k_ke_group(int nlocal, const double4 *__restrict__ xt,
                                                  const double *__restrict__ vx,
                                                  const double *__restrict__ vy,
                                                  const double *__restrict__ vz,
                                                  const int *__restrict__ mask,
                                                  const double *__restrict__ mass, int groupbit,
                                                  double *__restrict__ out) {
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
    if (mask[i] & groupbit) {
      const double w = reinterpret_cast<const double *>(xt)[4 * (size_t)i + 3];
      const double m = mass[d2type(w)], a = vx[i], b = vy[i], c = vz[i];
      v[0] += (a * a + b * b + c * c) * m;
      v[1] += m * a * a; v[2] += m * b * b; v[3] += m * c * c;
      v[4] += m * a * b; v[5] += m * a * c; v[6] += m * b * c;
    }
  }
  __shared__ double red[7 * 32];
  block_sum<7>(v, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 7; k++) atomicAdd(&out[k], v[k]);
}

// fix langevin (FixLangevin::post_force_templated<0,0,0,0,ZERO>, fix_langevin.cpp:383-507: constant
// or equal-style target temperature, per-type masses, no bias, no tally):
//   f_i += gfactor1[type] v_i + gfactor2[type] sqrt(T) (u - 0.5),  u uniform in (0,1), three per atom,
// each operation rounded separately like the reference loop.  The reference draws u from one
// sequential Marsaglia stream in host atom order, an order a bin-sorted, multi-sub-domain device
// layout does not have; the device's own stream is counter based instead -- Philox-4x32-10 keyed
// by the fix's seed with the counter (atom tag, timestep): one block of four 32-bit words per
// atom and step, independent of atom order, sub-domain count and launch shape.  For verification
// against the reference a host may hand in the uniforms itself (`uni`, indexed by tag), e.g. drawn
// from the reference's RanMars in tag order.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// gf[t] = gfactor1[t], gf[ntypes+1+t] = gfactor2[t]*tsqrt.  fsum (nullable) += sum of the random
// forces of the group (zero yes: the host subtracts the group mean afterwards, k_add_force).
__global__ void __launch_bounds__(256) k_langevin(
    int nlocal, const double4 *__restrict__ xt, const double *__restrict__ vx, const double *__restrict__ vy,
    const double *__restrict__ vz, double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz,
    const int *__restrict__ tag, const int *__restrict__ mask, int groupbit, int ntypes,
    const double *__restrict__ gf, uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo, uint32_t step_hi,
    const double *__restrict__ uni, long long nuni, double *__restrict__ fsum) {
  double s[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
    if (!(mask[i] & groupbit)) continue;
    const int t = d2type(reinterpret_cast<const double *>(xt)[4 * (size_t)i + 3]);
    const double g1 = gf[t], g2 = gf[ntypes + 1 + t];
    double u[3];
    const long long k = 3LL * (tag[i] - 1);
    if (uni && k + 2 < nuni) {
      u[0] = uni[k]; u[1] = uni[k + 1]; u[2] = uni[k + 2];
    } else {
      uint32_t r[4];
      philox4x32_10((uint32_t)tag[i], step_lo, step_hi, 0u, seed_lo, seed_hi, r);
#pragma unroll
      for (int d = 0; d < 3; d++) u[d] = ((double)r[d] + 0.5) * 2.3283064365386963e-10;  // 2^-32
    }
    const double fr0 = __dmul_rn(g2, __dadd_rn(u[0], -0.5));
    const double fr1 = __dmul_rn(g2, __dadd_rn(u[1], -0.5));
    const double fr2 = __dmul_rn(g2, __dadd_rn(u[2], -0.5));
    fx[i] = __dadd_rn(fx[i], __dadd_rn(__dmul_rn(g1, vx[i]), fr0));
    fy[i] = __dadd_rn(fy[i], __dadd_rn(__dmul_rn(g1, vy[i]), fr1));
    fz[i] = __dadd_rn(fz[i], __dadd_rn(__dmul_rn(g1, vz[i]), fr2));
    s[0] += fr0; s[1] += fr1; s[2] += fr2;
  }
  if (fsum) {  // uniform branch: fsum is a kernel argument
    __shared__ double red[3 * 32];
    block_sum<3>(s, red);
    if (threadIdx.x == 0)
      for (int d = 0; d < 3; d++) atomicAdd(&fsum[d], s[d]);
  }
}

// f_i += df for the atoms of a group (fix langevin zero yes: df = -sum(fran)/count, :481-497)
__global__ void __launch_bounds__(256) k_add_force(int nlocal, double *__restrict__ fx,
                                                   double *__restrict__ fy, double *__restrict__ fz,
                                                   const int *__restrict__ mask, int groupbit, double d0,
                                                   double d1, double d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  if (mask[i] & groupbit) {
    fx[i] = __dadd_rn(fx[i], d0);
    fy[i] = __dadd_rn(fy[i], d1);
    fz[i] = __dadd_rn(fz[i], d2);
  }
}

// host array <-> device layout converters (b200_set_atoms / b200_get_atoms)
__global__ void __launch_bounds__(256) k_pack_xt(int n, const double *__restrict__ x3,
                                                 const int *__restrict__ type,
                                                 double4 *__restrict__ xt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xt[i] = make_double4(x3[3 * (size_t)i], x3[3 * (size_t)i + 1], x3[3 * (size_t)i + 2],
                       type2d(type[i]));
}
__global__ void __launch_bounds__(256) k_unpack_xt(int n, const double4 *__restrict__ xt,
                                                   double *__restrict__ x3, int *__restrict__ type) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xt[i];
  if (x3) { x3[3 * (size_t)i] = p.x; x3[3 * (size_t)i + 1] = p.y; x3[3 * (size_t)i + 2] = p.z; }
  if (type) type[i] = d2type(p.w);
}
__global__ void __launch_bounds__(256) k_aos_to_soa(int n, const double *__restrict__ a3,
                                                    double *__restrict__ x, double *__restrict__ y,
                                                    double *__restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = a3[3 * (size_t)i]; y[i] = a3[3 * (size_t)i + 1]; z[i] = a3[3 * (size_t)i + 2];
}
__global__ void __launch_bounds__(256) k_soa_to_aos(int n, const double *__restrict__ x,
                                                    const double *__restrict__ y,
                                                    const double *__restrict__ z,
                                                    double *__restrict__ a3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a3[3 * (size_t)i] = x[i]; a3[3 * (size_t)i + 1] = y[i]; a3[3 * (size_t)i + 2] = z[i];
}
