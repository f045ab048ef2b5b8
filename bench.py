#!/usr/bin/env python3
"""bench.py -- atom-timesteps/s of the B200-native short-range MD hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload lj32m|lj4m|lj32k|eam16m|eam2m|eam32k] [--precision double|mixed]

A "step" is ONE MD timestep of the hot path (fix nve + ghost halo + pair forces, with the
neighbour list rebuilt on the reference's own schedule) over all atoms of a synthetic LJ-melt /
Cu-EAM FCC lattice built exactly like bench/in.lj / bench/in.eam of the reference.

Default workload (BASELINE.json config the target is quoted on): LJ melt, 32,000,000 atoms,
lj/cut 2.5 sigma, skin 0.3, rebuild every 20 steps, NVE -- strong-scaled over N GPUs.

Prints ONE JSON line (see the keys at the bottom).  `value` is device-resident steady-state
throughput (CUDA events around K timesteps, max over ranks); `e2e` is the same metric through
the C ABI with HOST buffers: upload from pinned host memory, Verlet setup, K timesteps with the
thermo read-back, download of x/v/f -- all inside the timed region.

--impl reference times the UNMODIFIED reference (oracle/_ref/lmp_ref, built by
oracle/build_ref.py) on the host cores with its OPENMP package on a bounded sample of the
same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# every kernel of libb200md.so is loaded when the library is, not at its first launch (the
# kernels that migrate atoms between sub-domains first run inside the timed region otherwise)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (kind, cells per edge at N=1, scaling)
    "lj32m": ("lj", 200, "strong"),
    "lj4m": ("lj", 100, "strong"),
    "lj32k": ("lj", 20, "strong"),
    "eam16m": ("eam", 160, "strong"),
    "eam2m": ("eam", 80, "weak"),
    "eam32k": ("eam", 20, "strong"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_system(kind, cells):
    from lammps_b200 import eam as eam_mod
    from lammps_b200 import lattice, pair_lj
    if kind == "lj":
        x, lo, hi = lattice.fcc_block("lj", 0.8442, cells)
        mass = np.array([0.0, 1.0])
        typ = np.ones(len(x), np.int32)
        v = lattice.velocity_create(x, typ, mass, 1.44, 87287, "lj")
        return dict(kind="lj", units="lj", x=x, v=v, type=typ, mass=mass, lo=lo, hi=hi, skin=0.3,
                    every=20, delay=0, check=False, dt=0.005,
                    tables=pair_lj.lj_cut_tables(1, {(1, 1): (1.0, 1.0, 2.5)}, 2.5))
    d = np.load(ROOT / "lammps_b200" / "data" / "Cu_u3_funcfl.npz")
    f = eam_mod.Funcfl(float(d["mass"]), int(d["nrho"]), float(d["drho"]), int(d["nr"]),
                       float(d["dr"]), float(d["cut"]), d["frho"], d["zr"], d["rhor"])
    T = eam_mod.funcfl_tables([f], [0])
    x, lo, hi = lattice.fcc_block("metal", 3.615, cells)
    typ = np.ones(len(x), np.int32)
    v = lattice.velocity_create(x, typ, T.mass, 1600.0, 376847, "metal")
    return dict(kind="eam", units="metal", x=x, v=v, type=typ, mass=T.mass, lo=lo, hi=hi, skin=1.0,
                every=1, delay=5, check=True, dt=0.005, tables=T.as_dict())


def configure(e, s, x, v, typ, tag, natoms_total):
    e.set_box(s["lo"], s["hi"])
    e.set_atoms(x, v, typ, tag, s["mass"], natoms_total=natoms_total)
    e.neighbor(s["skin"], every=s["every"], delay=s["delay"], check=s["check"])
    e.fix_nve(s["dt"])
    (e.pair_lj_cut if s["kind"] == "lj" else e.pair_eam)(s["tables"])


# algorithmic bytes/flops per atom-step of the dominant (pair) kernel, SURVEY.md 8(d):
#   LJ : list 4*Ps + x_i,type 32 + f_i write 24 + f clear 24 ; flops 8*Ps + 20*Pc
#   EAM: 2 passes over the list, rho/fp traffic
# Ps = stored half-list pairs per atom (measured from the list), Pc = in-cutoff pairs.
# The bin-tile list moves the same list bytes in another shape (2 B x 2*Ps entries instead of
# 4 B x Ps) and needs no force clear, but evaluates owned-owned pairs from both sides; the
# roofline keeps SURVEY's half-list figures so that numbers stay comparable between rounds.
def pair_algorithmic(kind, ps, pc):
    if kind == "lj":
        return 4.0 * ps + 32 + 24 + 24, 8.0 * ps + 20.0 * pc
    return 2 * 4.0 * ps + 2 * 32 + 8 * 3 + 24 + 24, 2 * 8.0 * ps + 24.0 * pc + 45.0 * pc


def measure_pc_ratio(kind):
    """In-cutoff fraction of the stored half-list pairs, measured on the list itself: a 32k-atom
    melt of the same state point (an intensive property), after 40 steps, through the engine's
    list export.  Returns Pc/Ps."""
    from lammps_b200.engine import Engine
    s = build_system(kind, (20, 20, 20))
    e = Engine(int(os.environ.get("LOCAL_RANK", "0")), "double", s["units"])
    n = len(s["x"])
    configure(e, s, s["x"], s["v"], s["type"], np.arange(1, n + 1, dtype=np.int32), n)
    e.setup(0, 0)
    e.run(40, 0)
    # the list in force after those steps and the positions it is evaluated on
    a = e.get_atoms(ghosts=True, fields=("x",))
    _, pi, pj = e.neighbor_list()
    d = a["x"][pi] - a["x"][pj]
    rsq = (d * d).sum(axis=1)
    cutsq = float(np.max(s["tables"]["cutsq"])) if kind == "lj" else float(s["tables"]["cutforcesq"])
    ratio = float((rsq < cutsq).mean())
    e.close()
    return ratio


def fp_peaks():
    """FP64 / FP32 FMA peaks of this GPU from the repo's microbenchmark (tools/microbench/fp_peak,
    built by __graft_entry__.build); None when the binary is missing."""
    exe = ROOT / "tools" / "microbench" / "fp_peak"
    if not exe.exists():
        return None
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout
    except Exception:
        return None
    pk = {}
    for ln in out.splitlines():
        m = re.match(r"(DFMA|FFMA)\s+ILP=8 grid=296x1024 .* ([0-9.]+) TFLOP/s", ln)
        if m:
            pk["fp64" if m.group(1) == "DFMA" else "fp32"] = float(m.group(2))
    return pk or None


def also_block(args, local_rank):
    """Other configurations of BASELINE.json, measured after the timed region of the headline
    (one GPU): device-resident atom-steps/s over 100 steps after preconditioning."""
    from lammps_b200.engine import Engine
    res = {}

    def quick(workload, precision, steps=100, tilt=None):
        kind, cells1, _ = WORKLOADS[workload]
        s = build_system(kind, (cells1,) * 3)
        n = len(s["x"])
        e = Engine(local_rank, precision, s["units"])
        configure(e, s, s["x"], s["v"], s["type"], np.arange(1, n + 1, dtype=np.int32), n)
        if tilt is not None:
            # the same crystal in a sheared box: tilt factors are whole lattice constants, so the
            # lattice stays periodic; atoms outside the parallelepiped are wrapped at setup
            a = (s["hi"][0] - s["lo"][0]) / cells1
            e.set_box_triclinic(s["lo"], s["hi"], tilt[0] * a, tilt[1] * a, tilt[2] * a)
        e.setup(1, 1)
        e.run(20, 0)
        e.run(steps, 0)
        ms = e.last_run_ms()
        e.profiling(True)
        e.run(steps, 0)
        ph = e.phase_times()
        e.close()
        pair_ms, pair_calls = ph["pair"]
        return {"value": n * steps / (ms * 1e-3), "unit": "atom-steps/s", "natoms": n, "steps": steps,
                "ms_per_step": ms / steps, "pair_us_per_step": pair_ms * 1e3 / max(pair_calls, 1),
                "precision": precision}

    for name, wl, prec in (("lj4m_double", "lj4m", "double"), ("lj4m_mixed", "lj4m", "mixed"),
                           ("eam2m_double", "eam2m", "double"), ("eam2m_mixed", "eam2m", "mixed")):
        try:
            res[name] = quick(wl, prec)
        except Exception as ex:  # a side measurement must not take the headline down
            res[name] = {"error": str(ex)[:200]}
    try:  # triclinic box (bin tiles, half list by the tag rule: k_tile_build<...,TRI>)
        res["lj4m_triclinic_double"] = quick("lj4m", "double", tilt=(2.0, -1.0, 3.0))
        res["lj4m_triclinic_double"]["note"] = "prism box, tilt (2, -1, 3) lattice constants"
    except Exception as ex:
        res["lj4m_triclinic_double"] = {"error": str(ex)[:200]}
    # the reference's own small inputs, unmodified, through the LAMMPS package
    exe = ROOT / "lammps_b200" / "lammps_pkg" / "lmp_b200"
    inputs = ROOT / "lammps_b200" / "lammps_pkg" / "bench_inputs"
    for name, inp in (("in.lj_32k_lmp_b200", "in.lj"), ("in.eam_32k_lmp_b200", "in.eam")):
        if not exe.exists() or not (inputs / inp).exists():
            continue
        try:
            out = subprocess.run([str(exe), "-sf", "b200", "-in", inp], cwd=inputs, capture_output=True,
                                 text=True, timeout=300).stdout
            m = re.search(r"Loop time of ([0-9.eE+-]+) on \d+ procs for (\d+) steps with (\d+) atoms", out)
            t, nst, nat = float(m.group(1)), int(m.group(2)), int(m.group(3))
            res[name] = {"value": nat * nst / t, "unit": "atom-steps/s", "natoms": nat, "steps": nst,
                         "note": "unmodified bench input, lmp_b200 -sf b200, LAMMPS loop time"}
        except Exception as ex:
            res[name] = {"error": str(ex)[:200]}
    # thermostatted 4 M-atom melts through the LAMMPS package (stage-by-stage timestep: the
    # integrator cannot live inside the pair kernel when a fix acts between the two half-kicks)
    for name, fixes in (("lj4m_nvt_lmp_b200", "fix 1 all nvt temp 1.44 1.44 0.5"),
                        ("lj4m_langevin_lmp_b200", "fix 1 all nve\nfix 2 all langevin 1.44 1.44 1.0 48279")):
        if not exe.exists():
            continue
        try:
            script = ("units lj\nlattice fcc 0.8442\nregion box block 0 100 0 100 0 100\ncreate_box 1 box\n"
                      "create_atoms 1 box\nmass 1 1.0\nvelocity all create 1.44 87287 loop geom\n"
                      "pair_style lj/cut 2.5\npair_coeff 1 1 1.0 1.0 2.5\nneighbor 0.3 bin\n"
                      "neigh_modify delay 0 every 20 check no\n" + fixes + "\nthermo 100\nrun 20\nrun 100\n")
            out = subprocess.run([str(exe), "-sf", "b200", "-echo", "none"], input=script, capture_output=True,
                                 text=True, timeout=300).stdout
            m = re.findall(r"Loop time of ([0-9.eE+-]+) on \d+ procs for (\d+) steps with (\d+) atoms", out)[-1]
            t, nst, nat = float(m[0]), int(m[1]), int(m[2])
            res[name] = {"value": nat * nst / t, "unit": "atom-steps/s", "natoms": nat, "steps": nst,
                         "note": fixes.replace("\n", "; ") + ", lmp_b200 -sf b200, LAMMPS loop time"}
        except Exception as ex:
            res[name] = {"error": str(ex)[:200]}
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from lammps_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries print there meanwhile (NCCL's
    # version banner) is sent to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind, cells1, scaling = WORKLOADS[args.workload]

    from lammps_b200 import decomp
    grid = decomp.proc_grid(world, (1.0, 1.0, 1.0))
    if scaling == "weak":
        cells = tuple(cells1 * g for g in grid)
    else:
        cells = (cells1,) * 3
    t0 = time.time()
    s = build_system(kind, cells)
    natoms = len(s["x"])
    tag = np.arange(1, natoms + 1, dtype=np.int32)
    host_setup_s = time.time() - t0

    e = Engine(local_rank, args.precision, s["units"])
    if world > 1:
        myloc = decomp.rank_to_loc(rank, grid)
        sel = decomp.owned_mask(s["x"], s["lo"], s["hi"], grid, myloc)
        xs, vs, ts, tg = s["x"][sel], s["v"][sel], s["type"][sel], tag[sel]
        e.set_decomposition(grid, myloc)
        decomp.init_comm(e, dist, rank, world)
    else:
        xs, vs, ts, tg = s["x"], s["v"], s["type"], tag
    nown = len(xs)

    # pinned host staging for the end-to-end leg
    pin = {k: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
           for k, a in (("x", xs), ("v", vs), ("type", ts), ("tag", tg))}
    hx, hv, ht, hg = (pin[k].numpy() for k in ("x", "v", "type", "tag"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(val):
        if world == 1:
            return val
        t = torch.tensor([val], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg
    configure(e, s, hx, hv, ht, hg, natoms)
    e.setup(1, 1)
    # preconditioning: one full neighbour-list period, so that every code path of a timestep
    # (rebuild with atom migration between sub-domains included) has run once -- first-use
    # allocations, CUDA IPC mappings and NCCL connections are start-up cost, not throughput --
    # and the W warm-up steps after it leave the list `W` steps old, as they would mid-run
    precond = s["every"] if not s["check"] else 10
    e.run(precond, 0)
    e.run(args.warmup, 0)
    st0 = e.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    e.run(args.steps, 0)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = maxreduce(e.last_run_ms())
    clocks = sampler.stop() if rank == 0 else None
    st = e.stats()
    launches = st["launches"] - st0["launches"]
    rebuilds = st["nbuilds"] - st0["nbuilds"]
    value = natoms * args.steps / (dev_ms * 1e-3)
    # per-phase times come from a second, untimed pass of the same length: the phase events sit
    # on one stream, so with them on the engine keeps halo and pair kernels serial (no overlap)
    e.profiling(True)
    e.run(args.steps, 0)
    prof_ms = maxreduce(e.last_run_ms())
    ph = e.phase_times()
    e.profiling(False)

    # ---------------- end-to-end leg through the C ABI with host buffers
    e2e_steps = args.steps
    cap = int(nown * 1.25) + 4096   # room for migration imbalance between ranks
    out = {k: torch.empty((cap, 3), dtype=torch.float64).pin_memory().numpy() for k in ("x", "v", "f")}
    barrier()
    t0 = time.perf_counter()
    e.step = 0
    configure(e, s, hx, hv, ht, hg, natoms)       # H2D: x,v (24 B each), type, tag (4 B each)
    e.sync()
    t1 = time.perf_counter()
    e.setup(1, 1)
    e.sync()
    t2 = time.perf_counter()
    th = e.run(e2e_steps, 100)                    # thermo tallies read back every 100 steps
    t3 = time.perf_counter()
    got = e.get_atoms(fields=("x", "v", "f"), into=out)   # D2H: x,v,f into pinned host buffers
    barrier()
    t4 = time.perf_counter()
    e2e_s = maxreduce(t4 - t0)
    e2e_parts = {"upload_ms": (t1 - t0) * 1e3, "setup_ms": (t2 - t1) * 1e3, "run_ms": (t3 - t2) * 1e3,
                 "download_ms": (t4 - t3) * 1e3}
    e2e_value = natoms * e2e_steps / e2e_s
    h2d = nown * (24 + 24 + 4 + 4)
    d2h = len(got["x"]) * 72 + len(th) * 80

    # ---------------- roofline of the dominant kernel (pair)
    hbm_peak, peak_src = peaks()
    pair_ms, pair_calls = ph["pair"]
    ps = st["npairs"] / max(nown, 1)
    pc_ratio = measure_pc_ratio(kind) if rank == 0 else 0.0     # measured on the list
    pc = ps * pc_ratio
    bytes_atom, flops_atom = pair_algorithmic(kind, ps, pc)
    pair_s = pair_ms * 1e-3 / max(pair_calls, 1)
    achieved = bytes_atom * nown / pair_s / 1e9
    # DRAM traffic per launch from the committed `ncu --set full` capture of the same kernel
    # (dram__bytes_read.sum + dram__bytes_write.sum per atom, profiles/ncu_traffic.json), scaled
    # to this launch's atom count; null when no capture exists for this kernel / precision
    traffic, traffic_src = None, None
    tiled = st["list_kind"] == 1
    key = f"{kind}_{args.precision}" + ("_tile" if tiled else "")
    tj = ROOT / "profiles" / "ncu_traffic.json"
    if tj.exists():
        t = json.loads(tj.read_text()).get(key)
        if t:
            traffic = t["dram_bytes_per_atom"] * nown
            traffic_src = t["source"]
    kern = {"lj_double": "k_pair_lj", "lj_mixed": "k_pair_lj_mixed+k_merge_ff",
            "lj_double_tile": "k_tile_lj2<EV,ONETYPE,ILP,threads,CTAs/SM,NVE> (fix nve fused on plain steps)",
            "lj_mixed_tile": "k_tile_lj2f<EV,ONETYPE,threads,CTAs/SM,NVE,GQ> (fix nve fused on plain steps)",
            "eam_double": "k_eam_rho_one+k_eam_embed+k_eam_force_one (+ rho/fp halo)",
            "eam_mixed": "k_eam_rho_mixed+k_eam_embed+k_eam_force_mixed+k_merge_ff (+ rho/fp halo)",
            "eam_double_tile": "k_tile_eam2_rho+k_tile_eam2_force (+ fp halo)",
            "eam_mixed_tile": "k_tile_eam_rho+k_eam_embed+k_tile_eam_force (+ rho/fp halo)"}
    note = ("the tile pair kernel is bound by shared-memory load wavefronts (one LDS.128 + one LDS.64 per "
            "list entry from scattered staged atoms) and then by the FP64 pipe; its DRAM traffic equals the "
            "algorithmic bytes; on plain steps the same launch also carries fix nve (final + initial "
            "integrate), whose 140 algorithmic B/atom are NOT counted in `achieved` (see fused_nve): "
            "DESIGN.md section 4" if tiled else
            "the flat pair kernels are bound by the L1TEX data pipe (one wavefront per gathered record, "
            "spline-table read and RED), not by DRAM: see DESIGN.md section 4")
    # fix nve rides in the pair launch on plain steps (k_tile_lj2<...,NVE>): the launch then does the
    # work of two phases of the reference; reported beside the pair-only figure, never inside it
    fused_calls = ph["initial_integrate"][1] < max(pair_calls // 2, 1) and tiled
    nve_bytes = 140.0
    roofline = {"bound": "hbm", "kernel": kern[key],
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "bytes_per_atom": bytes_atom, "pairs_stored_per_atom": ps,
                "list_entries_per_atom": st["list_entries"] / max(nown, 1),
                "us_per_launch": pair_s * 1e6, "flop_per_atom": flops_atom,
                "tflops": flops_atom * nown / pair_s / 1e12,
                "share_of_step": min(1.0, pair_s * args.steps / max(dev_ms * 1e-3, 1e-12)),
                "pairs_in_cutoff_per_atom": pc, "in_cutoff_fraction_measured": pc_ratio,
                "fused_nve": ({"nve_bytes_per_atom": nve_bytes,
                               "achieved_with_nve_bytes": (bytes_atom + nve_bytes) * nown / pair_s / 1e9,
                               "frac_with_nve_bytes": (bytes_atom + nve_bytes) * nown / pair_s / 1e9 / hbm_peak,
                               "note": "SURVEY 8(d) bytes of FixNVE final+initial integrate, done by the same "
                                       "launch; `achieved`/`frac` above count the pair bytes only"}
                              if fused_calls else None),
                "note": note}
    phases = {k: {"ms": round(t, 3), "calls": c} for k, (t, c) in ph.items() if c}
    # the limiting phases in one flat record (microseconds per timestep of the profiling pass)
    phase_us = {k: round(t * 1e3 / args.steps, 1) for k, (t, c) in ph.items() if c}

    if rank != 0:
        return
    res = {
        "metric": "atom-timesteps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64" if args.precision == "double" else "f32 pair math / f64 accumulate",
        "data": "synthetic",
        "config": {"workload": args.workload,
                   "workload_detail": f"{kind} {'LJ melt' if kind == 'lj' else 'Cu EAM'} "
                                      f"{natoms} atoms, fcc {cells[0]}x{cells[1]}x{cells[2]} cells, NVE, "
                                      f"skin {s['skin']}, neigh every {s['every']} delay {s['delay']} "
                                      f"check {'yes' if s['check'] else 'no'}",
                   "natoms": natoms, "proc_grid": list(grid), "precision": args.precision,
                   "l2": "working set (>= 100 B/atom x natoms) exceeds the 126 MB L2; no flush"
                         if natoms >= 2_000_000 else "working set fits L2 (small reference case)",
                   "rebuilds_in_timed_region": rebuilds,
                   "preconditioning_steps": precond,
                   "phase_us_per_step_serialised": phase_us,
                   "halo": {0: "none (one sub-domain, periodic self images)",
                            1: "NCCL send/recv between the 26 neighbour sub-domains",
                            2: "peer-memory stores over NVLink (CUDA IPC) + arrival/ack flags"}[
                                st["halo_transport"]],
                   "halo_overlap": (f"interior tiles ({st['tiles_interior']}) on a second stream beside the "
                                    f"halo and the boundary tiles ({st['tiles_boundary']})")
                                   if st["halo_overlap"] else "none",
                   "list": (f"bin tiles {st['tile'][0]}x{st['tile'][1]}x{st['tile'][2]} bins, 16-bit entries, "
                            f"<= {st['tile_stage_max']} atoms staged in shared memory per tile")
                           if st["list_kind"] == 1 else
                           f"flat int32 half list, {st['lanes_per_atom']} lanes per atom"},
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d / e2e_steps,
                "d2h_bytes_per_step": d2h / e2e_steps, "steps": e2e_steps,
                "includes": "pinned-host upload, Verlet setup (ghosts+list+forces), run, thermo "
                            "read-back, x/v/f download",
                "parts_ms_rank0": {k: round(v, 2) for k, v in e2e_parts.items()}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "phases": phases,
        "phases_note": f"per-phase CUDA events from a second untimed pass of {args.steps} steps "
                       f"({prof_ms / args.steps:.4f} ms/step with the phases serialised on one stream)",
        "wall_s_timed_region": wall,
        "host_setup_s": host_setup_s,
    }
    # the same kernel against the FP pipe it computes on (north_star: pipe utilisation of the pair
    # kernels): measured FMA peaks from the microbenchmark; `executed` counts what the kernel
    # really issues (18 FP64 instructions = 2 x 18 flop per list entry, both sides of every pair)
    pk = fp_peaks()
    if pk and kind == "lj":
        pipe = "fp64" if args.precision == "double" else "fp32"
        if pipe in pk:
            entries = st["list_entries"] / max(nown, 1)
            executed = entries * 36.0 * nown / pair_s / 1e12
            roofline["pipe"] = {"bound": pipe, "peak_tflops": pk[pipe], "peak_source":
                                "tools/microbench/fp_peak (FMA, 296 x 1024 threads, ILP 8) on this GPU",
                                "algorithmic_tflops": flops_atom * nown / pair_s / 1e12,
                                "algorithmic_frac": flops_atom * nown / pair_s / 1e12 / pk[pipe],
                                "executed_tflops": executed, "executed_frac": executed / pk[pipe]}
    if world == 1 and not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(kind, budget_s=20.0)
    if world == 1 and args.workload == "lj32m" and not args.no_also:
        e.close()
        res["also"] = also_block(args, local_rank)
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(res), flush=True)


# ----------------------------------------------------------------------------- reference
def lmp_script(kind, cells, warm, steps):
    from oracle import ref_harness as R
    if kind == "lj":
        return R.lj_input(run=warm, cells=cells) + f"\nrun {steps}\n"
    return R.eam_input(run=warm, cells=cells) + f"\nrun {steps}\n"


def run_lmp_ref(kind, cells, warm, steps, threads):
    from oracle import ref_harness as R
    if not R.EXE.exists():
        return None
    out = R.run_exe(lmp_script(kind, cells, warm, steps), threads=threads,
                    suffix="omp" if threads > 1 else None)
    loops = re.findall(r"Loop time of ([0-9.eE+-]+) on \d+ procs for (\d+) steps with (\d+) atoms", out)
    t, nst, nat = loops[-1]
    return float(t), int(nst), int(nat)


def cpu_baseline(kind, budget_s=20.0):
    """The reference's CPU path (oracle/_ref, kind "reference") on a bounded sample."""
    cores = os.cpu_count() or 1
    rate_guess = 1.2e6 * cores if kind == "lj" else 0.45e6 * cores
    steps = 100
    natoms = rate_guess * budget_s / steps
    cells = int(max(10, min(120, round((natoms / 4.0) ** (1 / 3)))))
    r = run_lmp_ref(kind, cells, 0, steps, cores)
    if r is None:
        return {"value": None, "unit": "atom-steps/s", "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/lmp_ref not built"}
    t, nst, nat = r
    return {"value": nat * nst / t, "unit": "atom-steps/s", "cores": cores, "kind": "reference",
            "sample": f"lmp_ref -sf omp -pk omp {cores}: {kind} {nat} atoms x {nst} steps "
                      f"(loop time {t:.2f} s, setup excluded as in finish.cpp)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, cells1, scaling = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    rate_guess = 1.2e6 * cores if kind == "lj" else 0.45e6 * cores
    natoms = rate_guess * 90.0 / max(total, 1)
    cells = int(max(10, min(cells1, round((natoms / 4.0) ** (1 / 3)))))
    r = run_lmp_ref(kind, cells, args.warmup, args.steps, cores)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lmp_ref is not built"}))
        return
    t, nst, nat = r
    val = nat * nst / t
    sample = (f"lmp_ref -sf omp -pk omp {cores}: {kind} {nat} atoms x {nst} steps after {args.warmup} "
              f"warm-up steps (bounded sample of {args.workload})")
    print(json.dumps({
        "impl": "reference", "metric": "atom-timesteps/s", "value": val, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / nst * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload,
                   "workload_detail": f"bounded sample of {args.workload}: {nat} atoms", "natoms": nat},
        "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lj32m", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="double", choices=["double", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the side measurements (eam2m, lj4m, lmp_b200)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_b200(args)
        finally:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
