#!/usr/bin/env python
"""bench.py -- ms per KNP-EMI timestep (assembly + solve) on synthetic meshes, with the HBM roofline of the
dominant kernels and the CPU oracle timed beside it.

    python bench.py --gpus 1 --steps 5 --warmup 3                    # our arm, BASELINE config C3 (2D N=2048, 64 cells)
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 3   # CPU restatement of the reference path
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # weak scaling: per-GPU mesh fixed

A "step" is one pass of the reference's time-loop body (KNPEMIx_solver.py:365-468): gate ODE, assembly of A and b,
GMRES + AMG solve, field update, on one batch of synthetic input.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms per KNP-EMI timestep (assembly+solve)"


WORKLOADS = {
    # name: (config file, N in the file, cells_per_dim in the file, gdim, models)
    "c3": ("c3_square2048_cells64.yaml", 2048, 8, 2, ("NeuronalCT", "HH", "ATP")),
    "c4": ("c4_cube120_cells64_passive.yaml", 120, 4, 3, ("Passive",)),
    "c5": ("c5_cube256_tissue512_hh.yaml", 256, 8, 3, ("NeuronalCT", "HH", "ATP")),
}


def workload_yaml(kb, workload, n, cells_per_dim=None, rtol=None):
    fname, n0, m0, gdim, _ = WORKLOADS[workload]
    m = cells_per_dim or m0
    txt = open(os.path.join(os.path.dirname(kb.__file__), "configs", fname)).read()
    txt = txt.replace(f"N: {n0}", f"N: {n}").replace(f"cells_per_dim: {m0}", f"cells_per_dim: {m}")
    txt = txt.replace(f"!range [2, {2 + m0 ** gdim}]", f"!range [2, {2 + m ** gdim}]")
    if rtol is not None:
        txt = txt.replace("ksp_rtol: 1.0e-9", f"ksp_rtol: {rtol:.1e}")
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    f.write(txt)
    f.close()
    return f.name


def build_problem(kb, workload, n, local, cells_per_dim=None, rtol=None, restart=30, amg_form=None):
    """ProblemKNPEMI + SolverKNPEMI of one workload through the reference-facing API, preconditioner set up, t = 0."""
    cfg = workload_yaml(kb, workload, n, cells_per_dim, rtol)
    t = [time.time()]
    p = kb.ProblemKNPEMI(cfg, verbose=False, device=local)
    os.unlink(cfg)
    p.set_initial_conditions()
    ctor = {"NeuronalCT": kb.NeuronalCotransporters, "HH": kb.HodgkinHuxley, "ATP": kb.ATPPump, "Passive": kb.PassiveModel}
    p.init_ionic_models([ctor[nm](p) for nm in WORKLOADS[workload][4]])
    t.append(time.time())
    p.setup_variational_form()
    t.append(time.time())
    p.solver_config["view_ksp"] = False
    s = kb.SolverKNPEMI(p, solver_config=p.solver_config)
    s.gmres_restart = restart
    if amg_form:
        s.amg_form = amg_form
    s.setup_solver()
    p.setup_preconditioner(True)
    t.append(time.time())
    s.ctx.pc_setup(s.opts)
    s.ctx.set_time(0.0, 0)
    t.append(time.time())
    if p.comm.rank == 0 and s.ctx.n_rows > 1000000:
        print(f"bench.py: setup of {workload} N={n}: mesh + host mirror {t[1] - t[0]:.1f} s, device context (topology, CSR pattern, halo) "
              f"{t[2] - t[1]:.1f} s, P assembly {t[3] - t[2]:.1f} s, preconditioner setup (AMG hierarchies) {t[4] - t[3]:.1f} s",
              file=sys.stderr, flush=True)
    return p, s


def field_norms(kb, p):
    """Global L2 norms of the eight fields (int u^2 over the field's own subdomain), all-reduced over the ranks."""
    it, et = list(p.intra_tags), [p.extra_tag[0]]
    return [math.sqrt(p.comm.allreduce(p.l2_norm_squared(p.wh[s][f], it if s == 0 else et), op=kb.MPI.SUM))
            for s in range(2) for f in range(4)]


def parity_check(kb, local):
    """Rank-count independence of the distributed path, recorded in every bench line: two small fixed problems (BASELINE C3
    and C4 in miniature) are stepped on ALL ranks of this run and the L2 norms of the eight fields are compared with the
    values the CPU oracle produced for the same problems (tests/golden/parity_small.json, committed with its generating
    script; the oracle itself is not imported here).  Tolerance 1e-8 relative (north star), potentials relative to the
    potential scale."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "parity_small.json")))
    out = {"tolerance": 1e-8, "ok": True, "cases": {}}
    for name, wl in (("c3_mini", "c3"), ("c4_mini", "c4")):
        g = gold[name]
        p, s = build_problem(kb, wl, g["N"], local, cells_per_dim=g["cells_per_dim"], rtol=1e-12)
        its = [int(s.ctx.step(s.opts).iterations) for _ in range(g["steps"])]
        p._mark_device_newer()
        got, ref = field_norms(kb, p), g["norms"]
        scale = list(ref)
        scale[3] = scale[7] = max(ref[3], ref[7])
        err = max(abs(a - b) / sc for a, b, sc in zip(got, ref, scale))
        out["cases"][name] = {"max_rel_err": err, "iterations": its, "ranks": p.comm.size}
        out["ok"] = out["ok"] and bool(err < out["tolerance"])
        s.ctx.close()
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.index, self.proc = [], index, None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_workload(workload, n, restart=30):
    """The C++/OpenMP CPU baseline (oracle/cpu.py -> oracle/libknpemi_cpu.so: restatement of the reference's time-loop body
    with the same GMRES + Schur/SA-AMG algorithm, all host cores) set up on the FULL workload through the same config file,
    mesh generator, initial state and model list as the GPU arm (the host mirror of the reference classes parses the YAML;
    no device context is created).  Returns (baseline, description)."""
    import cgx_b200 as kb
    from oracle.cpu import CpuBaseline
    from oracle.knpemi import OracleParams
    _, _, _, gdim, model_names = WORKLOADS[workload]
    cfg = workload_yaml(kb, workload, n)
    pr = kb.ProblemKNPEMI(cfg, verbose=False)
    os.unlink(cfg)
    pr.set_initial_conditions()
    ctor = {"NeuronalCT": kb.NeuronalCotransporters, "HH": kb.HodgkinHuxley, "ATP": kb.ATPPump, "Passive": kb.PassiveModel}
    pr.init_ionic_models([ctor[nm](pr) for nm in model_names])
    mesh = pr.mesh
    region = None
    if pr.stimulus_region and pr.multiple_stimulus_directions:
        region = tuple((int(d), float(r[0]), float(r[1])) for d, r in zip(pr.stimulus_region_directions, pr.stimulus_region_range))
    elif pr.stimulus_region:
        region = (int(pr.stimulus_region_direction), float(pr.stimulus_region_range[0]), float(pr.stimulus_region_range[1]))
    ions = pr.ion_list
    p = OracleParams(dt=float(pr.dt.value), T=pr.T.value, F=pr.F.value, R=pr.R.value, C_M=pr.C_M.value,
                     z=tuple(i["z"].value for i in ions), D=tuple(i["Di"].value for i in ions), phi_rest=pr.phi_rest.value,
                     g_Na_bar=pr.g_Na_bar.value, g_K_bar=pr.g_K_bar.value, g_leak=tuple(i["g_leak"].value for i in ions),
                     g_leak_g=tuple(i["g_leak_g"].value for i in ions), g_syn_bar=pr.g_syn_bar.value, a_syn=pr.a_syn.value,
                     T_stim=pr.T_stim.value, scale_stimulus=bool(pr.scale_stimulus), intra_tags=tuple(pr.intra_tags),
                     extra_tag=pr.extra_tag[0], membrane_tags=tuple(pr.gamma_tags), stimulus_tags=tuple(pr.stimulus_tags),
                     stimulus_region=region, c_e_init=tuple(i["ke_init"].value for i in ions))
    cb = CpuBaseline(gdim, mesh.x, mesh.cells, mesh.cell_tags, mesh.mf_verts, mesh.mf_tags, p, [(nm, None) for nm in model_names],
                     restart=restart)
    vi, ve = cb.S
    u = np.concatenate([pr.wh[0][f]._data[vi] for f in range(4)] + [pr.wh[1][f]._data[ve] for f in range(4)])
    if pr.gating_variables:
        gates = np.stack([pr.n._data[cb.mverts], pr.m._data[cb.mverts], pr.h._data[cb.mverts]])
    else:
        gates = np.zeros((3, cb.n_mv))
    cb.set_state(u, gates)
    return cb, dict(rows=cb.n, nnz=cb.nnz, cells=int(mesh.cells.shape[0]))


def cpu_run(workload, n, warmup, steps, restart=30):
    t0 = time.time()
    cb, info = cpu_workload(workload, n, restart)
    cb.pc_setup()
    info["setup_s"] = time.time() - t0
    its, tot, asm = [], [], []
    for i in range(warmup + steps):
        k, ms = cb.step(1e-9)
        if i >= warmup:
            its.append(k)
            tot.append(ms["total"])
            asm.append(ms["assembly"])
    info.update(ms=float(np.mean(tot)), assembly_ms=float(np.mean(asm)), iterations=its, threads=cb.threads)
    cb.close()
    return info


def weak_n(base_n, world):
    return base_n if world == 1 else int(round(base_n * math.sqrt(world) / 8)) * 8


def run_reference(args):
    """Reference arm: the CPU implementation of the path (C++/OpenMP port of the oracle; the DOLFINx/PETSc stack itself is
    not installable here) on ALL host cores, on the same workload at the same size as the GPU arm -- no scaling factor."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    base_n = args.size if args.size else WORKLOADS[wl][1]
    n = weak_n(base_n, args.gpus) if wl == "c3" else base_n
    os.environ.pop("OMP_NUM_THREADS", None) if os.environ.get("OMP_NUM_THREADS") == "1" else None
    t0 = time.time()
    r = cpu_run(wl, n, args.warmup, args.steps, args.restart)
    line = {"metric": METRIC, "value": r["ms"], "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms"], "higher_is_better": False, "scaling": "weak" if wl == "c3" else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(wl, n, r["cells"]) + f", GMRES({args.restart}) + charge-conservation Schur PC (SA-AMG blocks) rtol 1e-9"
                                   + (", ICs perturbed as SURVEY 8(d)" if wl == "c3" else ""),
                       "dofs": r["rows"], "nnz": r["nnz"], "cells": r["cells"], "iterations_per_step": r["iterations"],
                       "timed_step_indices": [args.warmup + 1, args.warmup + args.steps], "setup_s": r["setup_s"]},
            "phases_ms": {"assembly": r["assembly_ms"], "solve": r["ms"] - r["assembly_ms"]},
            "cpu_baseline": {"value": r["ms"], "unit": "ms", "cores": r["threads"], "kind": "port", "extrapolated": False,
                             "sample": f"C++/OpenMP port of the CPU oracle (restatement of the DOLFINx/PETSc time-loop body; the stack itself is "
                                       f"not installable) on the FULL workload ({r['rows']} rows), {r['threads']} threads of {os.cpu_count()} host "
                                       f"cores, steps {args.warmup + 1}..{args.warmup + args.steps}, measured, not scaled"},
            "e2e": {"value": r["ms"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def workload_name(wl, n, n_cells):
    return {"c3": f"BASELINE C3: synthetic 2D tissue block N={n} (8x8 cells), Na/K/Cl + HH+ATP+KCC2",
            "c4": f"BASELINE C4: synthetic 3D tissue block N={n} (4x4x4 cells, {n_cells} tetrahedra), Na/K/Cl + passive membrane",
            "c5": f"BASELINE C5: synthetic dense-tissue-like 3D mesh N={n} (8x8x8 plate-stack cells, {n_cells} tetrahedra), "
                  f"Na/K/Cl + HH+ATP+KCC2, stimulus region"}[wl]


# ----------------------------------------------------------------------------------------------- GPU arm
def timed_steps(ctx, opts, steps, barrier, comm, kb, world):
    """K steps of the time loop, device-timed with the CUDA events knp_step records on its launching stream, bracketed by
    barrier + synchronize; returns per-step means (max over ranks) and the iteration counts."""
    barrier()
    tot = asm = sol = 0.0
    its = []
    wall0 = time.perf_counter()
    for _ in range(steps):
        info = ctx.step(opts)
        tm = ctx.last_timings()
        tot += tm["total"]
        asm += tm["gate"] + tm["facet"] + tm["rows"]
        sol += tm["solve"]
        its.append(int(info.iterations))
    barrier()
    wall = (time.perf_counter() - wall0) * 1e3 / steps
    ms_dev = comm.allreduce(tot / steps, op=kb.MPI.MAX)
    ms_wall = comm.allreduce(wall, op=kb.MPI.MAX)
    return dict(ms=max(ms_dev, ms_wall) if world > 1 else ms_dev, ms_device=ms_dev, ms_wall=ms_wall,
                assembly_ms=comm.allreduce(asm / steps, op=kb.MPI.MAX), solve_ms=comm.allreduce(sol / steps, op=kb.MPI.MAX),
                iterations=its, local_total_ms=tot / steps)


def run_ours(args):
    import torch
    import cgx_b200 as kb
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        # torchrun pins OMP to 1 thread; the host-side setup (topology, AMG hierarchy) is OpenMP code
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = args.workload
    gdim = WORKLOADS[wl][3]
    base_n = args.size if args.size else WORKLOADS[wl][1]
    if wl == "c3":
        # weak scaling: cells per GPU fixed -> N grows with sqrt(world), rounded to the 8 x 8 cell array
        n = weak_n(base_n, world)
        scaling = "weak"
    else:
        n, scaling = base_n, "strong"                  # BASELINE C4: one fixed mesh sharded over the GPUs
    t_setup = time.time()
    p, s = build_problem(kb, wl, n, local, restart=args.restart, amg_form=args.amg_form)
    ctx = s.ctx
    t_setup = time.time() - t_setup
    comm = p.comm

    def barrier():
        torch.cuda.synchronize()
        comm.Barrier()

    for _ in range(args.warmup):
        info = ctx.step(s.opts)
        if rank == 0:
            print(f"bench.py: warm-up step, {info.iterations} GMRES iterations", file=sys.stderr, flush=True)
    sampler = ClockSampler(local)
    u_start, g_start = ctx.get_state()          # the e2e leg below repeats exactly these K steps through host buffers
    t_start, i_start = ctx.get_time()
    barrier()
    sampler.start()
    l0 = kb.lib.launch_count()
    print(f"bench.py: {l0} kernel launches before the timed region (ncu -s)", file=sys.stderr, flush=True)
    T = timed_steps(ctx, s.opts, args.steps, barrier, comm, kb, world)
    launches = kb.lib.launch_count() - l0
    clocks = sampler.stop()
    ms, its = T["ms"], T["iterations"]

    # ---- end to end through the host-buffer C-ABI call (H2D of the state + step + D2H of the result every step)
    nst = ctx.n_cols
    u_pin = torch.empty(nst, dtype=torch.float64).pin_memory()
    g_pin = torch.empty(3 * ctx.n_mverts, dtype=torch.float64).pin_memory()
    u_pin.numpy()[:] = u_start
    g_pin.numpy()[:] = g_start.ravel()
    un, gn = u_pin.numpy(), g_pin.numpy()
    ctx.set_time(t_start, i_start)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.step_host(un, gn, s.opts)
    barrier()
    e2e_ms = comm.allreduce((time.perf_counter() - e0) * 1e3 / args.steps, op=kb.MPI.MAX)
    io_bytes = (nst + 3 * ctx.n_mverts) * 8

    # ---- roofline of the dominant kernels, measured live with CUDA events on the launching stream
    st = torch.cuda.Stream()
    sp = st.cuda_stream
    x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda")
    y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            fn()
        b.record(st)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    t_spmv = timeit(lambda: ctx.spmv(x.data_ptr(), y.data_ptr(), stream=sp), 20)
    t_asm = timeit(lambda: ctx.assemble(1e-4, stream=sp), 10)
    comm.Barrier()
    t_pc = timeit(lambda: ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp), 10)
    m = p.mesh
    nvert, ncell = m.x.shape[0], m.cells.shape[0]
    B_spmv = 12 * ctx.nnz + 20 * ctx.n_rows
    B_asm = (8 * ctx.nnz + 16 * ctx.n_rows + 8 * m.gdim * nvert + (4 * (m.gdim + 1) + 4) * ncell
             + 32 * ctx.n_mverts + 16 * ctx.sizes.n_mfacets)
    B_pc = ctx.pc_bytes() if hasattr(ctx, "pc_bytes") else None
    peak, peak_src = peaks()
    mean_its = float(np.mean(its))
    step_local = T["local_total_ms"]
    # per step: (its + 1) A-SpMVs (one true residual per restart cycle) and (its + 2) preconditioner applications
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram__bytes_read+write per launch from the ncu --set full captures
    if os.path.exists(tpath) and world == 1:
        traffic = json.load(open(tpath))
    kernels = {
        "assembly (facet_kernel + rows_edge_kernel)": {"ms": t_asm, "algorithmic_bytes": B_asm, "GB/s": B_asm / t_asm / 1e6,
                                                       "frac": B_asm / t_asm / 1e6 / peak, "share_of_step": t_asm / step_local,
                                                       "traffic": traffic.get(f"assembly@{wl}:N={n}")},
        "spmv A (spmv_stream_kernel<EPI_SET>, plain CSR, TMA-staged)": {
            "ms": t_spmv, "algorithmic_bytes": B_spmv, "GB/s": B_spmv / t_spmv / 1e6, "frac": B_spmv / t_spmv / 1e6 / peak,
            "share_of_step": (mean_its + 1) * t_spmv / step_local, "traffic": traffic.get(f"spmv_stream_kernel<0>@{wl}:N={n}")},
        "pc_apply (Schur preconditioner: two SA-AMG cycles + mass SpMV)": {
            "ms": t_pc, "algorithmic_bytes": B_pc, "GB/s": (B_pc / t_pc / 1e6) if B_pc else None,
            "frac": (B_pc / t_pc / 1e6 / peak) if B_pc else None, "share_of_step": (mean_its + 2) * t_pc / step_local,
            "traffic": traffic.get(f"pc_apply@{wl}:N={n}")},
    }
    # the roofline key names the kernel group with the largest share of the step
    top = max((k for k in kernels if kernels[k].get("frac") is not None), key=lambda k: kernels[k]["share_of_step"])
    kt = kernels[top]
    roof = {"bound": "hbm", "kernel": top, "achieved": kt["GB/s"], "peak": peak, "unit": "GB/s", "frac": kt["frac"],
            "traffic": kt.get("traffic"), "peak_source": peak_src, "algorithmic_bytes": kt["algorithmic_bytes"],
            "ms_per_launch": kt["ms"], "share_of_step": kt["share_of_step"]}
    dofs_global = int(comm.allreduce(float(ctx.n_rows), op=kb.MPI.SUM))
    nnz_global = int(comm.allreduce(float(ctx.nnz), op=kb.MPI.SUM))
    n_cells_global = int(p.global_mesh_info["n_cells"])
    names = {wl: workload_name(wl, n, n_cells_global)}
    ctx.close()
    del p, s, ctx, x, y

    # ---- BASELINE C4 (3D, ~10 M tetrahedra, passive membrane) sharded over the same GPUs: strong scaling, reported beside
    #      the headline so that every driver-run record holds it (skipped when it IS the headline)
    c4 = None
    if wl != "c4" and not args.skip_c4:
        tc = time.time()
        p4, s4 = build_problem(kb, "c4", args.c4_size, local, restart=args.restart, amg_form=args.amg_form)
        tc = time.time() - tc
        for _ in range(3):
            s4.ctx.step(s4.opts)
        T4 = timed_steps(s4.ctx, s4.opts, max(3, min(args.steps, 5)), barrier, comm, kb, world)
        c4 = {"workload": f"BASELINE C4: 3D tissue block N={args.c4_size} (4x4x4 cells), passive membrane, sharded over {world} GPU(s)",
              "scaling": "strong", "ms_per_step": T4["ms"], "assembly_ms": T4["assembly_ms"], "solve_ms": T4["solve_ms"],
              "iterations_per_step": T4["iterations"],
              "ms_per_iteration": T4["solve_ms"] / max(1.0, float(np.mean(T4["iterations"]))),
              "dofs": int(comm.allreduce(float(s4.ctx.n_rows), op=kb.MPI.SUM)),
              "nnz": int(comm.allreduce(float(s4.ctx.nnz), op=kb.MPI.SUM)), "setup_s": tc}
        s4.ctx.close()
        del p4, s4

    par = parity_check(kb, local) if not args.skip_parity else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args)
    if rank == 0:
        line = {"metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": False, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": names[wl] + f", GMRES({args.restart}) + charge-conservation Schur PC (SA-AMG blocks) rtol 1e-9"
                                       + (", ICs perturbed as SURVEY 8(d)" if wl != "c4" else ""),
                           "dofs": dofs_global, "nnz": nnz_global, "cells": n_cells_global,
                           "iterations_per_step": its, "timed_step_indices": [args.warmup + 1, args.warmup + args.steps],
                           "l2_policy": "inputs (A: %.1f GB per GPU) larger than L2" % (12 * nnz_global / world / 1e9),
                           "setup_s": t_setup},
                "phases_ms": {"assembly": T["assembly_ms"], "solve": T["solve_ms"]},
                "ms_per_iteration": T["solve_ms"] / max(1.0, mean_its),
                "clocks": clocks,
                "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes},
                "gpu_launches": int(launches),
                "roofline": roof, "kernels": kernels}
        if c4:
            line["c4"] = c4
        if par:
            line["parity_check"] = par
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if par and not par["ok"]:
        raise SystemExit(f"bench.py: parity check failed: {par}")


def cpu_baseline_leg(args):
    """cpu_baseline of the GPU arm's line (rank 0, N = 1): the C++/OpenMP CPU port on the same workload at the same size, all
    host cores, bounded to 2 timed steps after the same warm-up steps (about 10-30 s of CPU work)."""
    wl = args.workload
    n = args.size if args.size else WORKLOADS[wl][1]
    k = min(args.steps, 2)
    r = cpu_run(wl, n, args.warmup, k, args.restart)
    return {"value": r["ms"], "unit": "ms", "cores": r["threads"], "kind": "port", "extrapolated": False,
            "iterations": r["iterations"], "assembly_ms": r["assembly_ms"], "setup_s": r["setup_s"],
            "sample": f"C++/OpenMP port of the CPU oracle (restatement of the DOLFINx/PETSc time-loop body, same GMRES + Schur/SA-AMG "
                      f"algorithm) on the FULL workload ({r['rows']} rows), {r['threads']} threads of {os.cpu_count()} host cores, "
                      f"steps {args.warmup + 1}..{args.warmup + k}, measured, not scaled"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS),
                    help="headline workload: c3 = BASELINE configs[2] (weak-scaled per GPU), c4 = configs[3] (strong), "
                         "c5 = configs[4] (strong, sized for 8 GPUs)")
    ap.add_argument("--size", type=int, default=0, help="grid squares per side (default: the BASELINE size of the workload)")
    ap.add_argument("--c4-size", type=int, default=120, help="N of the C4 section reported beside the headline")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--restart", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--amg-form", default=None, choices=["schur", "block_jacobi"], help="override SolverKNPEMI.amg_form")
    args = ap.parse_args()
    if args.warmup < 3:
        print("bench.py: warm-up raised to 3 (timing rules)", file=sys.stderr)
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
