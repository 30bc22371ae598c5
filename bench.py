"""bench.py -- preconditioner applies/s of the two-level RAS apply (BASELINE.json metric).

One "step" = one deflated two-level Schwarz::apply (mu = 1) over the whole
decomposition.  Workload ("config3 slice", weak scaling): 3-D 7-point Poisson,
one subdomain of m^3 cells (+ overlap 1) per GPU, GenEO-shaped coarse space of
nu = 20 vectors per subdomain; 1/2/4/8 GPUs = 1x1x1 / 2x1x1 / 2x2x1 / 2x2x2
subdomains.  `value` counts subdomain-applies per second
(= global applies/s x number of subdomains), so it aggregates over GPUs like the
contract asks; `applies_per_s` is the global figure.

  python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
  python bench.py --gpus 8 ...                               # without torchrun: spawns the 8 ranks itself (torch.distributed.run)
  python bench.py --impl reference ...                       # CPU arm (oracle port on host cores), SAME subdomain size as the CUDA arm

Both arms run the same workload: one subdomain of m^3 cells per GPU, m = --cells (default: 160, the largest exact FP64 factor
that fits one B200 with its numeric phase -- DESIGN.md section 4; 128 when the host cannot hold the CPU arm's 160^3 factor,
a rule both arms evaluate identically, see default_cells()).  The CPU arm never shrinks on its own.
  python bench.py --scalar z --cells 64                      # auxiliary: BASELINE config 5 shape (3-D Helmholtz, complex FP64,
                                                             # ORAS + plane-wave coarse space) through hpddm_b200z_*; not the headline
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "preconditioner applies/sec (FP64) 3-D Poisson, 1 subdomain/GPU, two-level RAS (deflated) + GenEO nu=20"
METRIC_E = "preconditioner block-applies/sec (FP64) 3-D linear elasticity (Q1, 3 dof/node), 1 subdomain/GPU, two-level RAS (deflated), nu=30, mu=4 (Block-GMRES shape)"
METRIC_Z = "preconditioner applies/sec (complex FP64) 3-D Helmholtz, 1 subdomain/GPU, two-level ORAS (deflated) + plane-wave coarse space"


def cosine_modes(dims, nu):
    """Analytic low-frequency Neumann modes of the cell-centred Laplacian on the
    subdomain box (DCT-II vectors): a synthetic coarse space of GenEO shape."""
    w, h, t = dims
    ks = sorted(((i, j, k) for i in range(6) for j in range(6) for k in range(6)), key=lambda q: (q[0] / w) ** 2 + (q[1] / h) ** 2 + (q[2] / t) ** 2)[:nu]
    x = (np.arange(w) + 0.5) / w
    y = (np.arange(h) + 0.5) / h
    z = (np.arange(t) + 0.5) / t
    Z = np.empty((w * h * t, nu), order="F")
    for c, (i, j, k) in enumerate(ks):
        v = np.cos(np.pi * k * z)[:, None, None] * np.cos(np.pi * j * y)[None, :, None] * np.cos(np.pi * i * x)[None, None, :]
        v = v.reshape(-1)
        Z[:, c] = v / np.linalg.norm(v)
    return Z


def elasticity_modes(part, Nn, nu):
    """GenEO-shaped coarse space for the elasticity workload: the 6 rigid-body modes of the subdomain (the kernel of its Neumann
    matrix, what GenEO finds first) modulated by low-frequency cosines of the box -- nu vectors, column-normalised."""
    from hpddm_b200.examples.generate import rigid_body_modes
    R = rigid_body_modes(part, Nn)
    w, h, t = part["dims"][:3]
    x = (np.arange(w) + 0.5) / w
    y = (np.arange(h) + 0.5) / h
    z = (np.arange(t) + 0.5) / t
    ks = sorted(((i, j, k) for i in range(3) for j in range(3) for k in range(3)), key=lambda q: q[0] ** 2 + q[1] ** 2 + q[2] ** 2)
    Z = np.empty((R.shape[0], nu), order="F")
    c = 0
    for (i, j, k) in ks:
        m = (np.cos(np.pi * k * z)[:, None, None] * np.cos(np.pi * j * y)[None, :, None] * np.cos(np.pi * i * x)[None, None, :]).reshape(-1)
        for r in range(6):
            if c == nu:
                break
            v = R[:, r] * np.repeat(m, 3)
            Z[:, c] = v / np.linalg.norm(v)
            c += 1
    return Z


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.samples, self.stop_flag = dev, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [int(s[0]) for s in self.samples if s and s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(st, nnz_a, nu, mu=1, s=8):
    """BASELINE.md section 4 traffic model, per subdomain, s = 8 bytes (16 for complex scalars)."""
    n, h = st["n"], st["halo"]
    b_trsv = st["factor_bytes"] * (2 if st["symmetric"] else 1) + st["index_bytes"] + 4 * s * n * mu
    b_halo = (2 * s + 8) * n * mu + 2 * s * h * mu
    b_z = s * n * nu + s * n * mu + 8 * n
    b_spmv = (s + 4) * nnz_a + 4 * (n + 1) + 3 * s * n * mu
    return dict(trsv=b_trsv, halo=b_halo, z=b_z, spmv=b_spmv, apply=2 * b_z + b_spmv + b_trsv + 3 * b_halo + 3 * s * n * mu)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from hpddm_b200.examples.generate import generate3d, split_grid_3d
    from hpddm_b200 import Decomposition

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = split_grid_3d(world)
    m = args.m
    N = tuple(g * m for g in grid)
    cplx = args.scalar == "z"
    elas = args.workload == "elasticity"
    if elas:   # BASELINE config 4 shape: Q1 elasticity, m nodes per subdomain edge (3 m^3 dofs + overlap), block of mu right-hand sides
        from hpddm_b200.examples.generate import generate_elasticity3d
        part = generate_elasticity3d(rank, world, Nn=N, overlap=1, mu=args.mu, grid=grid, assembly="local")
    elif cplx:
        from hpddm_b200.examples.generate import generate_helmholtz3d
        part = generate_helmholtz3d(rank, world, N=N, overlap=1, mu=1, grid=grid, k=args.wavenumber, nu=args.nu)
    else:
        part = generate3d(rank, world, N=N, overlap=1, mu=1, grid=grid)
    n = part["ndof"]
    tdtype = torch.complex128 if cplx else torch.float64
    deco = Decomposition(local, dtype=np.complex128 if cplx else np.float64)
    if world > 1:
        deco.comm_init_torch()
    s = deco.add(rank)
    s.initialize(part["Mat"], part["o"], part["mapping"])
    s.setGridHint(*part["dims"])
    deco.multiplicityScaling([part["d"]])
    t0 = time.time()
    if cplx:
        s.callNumfact(A=part["MatRobin"], method="oras")   # ORAS: impedance transmission conditions (Prcndtnr::OG)
    else:
        s.callNumfact()
    deco.synchronize()
    t_fact = time.time() - t0
    s.setVectors(elasticity_modes(part, N, args.nu) if elas else (part["Z"] if cplx else cosine_modes(part["dims"], args.nu)))
    deco.buildTwo()
    st = s.statistics()
    nnz_a = st["nnz_a"]
    by = algorithmic_bytes(st, nnz_a, args.nu, args.mu, s=16 if cplx else 8)
    stream = torch.cuda.ExternalStream(deco.stream, device=local)
    mu = args.mu
    x_dev = torch.rand(n * mu, dtype=tdtype, device="cuda")
    y_dev = torch.empty_like(x_dev)
    x_pin = torch.rand(n * mu, dtype=tdtype).pin_memory()
    y_pin = torch.empty(n * mu, dtype=tdtype).pin_memory()
    # ordinary (pageable) memory, what an unchanged Krylov driver allocates with new K[] (include/HPDDM_GMRES.hpp:45-50)
    x_pag = np.random.RandomState(5).rand(n * mu).astype(np.complex128 if cplx else np.float64)
    y_pag = np.empty_like(x_pag)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    sampler.start()
    l0 = deco.launches
    ms_dev = timed(lambda: deco.apply_device([x_dev], [y_dev], mu, "deflated"), args.steps, args.warmup)
    launches = (deco.launches - l0) // (args.steps + args.warmup) * args.steps
    # dominant kernel alone: the local triangular solves (forward + backward sweeps)
    from hpddm_b200 import capi
    ms_trsv = timed(lambda: deco.api.check(deco.api.sub_solve(s.h, x_dev.data_ptr(), y_dev.data_ptr(), mu, capi.DEVICE)), args.steps, args.warmup)
    ms_e2e = timed(lambda: deco.apply_host_inplace([x_pin], [y_pin], mu, "deflated"), args.steps, args.warmup)
    # the same call on pageable memory: (a) outside a start()/end() bracket -> plain pageable copies every time;
    # (b) inside one -> the two ranges are pinned in place on their 4th sighting (steady state of the work vector of a Krylov cycle)
    ms_pag_cold = timed(lambda: deco.apply_host_inplace([x_pag], [y_pag], mu, "deflated"), args.steps, args.warmup)
    x0_pag = y_pag.copy()   # (kept alive: start() writes the exchanged initial guess back into it)
    deco.api.check(deco.api.start(deco.ctx, capi.ptr_array([x_pag]), capi.ptr_array([x0_pag]), mu, capi.HOST))
    ms_pag = timed(lambda: deco.apply_host_inplace([x_pag], [y_pag], mu, "deflated"), args.steps, max(args.warmup, 5))
    hostreg = int(deco.api.ctx_hostreg_count(deco.ctx))
    deco.end()
    # where one apply goes (device pointers, each piece timed alone on the library's stream)
    w_dev = torch.rand(n * mu, dtype=tdtype, device="cuda")
    phases = {
        "halo_exchange_ms": timed(lambda: deco.api.check(deco.api.exchange(deco.ctx, capi.ptr_array([w_dev]), mu, 0, capi.DEVICE)), args.steps, args.warmup) / args.steps,
        "deflation_ms": timed(lambda: deco.api.check(deco.api.deflation(deco.ctx, capi.ptr_array([x_dev]), capi.ptr_array([y_dev]), mu, capi.DEVICE)), args.steps, args.warmup) / args.steps,
        "gmv_ms": timed(lambda: deco.api.check(deco.api.gmv(deco.ctx, capi.ptr_array([x_dev]), capi.ptr_array([y_dev]), mu, capi.DEVICE)), args.steps, args.warmup) / args.steps,
        "sptrsv_ms": ms_trsv / args.steps,
    }
    sampler.stop_flag = True
    sampler.join(timeout=2)
    peak, peak_src = peaks()
    traffic = None
    try:  # DRAM bytes of one solve measured by ncu for this configuration (profiles/), if captured
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["cells_complex" if cplx else "cells"].get(str(m), {}).get("dram_bytes") if args.mu == 1 else None
    except Exception:
        pass
    trsv_gbs = by["trsv"] / (ms_trsv / args.steps * 1e-3) / 1e9
    # the same fraction on strict nnz(L) bytes (structural non-zeros only, no panel padding): the conservative figure
    sz = 16 if cplx else 8
    strict = by["trsv"] - st["factor_bytes"] * (2 if st["symmetric"] else 1) + 2 * sz * st["nnz_factor"]
    trsv_gbs_strict = strict / (ms_trsv / args.steps * 1e-3) / 1e9
    if elas:
        wl, par = (f"config4 slice: 3-D Q1 linear elasticity {N[0]}x{N[1]}x{N[2]} nodes x 3 dof, {world} subdomain(s) of {m}^3 nodes + overlap 1, face x = 0 clamped by penalisation, "
                   f"two-level RAS deflated, nu={args.nu}, mu={args.mu}"), f"{grid[0]}x{grid[1]}x{grid[2]} subdomains, 1/GPU"
    elif cplx:
        wl, par = (f"config5 slice: 3-D Helmholtz k={args.wavenumber} {N[0]}x{N[1]}x{N[2]}, complex FP64, {world} subdomain(s) of {m}^3 cells + overlap 1, "
                   f"two-level ORAS deflated, {args.nu} plane waves, mu={args.mu}"), f"{grid[0]}x{grid[1]}x{grid[2]} subdomains, 1/GPU"
    else:
        wl, par = workload_name(m, world, args.nu, args.mu)
    out = {
        "metric": METRIC_E if elas else (METRIC_Z if cplx else METRIC), "value": world * args.steps / (ms_dev * 1e-3), "unit": "subdomain-applies/s", "applies_per_s": args.steps / (ms_dev * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": {"workload": wl, "parallelism": par, "cells": m, "nu": args.nu, "mu": args.mu},
        "factor": {"n_loc": n, "nnz_factor": st["nnz_factor"], "factor_gb": st["factor_bytes"] / 1e9, "levels": st["levels"], "fronts": st["fronts"],
                   "numfact_s": round(t_fact, 3), "symbolic_s": round(st["symbolic_seconds"], 3), "l2": "inputs (factor panels) larger than L2, no flush needed"},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "subdomain-applies/s", "h2d_bytes_per_step": (16 if cplx else 8) * n * mu, "d2h_bytes_per_step": (16 if cplx else 8) * n * mu,
                "ms_per_step": ms_e2e / args.steps, "host_memory": "pinned (cudaHostAlloc)",
                "pageable": {"value": world * args.steps / (ms_pag * 1e-3), "ms_per_step": ms_pag / args.steps, "ranges_pinned_in_place": hostreg,
                             "note": "ordinary numpy buffers inside a start()/end() bracket: pinned in place (cudaHostRegister) on their 4th sighting, i.e. during warm-up"},
                "pageable_unregistered": {"value": world * args.steps / (ms_pag_cold * 1e-3), "ms_per_step": ms_pag_cold / args.steps,
                                          "note": "ordinary numpy buffers outside start()/end(): plain pageable cudaMemcpyAsync each way"}},
        "phases": phases,
        "gpu_launches": int(launches),
        "roofline": {"kernel": "supernodal SpTRSV sweeps (k_fwd + k_bwd, all levels)", "bound": "hbm", "achieved": trsv_gbs, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": trsv_gbs / peak, "frac_strict_nnz": trsv_gbs_strict / peak, "traffic": traffic, "algorithmic_bytes_per_launch_set": by["trsv"], "ms": ms_trsv / args.steps,
                     "apply_gbs": by["apply"] / (ms_dev / args.steps * 1e-3) / 1e9},
        "clocks": sampler.summary(),
    }
    if args.krylov:
        # full solves with mu right-hand sides, Krylov data resident in HBM: pseudo-block GMRES (every column its own Krylov
        # space, one block apply per iteration) vs block GMRES (one block Krylov space; the method of BASELINE config 4)
        rs = np.random.RandomState(1234 + rank)
        bvec = [np.asfortranarray(rs.uniform(size=(n, mu)) + (1j * rs.uniform(size=(n, mu)) if cplx else 0.0))]
        bvec = deco.exchange(bvec, scaled=True)
        out["krylov"] = {}
        for name, fn in (("gmres", deco.solve), ("bgmres", deco.solve_bgmres)):
            fn(bvec, correction="deflated", max_it=2)   # warm-up (graph capture, allocations)
            deco.synchronize()
            t0 = time.time()
            it_k, _, res = fn(bvec, correction="deflated")
            out["krylov"][name] = {"seconds": time.time() - t0, "iterations": it_k, "max_rel_residual": float(np.max(res)), "rhs": mu}
        # GCRO-DR(40, 10) (the Krylov method of BASELINE config 5): two consecutive solves, the second one starts from the recycled pair
        # kept in the context (NOTE: device backend of this driver not yet measured on hardware when this line was written)
        deco.recycle_destroy()
        for tag in ("gcrodr_first_solve", "gcrodr_second_solve"):
            deco.synchronize()
            t0 = time.time()
            it_k, _, res = deco.solve_gcrodr(bvec, correction="deflated", restart=40, recycle=10)
            out["krylov"][tag] = {"seconds": time.time() - t0, "iterations": it_k, "max_rel_residual": float(np.max(res)), "rhs": mu, "recycled_dim": deco.recycle_dim()}
        deco.recycle_destroy()
    if cplx or elas:
        out["cpu_baseline"] = None   # the CPU arm (oracle/cpu_ras.cpp) is the scalar Poisson workload; these benches are auxiliary
    elif args.cpu_baseline and rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, steps=5, reuse=True)
    if rank == 0:
        print(json.dumps(out))
    deco.close()
    if world > 1:
        dist.destroy_process_group()


def host_memory_bytes():
    """Memory this process may use: MemAvailable capped by the cgroup limit."""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            avail = min(avail, int(lim)) if avail else int(lim)
    except Exception:
        pass
    return avail or 0


CPU_ARM_BYTES = {160: 129e9, 128: 52e9}   # cpu_ras_estimate_bytes(m, 16) of oracle/cpu_ras.cpp (factor + update stack + working fronts)


def default_cells():
    """Subdomain edge both arms use unless --cells / HPDDM_B200_BENCH_M says otherwise: 160 (largest fit of the CUDA arm),
    128 if the host could not hold the CPU arm's factorisation at 160 -- a function of the box only, so the `--impl reference`
    process and the CUDA process pick the same size without talking to each other."""
    if os.environ.get("HPDDM_B200_BENCH_M"):
        return int(os.environ["HPDDM_B200_BENCH_M"])
    return 160 if host_memory_bytes() >= 1.15 * CPU_ARM_BYTES[160] else 128


def workload_name(m, world, nu, mu):
    from hpddm_b200.examples.generate import split_grid_3d
    grid = split_grid_3d(world)
    N = tuple(g * m for g in grid)
    ov = [m + (1 if g > 1 else 0) for g in grid]
    return (f"config3 slice: 3-D Poisson {N[0]}x{N[1]}x{N[2]}, {world} subdomain(s) of {m}^3 cells + overlap 1, two-level RAS deflated, nu={nu}, mu={mu}",
            f"{grid[0]}x{grid[1]}x{grid[2]} subdomains, 1/GPU")


def _cache_path(m, nu, threads):
    import socket
    return os.path.join("/tmp", f"hpddm_b200_cpu_arm_{socket.gethostname()}_m{m}_nu{nu}_t{threads}.json")


def usable_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (cpu.max)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def _cpu_lib():
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_build", "libcpu_ras.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    L = C.CDLL(path)
    L.cpu_ras_create.restype = C.c_void_p
    L.cpu_ras_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.cpu_ras_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.cpu_ras_destroy.argtypes = [C.c_void_p]
    L.cpu_ras_nnz_factor.restype = C.c_long
    L.cpu_ras_nnz_factor.argtypes = [C.c_void_p]
    L.cpu_ras_factor_seconds.restype = C.c_double
    L.cpu_ras_factor_seconds.argtypes = [C.c_void_p]
    L.cpu_ras_estimate_bytes.restype = C.c_double
    L.cpu_ras_estimate_bytes.argtypes = [C.c_int, C.c_int]
    return L


def cpu_baseline(args, m=None, steps=None, reuse=False, ranks=1):
    """CPU arm = oracle/cpu_ras.cpp: the reference's path restated for the host cores (supernodal
    sparse Cholesky + dtrsv/dgemv supernodal solves as CHOLMOD/MUMPS do, dgemv projections, OpenMP
    CSR SpMV), all host threads, ONE subdomain of exactly the CUDA arm's size (m^3 cells; never reduced here).
    The "bounded sample" is the number of applies, not the size.  reuse=True (the in-line leg of the CUDA arm):
    take the measurement the `--impl reference` run left on this box for the same (m, nu, threads) instead of
    factoring the same matrix on the host a second time."""
    L = _cpu_lib()
    # scipy's OpenBLAS is built for at most 128 threads *including callers*: stay well below
    avail = usable_cpus()
    # the host cores one of `ranks` concurrent subdomain processes would get, but never fewer than 16 (or all there are): the CPU
    # factorisation of the largest-fit subdomain must stay within minutes on a box that exposes few cores
    threads = max(1, min(64, max(avail // max(1, ranks), min(avail, 16))))
    m = m or args.m
    cache = _cache_path(m, args.nu, threads)
    if reuse and os.path.exists(cache) and time.time() - os.path.getmtime(cache) < 6 * 3600:
        cb = json.load(open(cache))
        cb["sample"] += f" [measured {int(time.time() - os.path.getmtime(cache))} s earlier on this box by `bench.py --impl reference`, not repeated]"
        return cb
    need = L.cpu_ras_estimate_bytes(m, threads)
    have = host_memory_bytes()
    if have and need > 0.95 * have:
        raise SystemExit(f"bench.py: the CPU arm needs {need / 1e9:.0f} GB of host memory for a {m}^3 subdomain, the box offers {have / 1e9:.0f} GB; "
                         f"run both arms with a smaller --cells (the arms never use different sizes)")
    Z = cosine_modes((m, m, m), args.nu)
    t0 = time.time()
    h = L.cpu_ras_create(m, args.nu, Z.ctypes.data, threads)
    t_create = time.time() - t0
    n = m ** 3
    x = np.random.RandomState(0).uniform(size=n)
    y = np.empty(n)
    for _ in range(2):
        L.cpu_ras_apply(h, x.ctypes.data, y.ctypes.data)
    k = steps or max(3, min(30, args.steps))
    t0 = time.time()
    for _ in range(k):
        L.cpu_ras_apply(h, x.ctypes.data, y.ctypes.data)
    dt = (time.time() - t0) / k
    nnz = L.cpu_ras_nnz_factor(h)
    tf = L.cpu_ras_factor_seconds(h)
    L.cpu_ras_destroy(h)
    cb = {"value": 1.0 / dt, "unit": "subdomain-applies/s", "cores": threads, "host_cpus": os.cpu_count(), "affinity_cpus": avail, "kind": "port", "ms_per_apply": dt * 1e3,
          "cells": m, "nnz_factor": int(nnz), "max_rss_gb": round(__import__("resource").getrusage(__import__("resource").RUSAGE_SELF).ru_maxrss / 1e6, 1), "effective_gbs": (2 * 8 * nnz + 8 * n * (2 * args.nu + 12 + 7 * 1.5)) / dt / 1e9,
          "sample": f"oracle/cpu_ras.cpp (supernodal Cholesky + BLAS-2 supernodal solves, OpenMP x {threads} threads), one subdomain of {m}^3 cells "
                    f"(the CUDA arm's size), nu={args.nu}, {k} timed applies after 2 warm-up, nnz(L)={nnz:.4g}, CPU analysis+numfact {tf:.1f}s (setup {t_create:.1f}s)"}
    try:
        json.dump(cb, open(cache, "w"))
    except Exception:
        pass
    return cb


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    if rank != 0:
        return
    # N subdomains = N reference processes sharing the host cores (MPI ranks x OpenMP threads): ONE of them is timed with
    # cores/N threads and the job is credited N times its rate -- optimistic for the CPU (the N - 1 concurrent ranks would
    # compete for memory bandwidth), never for this repo.  The measurement for a given (m, nu, threads) is taken once per box.
    cb = cpu_baseline(args, steps=args.steps, reuse=world > 1, ranks=world)
    wl, par = workload_name(args.m, world, args.nu, args.mu)
    if world > 1:
        cb = dict(cb, value=world * cb["value"], one_subdomain_value=cb["value"],
                  sample=cb["sample"] + f"; job value = {world} x the one-subdomain rate ({world} concurrent ranks x {cb['cores']} threads assumed interference-free)")
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": cb["unit"], "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * world / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": wl, "parallelism": par, "cells": args.m, "nu": args.nu, "mu": args.mu},
           "factor": {"nnz_factor": cb["nnz_factor"]}, "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def respawn_under_torchrun(args):
    """`python bench.py --gpus N` without a launcher: start the N ranks ourselves (one per GPU, as the driver does)."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    raise SystemExit(subprocess.call(cmd))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", dest="m", type=int, default=0, help="cells per subdomain edge, both arms (0 = default_cells(): 160, or 128 on a small host)")
    ap.add_argument("--nu", "--nvec", dest="nu", type=int, default=20, help="deflation vectors per subdomain (use --nvec under torchrun: its parser claims --nu)")
    ap.add_argument("--rhs", dest="mu", type=int, default=1, help="right-hand sides per apply (block methods)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--workload", default="poisson", choices=["poisson", "elasticity"],
                    help="poisson: the headline (config 3 slice); elasticity: config 4 shape (Q1 3 dof/node, use --rhs 4 --nu 30; --cells = nodes per subdomain edge)")
    ap.add_argument("--scalar", default="d", choices=["d", "z"], help="d: real FP64 Poisson (headline); z: complex FP64 Helmholtz / ORAS (config 5 shape, auxiliary)")
    ap.add_argument("--wavenumber", type=float, default=2.0)
    ap.add_argument("--krylov", action="store_true", help="also time a full GMRES solve: device-resident driver vs host-driven loop over the C ABI")
    args = ap.parse_args()
    if not args.m:
        args.m = 64 if args.workload == "elasticity" else default_cells()
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but the launcher started WORLD_SIZE={os.environ['WORLD_SIZE']} ranks")
    if "WORLD_SIZE" not in os.environ and args.gpus > 1 and args.impl != "reference":
        respawn_under_torchrun(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
