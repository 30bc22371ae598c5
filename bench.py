"""bench.py -- preconditioner applies/s of the two-level RAS apply (BASELINE.json metric).

One "step" = one deflated two-level Schwarz::apply (mu = 1) over the whole
decomposition.  Workload ("config3 slice", weak scaling): 3-D 7-point Poisson,
one subdomain of m^3 cells (+ overlap 1) per GPU, GenEO-shaped coarse space of
nu = 20 vectors per subdomain; 1/2/4/8 GPUs = 1x1x1 / 2x1x1 / 2x2x1 / 2x2x2
subdomains.  `value` counts subdomain-applies per second
(= global applies/s x number of subdomains), so it aggregates over GPUs like the
contract asks; `applies_per_s` is the global figure.

  python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
  python bench.py --impl reference ...                       # CPU arm (oracle port on host cores)
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
    if cplx:
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
    s.setVectors(part["Z"] if cplx else cosine_modes(part["dims"], args.nu))
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
    sampler.stop_flag = True
    sampler.join(timeout=2)
    peak, peak_src = peaks()
    traffic = None
    try:  # DRAM bytes of one solve measured by ncu for this configuration (profiles/), if captured
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["cells_complex" if cplx else "cells"].get(str(m), {}).get("dram_bytes") if args.mu == 1 else None
    except Exception:
        pass
    trsv_gbs = by["trsv"] / (ms_trsv / args.steps * 1e-3) / 1e9
    out = {
        "metric": METRIC_Z if cplx else METRIC, "value": world * args.steps / (ms_dev * 1e-3), "unit": "subdomain-applies/s", "applies_per_s": args.steps / (ms_dev * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": {"workload": (f"config5 slice: 3-D Helmholtz k={args.wavenumber} {N[0]}x{N[1]}x{N[2]}, complex FP64, {world} subdomain(s) of {m}^3 cells + overlap 1 (n_loc={n}), "
                                f"two-level ORAS deflated, {args.nu} plane waves, mu={args.mu}") if cplx else
                               f"config3 slice: 3-D Poisson {N[0]}x{N[1]}x{N[2]}, {world} subdomain(s) of {m}^3 cells + overlap 1 (n_loc={n}), two-level RAS deflated, nu={args.nu}, mu={args.mu}",
                   "parallelism": f"{grid[0]}x{grid[1]}x{grid[2]} subdomains, 1/GPU", "l2": "inputs (factor panels) larger than L2, no flush needed",
                   "nnz_factor": st["nnz_factor"], "factor_gb": st["factor_bytes"] / 1e9, "levels": st["levels"], "fronts": st["fronts"],
                   "numfact_s": round(t_fact, 3), "symbolic_s": round(st["symbolic_seconds"], 3)},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "subdomain-applies/s", "h2d_bytes_per_step": (16 if cplx else 8) * n * mu, "d2h_bytes_per_step": (16 if cplx else 8) * n * mu,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "supernodal SpTRSV sweeps (k_fwd + k_bwd, all levels)", "bound": "hbm", "achieved": trsv_gbs, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": trsv_gbs / peak, "traffic": traffic, "algorithmic_bytes_per_launch_set": by["trsv"], "ms": ms_trsv / args.steps,
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
    if cplx:
        out["cpu_baseline"] = None   # the CPU arm (oracle/cpu_ras.cpp) is real-valued; the complex bench is auxiliary
    elif args.cpu_baseline and rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline(args, m=args.cpu_m or None, steps=5, budget_s=60.0)
    if rank == 0:
        print(json.dumps(out))
    deco.close()
    if world > 1:
        dist.destroy_process_group()


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
    return L


def cpu_baseline(args, m=None, steps=None, budget_s=150.0):
    """CPU arm = oracle/cpu_ras.cpp: the reference's path restated for the host cores (supernodal
    sparse Cholesky + dtrsv/dgemv supernodal solves as CHOLMOD/MUMPS do, dgemv projections, OpenMP
    CSR SpMV), all host threads, ONE subdomain of the same workload.  The subdomain edge is reduced
    (and reported) if the CPU factorisation would not fit the time budget."""
    L = _cpu_lib()
    # scipy's OpenBLAS is built for at most 128 threads *including callers*: stay well below
    avail = usable_cpus()
    threads = max(1, min(avail, 64))
    m = m or args.m
    # probe: factorisation time scales ~ m^6
    probe = min(m, 48)
    t0 = time.time()
    Zp = cosine_modes((probe,) * 3, args.nu)
    h = L.cpu_ras_create(probe, args.nu, Zp.ctypes.data, threads)
    t_probe = time.time() - t0
    L.cpu_ras_destroy(h)
    while m > probe and t_probe * (m / probe) ** 6 > budget_s:
        m -= 16
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
    return {"value": 1.0 / dt, "unit": "subdomain-applies/s", "cores": threads, "host_cpus": os.cpu_count(), "affinity_cpus": avail, "kind": "port", "ms_per_apply": dt * 1e3,
            "effective_gbs": (2 * 8 * nnz + 8 * n * (2 * args.nu + 12 + 7 * 1.5)) / dt / 1e9,
            "sample": f"oracle/cpu_ras.cpp (supernodal Cholesky + BLAS-2 supernodal solves, OpenMP x {threads} threads), one subdomain of {m}^3 cells"
                      f"{'' if m == args.m else f' (reduced from {args.m}^3 to fit the CPU time budget)'}, nu={args.nu}, {k} applies, nnz(L)={nnz:.3g}, CPU analysis+numfact {tf:.1f}s (setup {t_create:.1f}s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=args.steps)
    # the host cores are shared by all subdomains of the job: N subdomains advance at the rate of one
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": cb["unit"], "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": cb["sample"]}, "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", dest="m", type=int, default=int(os.environ.get("HPDDM_B200_BENCH_M", 128)), help="cells per subdomain edge")
    ap.add_argument("--nu", type=int, default=20)
    ap.add_argument("--rhs", dest="mu", type=int, default=1, help="right-hand sides per apply (block methods)")
    ap.add_argument("--cpu-cells", dest="cpu_m", type=int, default=0, help="subdomain edge of the CPU sample (0 = same as --cells)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--scalar", default="d", choices=["d", "z"], help="d: real FP64 Poisson (headline); z: complex FP64 Helmholtz / ORAS (config 5 shape, auxiliary)")
    ap.add_argument("--wavenumber", type=float, default=2.0)
    ap.add_argument("--krylov", action="store_true", help="also time a full GMRES solve: device-resident driver vs host-driven loop over the C ABI")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
