"""Multi-GPU parity driver (launch with torchrun, one rank per GPU):
every rank hosts one subdomain, halo + coarse gather go over NCCL; each rank
rebuilds the whole decomposition with the CPU oracle (small sizes) and checks its
own slice of apply / deflation / GMV / GMRES.  Exit code != 0 on mismatch.
PARITY_SCALAR=z runs the complex instantiation (hpddm_b200z_*): 3-D Helmholtz, ORAS, plane-wave coarse vectors.
PARITY_BOOT=nccl (default): NCCL bootstrap, collectives over the peer-memory fabric (HPDDM_B200_HALO=nccl forces NCCL everywhere);
PARITY_BOOT=host: control plane over torch.distributed/gloo through hpddm_b200_ctx_comm_init_host, no NCCL at all;
PARITY_SAME_GPU=1 (with PARITY_BOOT=host): every rank uses device 0 -- several processes sharing one GPU over CUDA IPC.
PARITY_STANDIN=<library> (CPU, with PARITY_BOOT=host): load the host stand-in of the device layer (tests/native/device_mock.cpp, built by
tests/tools/run_gpu_tests_on_stand_in.py) instead of libhpddm_b200.so -- the N > 1 host logic of the library between gloo processes on
a machine without a GPU (tests/test_multiproc_cpu.py).  PARITY_GCRODR=1 adds the device GCRO-DR / BGCRO-DR drivers (two solves each)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from hpddm_b200 import Decomposition, KrylovOperator
    from hpddm_b200.examples.generate import generate_world, split_grid_3d
    from oracle.krylov import OracleOperator, bgmres, cg, gmres
    from oracle.schwarz import ADDITIVE, BALANCED, DEFLATED, SchwarzWorld

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    boot = os.environ.get("PARITY_BOOT", "nccl")
    if os.environ.get("PARITY_SAME_GPU"):
        local = 0
    standin = os.environ.get("PARITY_STANDIN")
    if standin:
        from hpddm_b200 import capi
        capi.LIB_PATH = standin          # this process only; the product never loads the stand-in
        local = 0
    else:
        torch.cuda.set_device(local)
    if boot == "host":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = split_grid_3d(world)
    m = int(os.environ.get("PARITY_M", 10))
    N = tuple(g * m for g in grid)
    cplx = os.environ.get("PARITY_SCALAR", "d") == "z"
    if cplx:
        from hpddm_b200.examples.generate import generate_helmholtz3d
        from oracle.schwarz import OG
        parts = [generate_helmholtz3d(r, world, N=N, overlap=1, mu=2, grid=grid, k=2.0, nu=3) for r in range(world)]
        w = SchwarzWorld(parts, method=OG)
        w.multiplicity_scaling()
        w.numfact([p["MatRobin"] for p in parts])
        w.set_vectors([p["Z"] for p in parts])
    else:
        parts = generate_world(world, dim=3, N=N, overlap=1, mu=2, grid=grid, neumann=True)
        w = SchwarzWorld(parts)
        w.multiplicity_scaling()
        w.numfact()
        w.solve_gevp([p["MatNeumann"] for p in parts], nu=3)
    if os.environ.get("PARITY_NONUNIFORM"):   # different number of deflation vectors per rank (reference: -nonuniform)
        w.set_vectors([z[:, :1 + (r % 3)] for r, z in enumerate(w.Z)])
    w.build_coarse()
    deco = Decomposition(local, dtype=np.complex128 if cplx else np.float64)
    if boot == "host":
        deco.comm_init_host_torch()
    else:
        deco.comm_init_torch()
    p = parts[rank]
    s = deco.add(rank)
    s.initialize(p["Mat"], p["o"], p["mapping"])
    s.setGridHint(*p["dims"])
    d = deco.multiplicityScaling([p["d"]])[0]
    ok = np.abs(d - w.d[rank]).max() < 1e-15
    if cplx:
        s.callNumfact(A=p["MatRobin"], method="oras")
    else:
        s.callNumfact()
    s.setVectors(w.Z[rank])
    deco.buildTwo()
    E = deco.getCoarse()
    errs = {"E": np.abs(E - w.E).max() / np.abs(w.E).max()}
    rs = np.random.RandomState(7)
    x_all = [np.asfortranarray(rs.standard_normal(q["f"].shape) + (1j * rs.standard_normal(q["f"].shape) if cplx else 0.0)) for q in parts]
    for name, corr in (("one-level", None), ("deflated", DEFLATED), ("additive", ADDITIVE), ("balanced", BALANCED)):
        ref = w.apply(x_all, corr)[rank]
        got = deco.apply([x_all[rank]], corr)[0]
        errs[name] = np.abs(got - ref).max() / np.abs(ref).max()
    ref = w.GMV(x_all)[rank]
    errs["gmv"] = np.abs(deco.GMV([x_all[rank]])[0] - ref).max() / np.abs(ref).max()
    ref = w.deflation(x_all)[rank]
    errs["deflation"] = np.abs(deco.deflation([x_all[rank]])[0] - ref).max() / np.abs(ref).max()
    b_all = w.exchange([q["f"].copy() for q in parts])
    it_ref, x_ref, _ = gmres(OracleOperator(w, DEFLATED), b_all)
    it_gpu, x_gpu, _ = gmres(KrylovOperator(deco, DEFLATED), [b_all[rank]])
    errs["gmres_x"] = np.abs(x_gpu[0] - x_ref[rank]).max() / np.abs(x_ref[rank]).max()
    it_dev, x_dev, _ = deco.solve([b_all[rank]], correction=DEFLATED)     # device-resident driver (hpddm_b200[z]_solve)
    errs["gmres_dev_x"] = np.abs(x_dev[0] - x_ref[rank]).max() / np.abs(x_ref[rank]).max()
    # block GMRES (2 right-hand sides in one block Krylov space) and, on the symmetric one-level ASM variant, CG: device drivers
    # whose reductions (Gram matrices, D-weighted products) cross the processes through NCCL
    b2_all = w.exchange([q["f"][:, :2].copy() for q in parts])
    it_bref, x_bref = bgmres(OracleOperator(w, DEFLATED), b2_all)
    it_bdev, x_bdev, _ = deco.solve_bgmres([b2_all[rank]], correction=DEFLATED)
    errs["bgmres_dev_x"] = np.abs(x_bdev[0] - x_bref[rank]).max() / np.abs(x_bref[rank]).max()
    it_cdev = it_cref = 0
    if not cplx:
        from oracle.schwarz import SY
        w.type = SY
        s.callNumfact(method="asm")
        it_cref, x_cref = cg(OracleOperator(w, None), b2_all, tol=1e-8)
        it_cdev, x_cdev, _ = deco.solve_cg([b2_all[rank]], correction=None, tol=1e-8)
        errs["cg_dev_x"] = np.abs(x_cdev[0] - x_cref[rank]).max() / np.abs(x_cref[rank]).max()
    if os.environ.get("PARITY_GCRODR"):
        # recycling drivers on all ranks in lockstep: two solves sharing the pair kept in every rank's context, against the oracle
        from oracle.gcrodr import bgcrodr, gcrodr
        for name, fn, dev in (("gcrodr", gcrodr, deco.solve_gcrodr), ("bgcrodr", bgcrodr, deco.solve_bgcrodr)):
            deco.recycle_destroy()
            state = None
            for k_solve, rhs in enumerate((b2_all, [2.0 * v[:, ::-1] + 1.0 for v in b2_all])):
                rhs = w.exchange([np.asfortranarray(v).copy() for v in rhs]) if k_solve else rhs
                it_o, x_o, state = fn(OracleOperator(w, DEFLATED), rhs, restart=8, recycle=3, tol=1e-8, state=state)
                it_d, x_d, _ = dev([rhs[rank]], correction=DEFLATED, restart=8, recycle=3, tol=1e-8)
                errs[f"{name}{k_solve}_dev_x"] = np.abs(x_d[0] - x_o[rank]).max() / np.abs(x_o[rank]).max()
                if it_d != it_o or errs[f"{name}{k_solve}_dev_x"] > 1e-6:
                    ok = False
                    print(f"rank {rank}: {name} solve {k_solve}: {it_d} iterations, oracle {it_o}", flush=True)
    bad = (not ok) or it_gpu != it_ref or it_dev != it_ref or errs["gmres_dev_x"] > 1e-7 or it_bdev != it_bref or errs["bgmres_dev_x"] > 1e-7 or \
        it_cdev != it_cref or errs.get("cg_dev_x", 0.0) > 1e-6 or any(v > 1e-10 for k, v in errs.items() if not (k.startswith("gmres") or k.endswith("_dev_x"))) or errs["gmres_x"] > 1e-7
    want = os.environ.get("PARITY_EXPECT_TRANSPORT")
    if want and deco.transport != want:
        bad = True
    print(f"rank {rank}/{world} {'complex' if cplx else 'real'} [{deco.transport}]: d_ok={ok} it_gpu={it_gpu} it_dev={it_dev} it_ref={it_ref} bgmres={it_bdev}/{it_bref} cg={it_cdev}/{it_cref} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + (" FAIL" if bad else " OK"), flush=True)
    t = torch.tensor([1.0 if bad else 0.0], device="cuda" if boot != "host" else "cpu")
    dist.all_reduce(t)
    deco.close()
    dist.destroy_process_group()
    sys.exit(1 if t.item() > 0 else 0)


if __name__ == "__main__":
    main()
