"""Runs oracle/_ref/ref_driver (the UNMODIFIED reference headers + examples/generate.cpp,
built by this directory's Makefile) and stores its per-rank dumps as tests/golden/*.npz.

    python oracle/ref_build/make_goldens.py [case ...]      (default: every case)

The goldens pin (a) the driver-side generator, (b) the oracle restatement and (c) the CUDA
path against outputs of the reference itself (tests/test_golden_reference.py).
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
DRIVER_Z = os.path.join(ROOT, "oracle", "_ref", "ref_driver_z")  # the same driver, K = std::complex<double> (-DFORCE_COMPLEX)
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE config 1 (examples/schwarz.cpp test line, Makefile:310 of the reference)
    "config1_100x100_p4_ras": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_gmres_restart=25", "-hpddm_max_it", "80", "-Nx", "100", "-Ny", "100"]),
    # small cases with every hot-path output
    "small_40x40_p4_ras": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-Nx", "40", "-Ny", "40"]),
    "small_40x40_p4_twolevel_nu3": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "40", "-Ny", "40"]),
    "small_48x30_p6_ov2_twolevel_nu2": dict(np=6, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "2", "-Nx", "48", "-Ny", "30", "-overlap", "2"]),
    "small_36x36_p4_symcsr_twolevel_nu2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "2", "-Nx", "36", "-Ny", "36", "-symmetric_csr", "1"]),
    # several right-hand sides advancing together (IterativeMethod::GMRES with mu > 1: per-column hasConverged, GMRES.hpp:90-158),
    # with restarts so that columns converge in different cycles
    "small_40x40_p4_mu3_gmres": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "3"]),
    "small_40x40_p4_mu3_twolevel_restart5": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "40", "-Ny", "40",
                                                             "-generate_random_rhs", "3", "-hpddm_gmres_restart", "5", "-hpddm_tol", "1e-9"]),
    "complex_40x40_p4_mu2_twolevel": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "40", "-Ny", "40",
                                                              "-generate_random_rhs", "2"]),
    # IterativeMethod::CG (HPDDM_CG.hpp:31-168) with the symmetric one-level method (ASM), 1 and 3 right-hand sides
    "small_40x40_p4_asm_cg": dict(np=4, args=["-hpddm_schwarz_method", "asm", "-hpddm_krylov_method", "cg", "-Nx", "40", "-Ny", "40"]),
    "small_40x40_p4_asm_cg_mu3": dict(np=4, args=["-hpddm_schwarz_method", "asm", "-hpddm_krylov_method", "cg", "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "3", "-hpddm_tol", "1e-8"]),
    # IterativeMethod::BGMRES (GMRES.hpp:160-313), 4 right-hand sides (the Krylov method of BASELINE config 4), with and without restarts
    "small_40x40_p4_bgmres_mu4": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgmres", "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "4"]),
    "small_40x40_p4_bgmres_mu4_twolevel_restart6": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-hpddm_krylov_method", "bgmres",
                                                                    "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "4", "-hpddm_gmres_restart", "6", "-hpddm_verbosity", "4"]),
    "complex_40x40_p4_bgmres_mu3": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgmres", "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "3"]),
    # non-homogeneous Dirichlet data imposed by penalisation on one side (diag = 1e30, f = 1e30 g): boundaryConditions, the x = b / diag
    # of Schwarz::start, the penalised entries of ||b|| (iterative.hpp:461-468) and the masked rows of computeResidual (schwarz.hpp:761-803)
    "small_40x40_p4_penalised_ras": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-Nx", "40", "-Ny", "40", "-penalise", "1"]),
    "small_40x40_p4_penalised_twolevel_mu2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "40", "-Ny", "40",
                                                              "-generate_random_rhs", "2", "-penalise", "1"]),
    "small_40x40_p4_penalised_bgmres_mu4": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgmres", "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "4", "-penalise", "1"]),
    "complex_40x40_p4_penalised_ras": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-Nx", "40", "-Ny", "40", "-penalise", "1"]),
    # IterativeMethod::GCRODR (GCRODR.hpp:35-444; the Krylov method of BASELINE config 5): restarted cycles with a recycled subspace,
    # successive solves reusing the pair (U, C) kept in A.storage() (-solves N, see ref_driver.cpp), several right-hand sides,
    # a two-level preconditioner, complex scalars; verbosity 3 keeps the reference's residual history in the golden's log
    "small_40x40_p4_gcrodr_m8_k4_solves3": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "4",
                                                           "-Nx", "40", "-Ny", "40", "-solves", "3", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_gcrodr_m40_k10_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_recycle", "10",
                                                             "-Nx", "40", "-Ny", "40", "-solves", "2", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_gcrodr_m8_k3_mu2_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "3",
                                                               "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "2", "-solves", "2", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_gcrodr_m6_k2_twolevel_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3",
                                                                    "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "6", "-hpddm_recycle", "2", "-hpddm_tol", "1e-9",
                                                                    "-Nx", "40", "-Ny", "40", "-solves", "2", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_gcrodr_m8_k3_sr_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "3",
                                                              "-hpddm_recycle_target", "SR", "-hpddm_tol", "1e-7", "-Nx", "40", "-Ny", "40", "-solves", "2", "-hpddm_verbosity", "3"]),
    # -hpddm_recycle_same_system: the reference raises the option from 1 to 2 after the first converged solve -> solves 2 and 3 use the
    # stored pair as is (no A M^-1 U product, no update of the pair)
    "small_40x40_p4_gcrodr_m12_k4_same_solves3": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "12", "-hpddm_recycle", "4",
                                                                 "-hpddm_recycle_same_system", "1", "-Nx", "40", "-Ny", "40", "-solves", "3", "-hpddm_verbosity", "3"]),
    # IterativeMethod::BGCRODR (GCRODR.hpp:445-907): one block Krylov space and one recycled pair of mu k columns for all right-hand sides
    "small_40x40_p4_bgcrodr_m8_k3_mu2_solves3": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "3",
                                                                "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "2", "-solves", "3", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_bgcrodr_m40_k5_mu3_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgcrodr", "-hpddm_recycle", "5",
                                                                 "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "3", "-solves", "2", "-hpddm_verbosity", "3"]),
    "small_40x40_p4_bgcrodr_m8_k4_solves2": dict(np=4, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "4",
                                                            "-Nx", "40", "-Ny", "40", "-solves", "2", "-hpddm_verbosity", "3"]),
    "complex_40x40_p4_bgcrodr_m8_k2_mu2_solves2": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "bgcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "2",
                                                                          "-Nx", "40", "-Ny", "40", "-generate_random_rhs", "2", "-solves", "2", "-hpddm_verbosity", "3"]),
    "complex_40x40_p4_gcrodr_m8_k3_solves2": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_krylov_method", "gcrodr", "-hpddm_gmres_restart", "8", "-hpddm_recycle", "3",
                                                                     "-Nx", "40", "-Ny", "40", "-solves", "2", "-hpddm_verbosity", "3"]),
    # complex scalars (the reference's FORCE_COMPLEX build; damped-Helmholtz-like shift of the generator's matrix, see ref_driver.cpp)
    "complex_40x40_p4_ras": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-Nx", "40", "-Ny", "40"]),
    "complex_40x40_p4_twolevel_nu3": dict(np=4, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "40", "-Ny", "40"]),
    "complex_48x30_p6_ov2_twolevel_nu2": dict(np=6, z=True, args=["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "2", "-Nx", "48", "-Ny", "30", "-overlap", "2"]),
}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            (ln,) = struct.unpack("i", h)
            name = f.read(ln).decode()
            t = f.read(1).decode()
            (cnt,) = struct.unpack("q", f.read(8))
            size, dtype = {"d": (8, np.float64), "i": (4, np.int32), "z": (16, np.complex128)}[t]
            out[name] = np.frombuffer(f.read(cnt * size), dtype=dtype).copy()
    return out


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle/ref_build` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, HPDDM_SHIM_NP=str(case["np"]), HPDDM_REF_DUMP=os.path.join(tmp, "g"))
            res = subprocess.run([DRIVER_Z if case.get("z") else DRIVER] + case["args"], env=env, cwd=tmp, capture_output=True, text=True, timeout=600)
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("ref_driver:")]
            print(name, line)
            blob = {"args": np.array(" ".join(case["args"])), "np": np.array(case["np"]), "log": np.array(res.stdout[-20000:])}
            for r in range(case["np"]):
                for k, v in read_dump(os.path.join(tmp, f"g_{r}.bin")).items():
                    blob[f"r{r}_{k}"] = v
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


if __name__ == "__main__":
    main()
