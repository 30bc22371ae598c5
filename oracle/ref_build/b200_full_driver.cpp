// The reference's Krylov drivers (IterativeMethod::solve -> GMRES/CG, include/HPDDM_iterative.hpp,
// include/HPDDM_GMRES.hpp) and generator (examples/generate.cpp), UNMODIFIED, on top of
// HPDDM::B200Schwarz (hpddm_b200/host/HPDDM_B200.hpp): the whole preconditioner apply, GMV and the halo
// exchanges run on the GPUs (one rank per GPU, NCCL), the Krylov recurrences stay reference host code.
// Mirrors examples/schwarz.cpp:40-147.  Built by oracle/ref_build/Makefile -> oracle/_ref/b200_full_driver.
#include "schwarz.hpp"  // reference examples header: K, symCoarse, generate()
#include "HPDDM_B200.hpp"

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int rankWorld, sizeWorld;
  MPI_Comm_size(MPI_COMM_WORLD, &sizeWorld);
  MPI_Comm_rank(MPI_COMM_WORLD, &rankWorld);
  {
    const char *nd = getenv("HPDDM_B200_NDEV");
    const int ndev = nd ? atoi(nd) : 1;
    setenv("HPDDM_B200_DEVICE", std::to_string(rankWorld % std::max(1, ndev)).c_str(), 1);
  }
  HPDDM::Option &opt = *HPDDM::Option::get();
  opt.parse(argc, argv, rankWorld == 0,
            {std::forward_as_tuple("overlap=<1>", "Number of grid points in the overlap.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("Nx=<100>", "Number of grid points in the x-direction.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("Ny=<100>", "Number of grid points in the y-direction.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("generate_random_rhs=<0>", "Number of generated random right-hand sides.", HPDDM::Option::Arg::integer),
             std::forward_as_tuple("symmetric_csr=(0|1)", "Assemble symmetric matrices.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("nonuniform=(0|1)", "Use a different number of eigenpairs to compute on each subdomain.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("deflation_vectors=<0>", "Number of analytic deflation vectors per subdomain.", HPDDM::Option::Arg::integer),
             std::forward_as_tuple("device_krylov=(0|1)", "Device-resident Krylov solve (B200Schwarz::solve) instead of the reference's host driver.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("solves=<1>", "Number of successive solves with the same operator (as oracle/ref_build/ref_driver.cpp).", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("prefix=<string>", "Use a prefix.", HPDDM::Option::Arg::argument)});
  if (rankWorld != 0) opt.remove("verbosity");
  std::vector<std::vector<int>> mapping;
  mapping.reserve(8);
  std::list<int> o;
  HPDDM::MatrixCSR<K> *Mat, *MatNeumann = nullptr;
  K *f, *sol;
  HPDDM::underlying_type<K> *d = nullptr;
  int ndof;
  generate(rankWorld, sizeWorld, o, mapping, ndof, Mat, MatNeumann, d, f, sol);
  int mu = opt.app()["generate_random_rhs"];
  int status = 0;
  {
    HPDDM::B200Schwarz<K, HPDDM::OptionsPrefix<K>> A;
    static MPI_Comm world = MPI_COMM_WORLD;
    if (getenv("HPDDM_B200_BOOT") && !strcmp(getenv("HPDDM_B200_BOOT"), "host"))  // control plane over MPI_Allgather, no NCCL
      A.setCommunicatorHost(rankWorld, sizeWorld, [](const void *s, void *r, size_t bytes, void *comm) -> int { return MPI_Allgather(s, (int)bytes, MPI_BYTE, r, (int)bytes, MPI_BYTE, *static_cast<MPI_Comm *>(comm)); }, &world);
    else A.setCommunicator(rankWorld, sizeWorld, [](void *id) { MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD); });
    A.initialize(Mat, o, mapping);
    decltype(mapping)().swap(mapping);
    A.multiplicityScaling(d);
    A.initialize(d);
    if (mu != 0) A.exchange<true>(f, mu);
    else mu = 1;
    const int nu = (int)opt.app()["deflation_vectors"];
    K **deflation = nullptr;
    if (nu > 0) {
      deflation = new K *[nu];
      *deflation = new K[(size_t)nu * ndof];
      for (int k = 0; k < nu; ++k) {
        deflation[k] = *deflation + (size_t)k * ndof;
        for (int i = 0; i < ndof; ++i) deflation[k][i] = k == 0 ? 1.0 : std::cos(3.141592653589793 * k * (i + 0.5) / ndof) + 0.1 * ((i * 7 + k) % 5);
      }
      A.setVectors(deflation, nu);
      A.buildTwo(MPI_COMM_WORLD, HPDDM_B200_CORRECTION_DEFLATED);
    }
    A.callNumfact();
    // -device_krylov 1: the device-resident route of the C++ mirror (B200Schwarz::solve -> hpddm_b200[z]_solve / _bgmres / _cg / _gcrodr /
    // _bgcrodr by -hpddm_krylov_method: Krylov basis and, for the recycling drivers, the pair (U, C) stay in HBM) with the options the
    // reference's drivers read (iterative.hpp:192-218); default: the reference's own IterativeMethod::solve on top of the GPU apply
    const bool device_krylov = opt.app().find("device_krylov") != opt.app().cend() && opt.app()["device_krylov"] == 1;
    auto solve = [&](const K *rhs, K *x) -> int {
      if (!device_krylov) return HPDDM::IterativeMethod::solve(A, rhs, x, mu, MPI_COMM_WORLD);
      return A.solve(rhs, x, mu, opt.val<char>("krylov_method", HPDDM_KRYLOV_METHOD_GMRES), opt.val<unsigned short>("gmres_restart", 40), opt.val<unsigned short>("max_it", 100),
                     opt.val("tol", 1.0e-6), opt.val<int>("recycle", 0));
    };
    int it = solve(f, sol);
    // -solves N: further solves with the same operator and the right-hand sides of ref_driver.cpp (f_s[:, nu] = (1 + nu / 2) A w_s + f[:, nu] / 4)
    const int solves = opt.app().find("solves") != opt.app().cend() ? (int)opt.app()["solves"] : 1;
    for (int sidx = 2; sidx <= solves; ++sidx) {
      std::vector<K> ws(ndof), Aw(ndof), fs((size_t)mu * ndof);
      for (int i = 0; i < ndof; ++i) ws[i] = std::cos(0.19 * sidx * i + 0.07 * rankWorld) + 0.25;
      A.exchange<true>(ws.data(), 1);
      A.GMV(ws.data(), Aw.data(), 1);
      for (int c = 0; c < mu; ++c)
        for (int i = 0; i < ndof; ++i) fs[(size_t)c * ndof + i] = K(1.0 + 0.5 * c) * Aw[i] + K(0.25) * f[(size_t)c * ndof + i];
      std::fill_n(sol, (size_t)mu * ndof, K());
      const int its = solve(fs.data(), sol);
      if (rankWorld == 0) std::cout << "b200_full_driver: solve " << sidx << ", it " << its << std::endl;
      if (its < 0 || its > 45) status = 1;
      if (sidx == solves) std::copy(fs.begin(), fs.end(), f);  // the residual check below is for the last system solved
    }
    HPDDM::underlying_type<K> *storage = new HPDDM::underlying_type<K>[2 * mu];
    A.computeResidual(sol, f, storage, mu);
    if (rankWorld == 0)
      for (unsigned short nu = 0; nu < mu; ++nu) std::cout << (nu == 0 ? " --- residual = " : "                ") << std::scientific << storage[1 + 2 * nu] << " / " << storage[2 * nu] << std::endl;
    if (it > 45) status = 1;
    for (unsigned short nu = 0; nu < mu; ++nu)
      if (storage[1 + 2 * nu] / storage[2 * nu] > 1.0e-2) status = 1;
    if (rankWorld == 0) std::cout << "b200_full_driver: " << sizeWorld << " ranks, it " << it << ", status " << status << std::endl;
    delete[] storage;
    if (deflation) {
      delete[] *deflation;
      delete[] deflation;
    }
  }
  delete Mat;
  delete[] d;
  delete MatNeumann;
  delete[] sol;
  delete[] f;
  MPI_Finalize();
  return status;
}
