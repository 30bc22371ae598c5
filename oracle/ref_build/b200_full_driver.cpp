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
    A.setCommunicator(rankWorld, sizeWorld, [](void *id) { MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD); });
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
    int it = HPDDM::IterativeMethod::solve(A, f, sol, mu, MPI_COMM_WORLD);
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
