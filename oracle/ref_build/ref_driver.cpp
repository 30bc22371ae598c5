// Golden-vector generator: drives the UNMODIFIED reference headers (HPDDM::Schwarz, the
// reference's examples/generate.cpp, IterativeMethod::solve) and dumps inputs/outputs of the
// hot path per rank.  Built by oracle/ref_build/Makefile against the in-box MPI shim, the
// dense LAPACK SUBDOMAIN/COARSEOPERATOR plugins (-DLAPACKSUB -DDLAPACK) and scipy's OpenBLAS.
// Mirrors the control flow of the reference's examples/schwarz.cpp:40-147 (which cannot dump).
// Built twice: K = double, and -DFORCE_COMPLEX (K = std::complex<double>, examples/schwarz.hpp:56-60); the complex
// build shifts the generator's Poisson matrix to a damped-Helmholtz-like operator  A - (k^2 - i sigma) I  and makes
// the right-hand side, the probe vector and the deflation vectors genuinely complex, so that every conjugation
// convention of the reference (Wrapper<K>::transc, conj in the inner products) is exercised.
// TEST INFRASTRUCTURE ONLY.
#include "schwarz.hpp"  // the reference's examples/schwarz.hpp (K, symCoarse, generate())

#include <cstdio>
#include <string>

static FILE *g_out = nullptr;
static void dump(const char *name, char type, const void *data, long long count) {
  int len = (int)strlen(name);
  fwrite(&len, sizeof(int), 1, g_out);
  fwrite(name, 1, len, g_out);
  fwrite(&type, 1, 1, g_out);
  fwrite(&count, sizeof(long long), 1, g_out);
  fwrite(data, type == 'z' ? 16 : (type == 'd' ? 8 : 4), count, g_out);
}
static void dumpd(const char *name, const double *d, long long n) { dump(name, 'd', d, n); }
static void dumpd(const char *name, const std::complex<double> *d, long long n) { dump(name, 'z', d, n); }
#ifdef FORCE_COMPLEX
static inline K cplx_probe(double re, double im) { return K(re, im); }
#else
static inline K cplx_probe(double re, double) { return re; }
#endif
static void dumpi(const char *name, const int *d, long long n) { dump(name, 'i', d, n); }

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  int rankWorld, sizeWorld;
  MPI_Comm_size(MPI_COMM_WORLD, &sizeWorld);
  MPI_Comm_rank(MPI_COMM_WORLD, &rankWorld);
  HPDDM::Option &opt = *HPDDM::Option::get();
  opt.parse(argc, argv, false,
            {std::forward_as_tuple("overlap=<1>", "Number of grid points in the overlap.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("Nx=<100>", "Number of grid points in the x-direction.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("Ny=<100>", "Number of grid points in the y-direction.", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("generate_random_rhs=<0>", "Number of generated random right-hand sides.", HPDDM::Option::Arg::integer),
             std::forward_as_tuple("symmetric_csr=(0|1)", "Assemble symmetric matrices.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("nonuniform=(0|1)", "Use a different number of eigenpairs to compute on each subdomain.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("deflation_vectors=<0>", "Number of analytic deflation vectors per subdomain (golden runs).", HPDDM::Option::Arg::integer),
             std::forward_as_tuple("penalise=(0|1)", "Impose non-homogeneous Dirichlet data on the side y = 0 by penalisation (golden runs).", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("device_krylov=(0|1)", "Full-seam build only: device-resident Krylov solve (Schwarz<B200Sub, ...>::solveOnDevice) instead of IterativeMethod::solve.", HPDDM::Option::Arg::argument),
             std::forward_as_tuple("solves=<1>", "Number of successive solves with the same operator (golden runs of the recycling drivers).", HPDDM::Option::Arg::positive),
             std::forward_as_tuple("prefix=<string>", "Use a prefix.", HPDDM::Option::Arg::argument)});
  std::string out = getenv("HPDDM_REF_DUMP") ? getenv("HPDDM_REF_DUMP") : "golden";
  out += "_" + std::to_string(rankWorld) + ".bin";
  g_out = fopen(out.c_str(), "wb");
  std::vector<std::vector<int>> mapping;
  mapping.reserve(8);
  std::list<int> o;
  HPDDM::MatrixCSR<K> *Mat, *MatNeumann = nullptr;
  K *f, *sol;
  HPDDM::underlying_type<K> *d = nullptr;
  int ndof;
  generate(rankWorld, sizeWorld, o, mapping, ndof, Mat, MatNeumann, d, f, sol);
  // -generate_random_rhs N: N random right-hand sides from the reference's generator (made consistent with exchange<true>
  // like examples/schwarz.cpp:98) solved together -- pins the multi-RHS semantics of IterativeMethod::GMRES / BGMRES
  const int mu = std::max(1, (int)opt.app()["generate_random_rhs"]);
#ifdef FORCE_COMPLEX
  // A <- A - (k^2 - i sigma) I : indefinite, complex symmetric, non-Hermitian (full CSR storage only)
  for (int i = 0; i < ndof; ++i) {
    for (int k = Mat->ia_[i]; k < Mat->ia_[i + 1]; ++k)
      if (Mat->ja_[k] == i) Mat->a_[k] += K(-3.0, 1.0);
    f[i] = K(std::real(f[i]), 0.25 * std::sin(0.05 * i + rankWorld));
  }
#endif
  if (opt.app().find("penalise") != opt.app().cend() && opt.app()["penalise"] == 1) {
    // Non-homogeneous Dirichlet data g on the side y = 0 of the domain, imposed the way HPDDM users do it (HPDDM_PEN,
    // include/HPDDM_define.hpp:48): diagonal = penalty, right-hand side = penalty * g.  Exercises Subdomain::boundaryConditions
    // The golden runs use the power of two next to HPDDM_PEN, 2^100 = 1.27e30, as the penalty: x = b / diag and diag * x are then
    // exact in every arithmetic (IEEE, with or without FMA contraction, real or complex), so the initial residual on these
    // rows is exactly zero everywhere.  With 1e30 itself fl(fl(b / 1e30) * 1e30) - b is 0 or +-ulp(1e30) ~ 1.4e14 depending
    // on how a platform rounds -- noise GMRES then has to work off, i.e. nothing an iteration-count parity test can pin.
    // (subdomain.hpp:310-336), the division of x in Schwarz::start (schwarz.hpp:501-505), the penalised entries of ||b||
    // (iterative.hpp:461-468) and the masked rows of Schwarz::computeResidual (schwarz.hpp:761-803).
    // Applied after the complex shift so that the penalised diagonal is exactly (1e30, 0): with a diagonal like (1e30, 1) the
    // initial residual b - A (b / diag) on these rows is an ulp of 1e30 ~ 1e14 of rounding noise that depends on how the platform
    // rounds complex products, i.e. nothing a parity test can pin.
    // Local layout of examples/generate.cpp:51-61: row k = (i - iStart) + (iEnd - iStart) * (j - jStart).
    const int Nx = opt.app()["Nx"], Ny = opt.app()["Ny"], overlap = opt.app()["overlap"];
    int xGrid = int(sqrt(sizeWorld));
    while (sizeWorld % xGrid != 0) --xGrid;
    const int yGrid = sizeWorld / xGrid, y = rankWorld / xGrid, x = rankWorld - xGrid * y;
    const int iStart = std::max(x * Nx / xGrid - overlap, 0), iEnd = std::min((x + 1) * Nx / xGrid + overlap, Nx);
    const int jStart = std::max(y * Ny / yGrid - overlap, 0);
    const double pen = std::ldexp(1.0, 100);
    if (jStart == 0)
      for (int i = iStart; i < iEnd; ++i) {
        const int k = i - iStart;
        for (int p = Mat->ia_[k]; p < Mat->ia_[k + 1]; ++p)
          if (Mat->ja_[p] == k) Mat->a_[p] = K(pen);
        for (int nu = 0; nu < mu; ++nu) f[k + nu * ndof] = K(pen) * cplx_probe(1.0 + 0.5 * std::sin(0.3 * i + nu), 0.25 * std::cos(0.2 * i));
      }
  }
  {
    int hdr[4] = {ndof, Mat->nnz_, (int)Mat->sym_, sizeWorld};
    dumpi("header", hdr, 4);
    dumpi("ia", Mat->ia_, ndof + 1);
    dumpi("ja", Mat->ja_, Mat->nnz_);
    dumpd("a", Mat->a_, Mat->nnz_);
    dumpd("d_ramp", d, ndof);
#ifdef FORCE_COMPLEX
    dumpd("f_local", f, ndof);  // perturbed per rank: not yet consistent on the overlap (made so below, like schwarz.cpp:98)
#else
    if (mu == 1) dumpd("f", f, ndof);
#endif
    std::vector<int> ov(o.begin(), o.end());
    dumpi("o", ov.data(), ov.size());
    for (size_t i = 0; i < mapping.size(); ++i) dumpi(("mapping" + std::to_string(i)).c_str(), mapping[i].data(), mapping[i].size());
  }
  HPDDM::Schwarz<SUBDOMAIN, COARSEOPERATOR, symCoarse, K> A;
  A.Subdomain::initialize(Mat, o, mapping);
  decltype(mapping)().swap(mapping);
  A.multiplicityScaling(d);
  A.initialize(d);
  dumpd("d", d, ndof);
#ifdef FORCE_COMPLEX
  A.exchange<true>(f, mu);  // consistent right-hand side (examples/schwarz.cpp:98 does the same for its random ones)
  dumpd("f", f, (long long)mu * ndof);
#else
  if (mu > 1) {
    A.exchange<true>(f, mu);
    dumpd("f", f, (long long)mu * ndof);
  }
#endif
  const int nuOpt = (int)opt.app()["deflation_vectors"];
  int nu = 0;
  if (nuOpt > 0) {
    // analytic deflation vectors (stand-in for the GenEO eigenvectors: Z is an INPUT of the hot path)
    nu = nuOpt;
    K **deflation = new K *[nu];
    *deflation = new K[(size_t)nu * ndof];
    for (int k = 0; k < nu; ++k) {
      deflation[k] = *deflation + (size_t)k * ndof;
      for (int i = 0; i < ndof; ++i)
        deflation[k][i] = cplx_probe(k == 0 ? 1.0 : std::cos(3.141592653589793 * k * (i + 0.5) / ndof) + 0.1 * ((i * 7 + k) % 5), 0.3 * std::sin(0.01 * (k + 1) * i));
    }
    dumpd("Z", *deflation, (long long)nu * ndof);
    A.setVectors(deflation);
    opt["geneo_nu"] = nu;
    if (!opt.set("schwarz_coarse_correction")) opt["schwarz_coarse_correction"] = HPDDM_SCHWARZ_COARSE_CORRECTION_DEFLATED;
    A.super::initialize(nu);
    A.buildTwo(MPI_COMM_WORLD);
  }
  A.callNumfact();
  // deterministic, consistent input vector
  std::vector<K> v(ndof), w(ndof), work(ndof);
  for (int i = 0; i < ndof; ++i) v[i] = cplx_probe(std::sin(0.37 * i + 0.11 * rankWorld) + 0.5, std::cos(0.23 * i) - 0.2 * rankWorld);
  A.exchange<true>(v.data(), 1);
  dumpd("v", v.data(), ndof);
  {
    std::vector<K> x(v);
    bool alloc = A.setBuffer();
    A.Subdomain::exchange(x.data(), 1);
    dumpd("subdomain_exchange_v", x.data(), ndof);
    A.GMV(v.data(), w.data(), 1);
    A.clearBuffer(alloc);
    dumpd("gmv_v", w.data(), ndof);
  }
  {
    std::vector<K> x(ndof, K());
    bool alloc = A.start(f, x.data(), 1);  // allocates halo buffers + uc_ like the Krylov drivers do
    const double saved = nuOpt > 0 ? opt["schwarz_coarse_correction"] : -1.0;
    if (nuOpt > 0) opt.remove("schwarz_coarse_correction");
    A.apply(v.data(), w.data(), 1, work.data());
    dumpd("apply_onelevel_v", w.data(), ndof);
    if (nuOpt > 0) {
      A.deflation<false>(v.data(), w.data(), 1);
      dumpd("deflation_v", w.data(), ndof);
      const char *names[3] = {"apply_deflated_v", "apply_additive_v", "apply_balanced_v"};
      const double vals[3] = {HPDDM_SCHWARZ_COARSE_CORRECTION_DEFLATED, HPDDM_SCHWARZ_COARSE_CORRECTION_ADDITIVE, HPDDM_SCHWARZ_COARSE_CORRECTION_BALANCED};
      for (int c = 0; c < 3; ++c) {
        opt["schwarz_coarse_correction"] = vals[c];
        A.apply(v.data(), w.data(), 1, work.data());
        dumpd(names[c], w.data(), ndof);
      }
      opt["schwarz_coarse_correction"] = saved;
    }
    A.end(alloc);
  }
  // Built a second time against the FULL seam of the CUDA library (-DB200SUB -DB200SCHWARZ: HPDDM::Schwarz<HPDDM::B200Sub, ...>,
  // oracle/_ref/ref_driver_b200_full): the same dumps then come from the GPU path and are compared with the goldens field by field;
  // -device_krylov 1 routes the solves through solveOnDevice (options read like IterativeMethod::options reads them).
#ifdef B200SCHWARZ
  const bool device_krylov = opt.app().find("device_krylov") != opt.app().cend() && opt.app()["device_krylov"] == 1;
#define REF_SOLVE(rhs, x) (device_krylov ? A.solveOnDevice(rhs, x, mu) : HPDDM::IterativeMethod::solve(A, rhs, x, mu, A.getCommunicator()))
#else
#define REF_SOLVE(rhs, x) HPDDM::IterativeMethod::solve(A, rhs, x, mu, A.getCommunicator())
#endif
  int it = REF_SOLVE(f, sol);
  std::vector<HPDDM::underlying_type<K>> storage(2 * mu);
  A.computeResidual(sol, f, storage.data(), mu);
  dumpi("iterations", &it, 1);
  dumpd("sol", sol, (long long)mu * ndof);
  dumpd("residual", storage.data(), 2 * mu);
  if (rankWorld == 0) printf("ref_driver: %d ranks, ndof %d, nu %d, it %d, residual %.3e / %.3e\n", sizeWorld, ndof, nu, it, storage[1], storage[0]);
  // -solves N: N - 1 further solves with the same operator and new right-hand sides, zero initial guess -- the recycled subspace
  // (U, C) kept in A.storage() by GCRO-DR (GCRODR.hpp:64-69,94-130,245) is reused: f_s[:, nu] = (1 + nu / 2) A w_s + f[:, nu] / 4 with
  // w_s a consistent vector (linear combinations of consistent vectors stay consistent)
  const int solves = opt.app().find("solves") != opt.app().cend() ? (int)opt.app()["solves"] : 1;
  for (int sidx = 2; sidx <= solves; ++sidx) {
    std::vector<K> ws(ndof), Aw(ndof), fs((size_t)mu * ndof);
    for (int i = 0; i < ndof; ++i) ws[i] = cplx_probe(std::cos(0.19 * sidx * i + 0.07 * rankWorld) + 0.25, std::sin(0.13 * i) + 0.1 * rankWorld);
    A.exchange<true>(ws.data(), 1);
    bool alloc = A.setBuffer();
    A.GMV(ws.data(), Aw.data(), 1);
    A.clearBuffer(alloc);
    for (int c = 0; c < mu; ++c)
      for (int i = 0; i < ndof; ++i) fs[(size_t)c * ndof + i] = K(1.0 + 0.5 * c) * Aw[i] + K(0.25) * f[(size_t)c * ndof + i];
    std::fill_n(sol, (size_t)mu * ndof, K());
    int its = REF_SOLVE(fs.data(), sol);
    A.computeResidual(sol, fs.data(), storage.data(), mu);
    const std::string tag = std::to_string(sidx);
    dumpd(("f" + tag).c_str(), fs.data(), (long long)mu * ndof);
    dumpi(("iterations" + tag).c_str(), &its, 1);
    dumpd(("sol" + tag).c_str(), sol, (long long)mu * ndof);
    dumpd(("residual" + tag).c_str(), storage.data(), 2 * mu);
    if (rankWorld == 0) printf("ref_driver: solve %d, it %d, residual %.3e / %.3e\n", sidx, its, storage[1], storage[0]);
  }
  fclose(g_out);
  delete[] d;
  delete MatNeumann;
  delete[] sol;
  delete[] f;
  MPI_Finalize();
  return 0;
}
