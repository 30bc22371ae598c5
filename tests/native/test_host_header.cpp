// Syntax / instantiation check of hpddm_b200/host/HPDDM_B200.hpp without the reference tree:
// a minimal MatrixCSR with the reference's member names (include/HPDDM_matrix.hpp:156-165).
// Both scalar types: K = double (hpddm_b200_*) and K = std::complex<double> (hpddm_b200z_*).
#include "HPDDM_B200.hpp"
#include <list>
namespace HPDDM {
template <class K>
class MatrixCSR {
public:
  K *a_; int *ia_, *ja_; int n_, m_, nnz_; bool sym_;
};
}
template <class K>
void instantiate() {
  HPDDM::MatrixCSR<K> *M = nullptr;
  HPDDM::B200Sub<K> S;
  S.numfact(M); S.solve((K *)nullptr, 1); S.solve((const K *)nullptr, (K *)nullptr, 1); S.inertia(M); S.deficiency();
  HPDDM::B200Schwarz<K> A;
  std::list<int> o; std::vector<std::vector<int>> r;
  A.setCommunicator(0, 1, [](void *) {}); A.setCommunicatorHost(0, 1, nullptr, nullptr);
  A.initialize(M, o, r); A.setGridHint(1, 1); A.multiplicityScaling(nullptr); double *d = nullptr; A.initialize(d);
  A.solveGEVP(M, 4); A.callNumfact(); K **ev = nullptr; A.setVectors(ev, 1); A.template buildTwo<0>(0);
  A.start(nullptr, (K *)nullptr, 1); A.apply((const K *)nullptr, (K *)nullptr, 1); A.template deflation<false>(nullptr, (K *)nullptr, 1);
  A.GMV(nullptr, (K *)nullptr, 1); A.template exchange<true>(nullptr, 1); A.end(); A.computeResidual(nullptr, nullptr, nullptr, 1);
  A.solve(nullptr, (K *)nullptr, 1, 0); A.solve(nullptr, (K *)nullptr, 4, 1); A.solve(nullptr, (K *)nullptr, 1, 2); A.solve(nullptr, (K *)nullptr, 2, 4, 30, 100, 1e-8, 10); A.solve(nullptr, (K *)nullptr, 2, 5, 30, 100, 1e-8, 10);
  (void)A.getVectors(); A.statistics();
  (void)A.getScaling(); (void)A.getDof(); (void)A.boundaryConditions(); (void)A.prefix();
}
int main() {
  volatile bool run = false;
  if (run) {  // instantiate every member, never run (no GPU in the build container)
    instantiate<double>();
    instantiate<std::complex<double>>();
  }
  return 0;
}
