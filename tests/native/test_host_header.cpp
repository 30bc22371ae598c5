// Syntax / instantiation check of hpddm_b200/host/HPDDM_B200.hpp without the reference tree:
// a minimal MatrixCSR with the reference's member names (include/HPDDM_matrix.hpp:156-165).
#include "HPDDM_B200.hpp"
#include <list>
namespace HPDDM {
template <class K>
class MatrixCSR {
public:
  K *a_; int *ia_, *ja_; int n_, m_, nnz_; bool sym_;
};
}
int main() {
  volatile bool run = false;
  if (run) {  // instantiate every member, never run (no GPU in the build container)
    HPDDM::MatrixCSR<double> *M = nullptr;
    HPDDM::B200Sub<double> S;
    S.numfact(M); S.solve((double *)nullptr, 1); S.solve((const double *)nullptr, (double *)nullptr, 1); S.inertia(M); S.deficiency();
    HPDDM::B200Schwarz<double> A;
    std::list<int> o; std::vector<std::vector<int>> r;
    A.setCommunicator(0, 1, [](void *) {});
    A.initialize(M, o, r); A.setGridHint(1, 1); A.multiplicityScaling(nullptr); double *d = nullptr; A.initialize(d);
    A.solveGEVP(M, 4); A.callNumfact(); double **ev = nullptr; A.setVectors(ev, 1); A.buildTwo<0>(0);
    A.start(nullptr, (double *)nullptr, 1); A.apply((const double *)nullptr, (double *)nullptr, 1); A.deflation<false>(nullptr, (double *)nullptr, 1);
    A.GMV(nullptr, (double *)nullptr, 1); A.exchange<true>(nullptr, 1); A.end(); A.computeResidual(nullptr, nullptr, nullptr, 1);
    (void)A.getScaling(); (void)A.getDof(); (void)A.boundaryConditions(); (void)A.prefix();
  }
  return 0;
}
