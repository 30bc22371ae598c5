"""Dry run of tests/test_gpu_golden.py on a machine without a GPU: the oracle stands in for the library behind the Decomposition
interface, so that the TEST CODE itself (names, golden keys, control flow of every golden case) is exercised before it is sent
to a GPU box.  Not a test of the product.

    python tests/tools/dryrun_gpu_golden.py [case ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.krylov import OracleOperator, bgmres, cg, gmres  # noqa: E402
from oracle.schwarz import SchwarzWorld  # noqa: E402
from tests.golden_util import cases  # noqa: E402


class _Sub:
    def __init__(self, deco, r):
        self.deco, self.r = deco, r

    def setVectors(self, Z):
        self.deco.Z[self.r] = np.asfortranarray(Z)

    def boundaryConditions(self):
        return dict(self.deco.w.boundary_conditions(self.r))


class FakeDecomposition:
    def __init__(self, parts, method):
        self.w = SchwarzWorld(parts, method=method)
        self.w.multiplicity_scaling()
        self.w.numfact()
        self.P = len(parts)
        self.Z = [None] * self.P
        self.subs = [_Sub(self, r) for r in range(self.P)]
        self.full = all(len(p["o"]) == self.P - 1 for p in parts)

    def multiplicityScaling(self, ds):
        return [d.copy() for d in self.w.d]

    def exchange(self, v, scaled=True):
        v = [np.array(x, order="F", copy=True) for x in v]
        return self.w.exchange(v) if scaled else self.w.subdomain_exchange(v)

    def GMV(self, v):
        return self.w.GMV(v)

    def apply(self, v, corr):
        return self.w.apply([np.array(x, order="F", copy=True) for x in v], corr)

    def buildTwo(self):
        self.w.set_vectors(self.Z)
        self.w.build_coarse(lapack_tr_quirk=True)

    def getCoarse(self):
        return self.w.E.T.copy()

    def setCoarse(self, E):
        pass

    def deflation(self, v):
        return self.w.deflation([np.array(x, order="F", copy=True) for x in v])

    def start(self, b, x):
        return self.w.start(b, x)

    def dot(self, x, y):
        return self.w.dot(x, y)

    def rhs_norm(self, b):
        return self.w.rhs_norm(b)

    def solve(self, b, correction=None, restart=40, max_it=100, tol=1e-6):
        it, x, _ = gmres(OracleOperator(self.w, correction), b, restart=restart, max_it=max_it, tol=tol)
        return it, x, np.zeros(b[0].shape[1])

    def solve_cg(self, b, correction=None, max_it=100, tol=1e-6):
        it, x = cg(OracleOperator(self.w, correction), b, max_it=max_it, tol=tol)
        return it, x, np.zeros(b[0].shape[1])

    def solve_bgmres(self, b, correction=None, restart=40, max_it=100, tol=1e-6):
        it, x = bgmres(OracleOperator(self.w, correction), b, restart=restart, max_it=max_it, tol=tol)
        return it, x, np.zeros(b[0].shape[1])

    # stand-ins for the device-resident GCRO-DR entry points (tests/tools/run_gcrodr_device.py)
    launches = 1

    def solve_gcrodr(self, b, correction=None, restart=40, recycle=10, max_it=100, tol=1e-6, target="SM", strategy="A", same_system=0):
        from oracle.gcrodr import gcrodr
        it, x, self.state = gcrodr(OracleOperator(self.w, correction), b, restart=restart, recycle=recycle, max_it=max_it, tol=tol, state=getattr(self, "state", None),
                                   target=target, strategy=strategy, same_system=same_system)
        return it, x, np.zeros(b[0].shape[1])

    def solve_bgcrodr(self, b, correction=None, restart=40, recycle=10, max_it=100, tol=1e-6, target="SM", strategy="A", same_system=0):
        from oracle.gcrodr import bgcrodr
        it, x, self.state = bgcrodr(OracleOperator(self.w, correction), b, restart=restart, recycle=recycle, max_it=max_it, tol=tol, state=getattr(self, "state", None),
                                    target=target, strategy=strategy, same_system=same_system)
        return it, x, np.zeros(b[0].shape[1])

    def recycle_dim(self):
        st = getattr(self, "state", None)
        return st["k"] if st and st.get("U") is not None else 0

    def recycle_destroy(self):
        self.state = None

    def computeResidual(self, sol, b, kind="l2"):
        return self.w.compute_residual(sol, b, kind)

    def close(self):
        pass


def main(names):
    import tests.test_gpu_golden as T
    T.build_gpu_decomposition = lambda parts, world, own_scaling=True, grid_hint=False, method="ras": FakeDecomposition(parts, method)
    import tests.tools.run_gcrodr_device as R
    R.build_gpu_decomposition = T.build_gpu_decomposition
    for name in names or cases():
        T.test_cuda_path_reproduces_the_reference(name)
        if "_gcrodr_" in name or "_bgcrodr_" in name:
            R.main(name)
        print("ok", name, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
