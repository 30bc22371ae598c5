"""CPU oracle for the HPDDM two-level RAS preconditioner-apply hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU arm.  The product path (``hpddm_b200`` + ``libhpddm_b200.so``) never
imports this package and fails loudly when the CUDA library is missing.

Every function cites the reference ``file:line`` (relative to the HPDDM tree,
hpddm/hpddm @ ce8f7bf, v2.4.0) whose algorithm it restates.

Pinning status: the restatement is checked (tests/test_oracle_vs_ref.py)
against golden vectors produced by the *unmodified* reference headers compiled
in the build container against an in-process MPI shim and the dense LAPACK
local solver (recipe: ``oracle/ref_build/``; outputs ``oracle/_ref/``; goldens:
``tests/golden/``).  The reference itself ships no apply()-level golden
vectors, so those self-generated vectors are the pin.
"""
