"""N > 1 host logic on CPU: two processes over gloo (tests/run_gloo_world2.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_size_2_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "tests", "run_gloo_world2.py")], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "gloo world-2 OK" in res.stdout


# ---- the PRODUCT's N > 1 path between real processes, without a GPU: the library's orchestration layer on the host stand-in of the
# device layer (tests/native/device_mock.cpp), control AND data plane over the host program's all-gather (torch.distributed / gloo
# through hpddm_b200_ctx_comm_init_host).  tests/run_multi_gpu_parity.py is the runner of the multi-GPU parity evidence, unchanged but
# for the library it loads: every rank checks its slice of multiplicityScaling, the assembled coarse operator, the four corrections,
# GMV, deflation, host-driven and device GMRES, BGMRES, CG -- and, here, the device GCRO-DR / BGCRO-DR drivers -- against the oracle
# of the whole decomposition.
import pytest  # noqa: E402


@pytest.fixture(scope="module")
def standin(tmp_path_factory):
    sys.path.insert(0, ROOT)
    from tests.tools.run_gpu_tests_on_stand_in import build
    return build(str(tmp_path_factory.mktemp("standin")), [])


def _parity(standin, nproc, port, **extra):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PARITY_BOOT="host", PARITY_STANDIN=standin, PARITY_GCRODR="1", PARITY_EXPECT_TRANSPORT="peer-memory fabric", **extra)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "run_multi_gpu_parity.py")], env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    assert out.count(" OK") >= nproc and "FAIL" not in out, out[-4000:]


def test_product_two_ranks_over_gloo_on_the_stand_in(standin):
    _parity(standin, 2, 29551, PARITY_M="6")


def test_product_four_ranks_nonuniform_coarse_space_over_gloo_on_the_stand_in(standin):
    _parity(standin, 4, 29552, PARITY_M="5", PARITY_NONUNIFORM="1")


def test_product_two_ranks_complex_over_gloo_on_the_stand_in(standin):
    _parity(standin, 2, 29553, PARITY_M="6", PARITY_SCALAR="z")
