"""GPU: N > 1 host logic of the PRODUCT on the driver's single-GPU box -- two processes (torchrun, gloo control plane through
hpddm_b200_ctx_comm_init_host) share device 0; halo sums, the coarse all-gather and the Krylov reductions go over the library's
peer-memory fabric (CUDA IPC).  tests/run_multi_gpu_parity.py checks every hot-path entry point and the three device Krylov
drivers of each rank against the CPU oracle of the whole decomposition."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(nproc, env_extra, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PARITY_BOOT="host", PARITY_SAME_GPU="1", PARITY_EXPECT_TRANSPORT="peer-memory fabric", **env_extra)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tests", "run_multi_gpu_parity.py")], env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    assert out.count(" OK") >= nproc and "FAIL" not in out, out[-4000:]


def test_two_ranks_share_one_gpu_over_the_peer_memory_fabric():
    _run(2, {"PARITY_M": "8"}, 29541)


def test_four_ranks_nonuniform_coarse_space_share_one_gpu():
    """2 x 2 x 1 subdomains, 1-3 deflation vectors per rank (the reference's -nonuniform): padded coarse layout + all-gather by peer stores"""
    _run(4, {"PARITY_M": "6", "PARITY_NONUNIFORM": "1"}, 29542)


def test_two_ranks_complex_share_one_gpu():
    _run(2, {"PARITY_M": "8", "PARITY_SCALAR": "z"}, 29543)
