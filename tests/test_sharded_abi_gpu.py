"""GPU: the row-sharded kNN of the C ABI (orbx_comm_* / orbx_knn2_query_sharded, include/orbx.h) driven from C, one process per GPU,
no Python in the ranks (tests/harness/knn2_sharded_driver.c) -- what the C++ backend of slam_backends/orb_slam_3 (CMakeLists.txt:109-117
links the libraries) would do.  One rank runs on any GPU box; the multi-rank case needs two or more GPUs and is launched through
torch.distributed.run --no-python exactly as a job launcher would."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path_factory.mktemp("sharded") / "knn2_sharded_driver")
    libdir = os.path.join(ROOT, "send_slam_b200")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-D_DEFAULT_SOURCE", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "harness", "knn2_sharded_driver.c"), "-L", libdir, "-lorbx", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def _env():
    env = dict(os.environ)
    try:                                   # the NCCL torch ships, for a process that does not import torch
        import nvidia.nccl
        cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        if os.path.exists(cand):
            env.setdefault("ORBX_NCCL_LIB", cand)
    except Exception:
        pass
    return env


def test_single_rank_communicator(driver):
    env = _env()
    env.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_PORT="29613")
    p = subprocess.run([driver, "200003", "333"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "knn2_sharded_driver OK rank 0 of 1" in p.stdout, (p.stdout, p.stderr)


def test_two_or_more_ranks(driver):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    n = min(n, 4)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29614", driver, "1000003", "777"], env=_env(), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.count("knn2_sharded_driver OK") == n, (p.stdout[-2000:], p.stderr[-2000:])
