"""Multi-GPU parity on real devices (skipped on a one-GPU box): the gswm_comm mailbox all-reduce over NVLink, between
processes (CUDA IPC, one rank per GPU under torch.distributed.run) and inside one process (peer access)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gswm(cuda_device):
    import gswm as g
    g.build()
    g._lib.lib()   # raises if the extension cannot be loaded: there is no fallback path
    return g


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_comm_between_processes(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MGPU OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    assert "nccl_c_abi=1" in res.stdout, "gswm_allreduce_counters was not exercised (libnccl not loadable through ctypes)"


def test_comm_local_two_devices(gswm):
    """ncclCommInitAll style: both ranks in this process, one GPU each, mailboxes reached through peer access."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    lib = gswm._lib.lib()
    hs = [C.c_void_p(), C.c_void_p()]
    for r in range(2):
        assert lib.gswm_comm_create(C.byref(hs[r]), r, r, 2, None) == 0
    assert lib.gswm_comm_connect_local((C.c_void_p * 2)(hs[0], hs[1]), 2) == 0
    bufs = [torch.tensor([r + 1, 10 * (r + 1), 0, 0, 0, 7], dtype=torch.int64, device=f"cuda:{r}") for r in range(2)]
    for epoch in range(4):
        for r in range(2):
            with torch.cuda.device(r):
                assert lib.gswm_comm_allreduce_counters(hs[r], bufs[r].data_ptr(), 6, torch.cuda.current_stream(r).cuda_stream) == 0
        for r in range(2):
            torch.cuda.synchronize(r)
        want = [3 * 2 ** epoch, 30 * 2 ** epoch, 0, 0, 0, 14 * 2 ** epoch]       # in place: every call doubles the sum
        assert bufs[0].tolist() == want and bufs[1].tolist() == want
    assert lib.gswm_comm_status(hs[0]) == 0 and lib.gswm_comm_status(hs[1]) == 0
    for r in range(2):
        lib.gswm_comm_destroy(hs[r])
