"""GPU, N > 1: the in-library data plane (b200_comm_*, b200_groth16_prove_sharded, b200_msm_sharded) under torchrun,
one rank per GPU (tools/mgpu_check.py holds the checks: golden proofs byte-identical over N GPUs, sharded == single-GPU
proof on a fresh instance, sharded MSM == bn254_msm).  Skipped on a box with one GPU; the single-device shard tests in
test_gpu_groth16.py and the gloo tests in test_multi_rank_gloo.py cover the host logic there."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_in_library_exchange_matches_golden(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mgpu_check.py"), "--fresh", "20000" if world == 2 else "0"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and f"MGPU_CHECK_OK world={world}" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])


def test_comm_rejects_bad_arguments(gpu):
    import ctypes as C
    h = C.c_void_p()
    tok = (C.c_uint8 * 128)()
    assert gpu.dll.b200_comm_create(None, 0, 1, C.byref(h)) != 0
    assert gpu.dll.b200_comm_create(tok, 2, 2, C.byref(h)) != 0   # rank out of range
    assert gpu.dll.b200_comm_create(tok, 0, 0, C.byref(h)) != 0
    assert gpu.dll.b200_comm_destroy(None) != 0
    assert gpu.dll.b200_comm_unique_id(None) != 0
    # a world of one is a valid communicator: the sharded entry points degenerate to the single-GPU ones
    assert gpu.dll.b200_comm_unique_id(tok) == 0
    assert gpu.dll.b200_comm_create(tok, 0, 1, C.byref(h)) == 0
    r, w = C.c_int(-1), C.c_int(-1)
    assert gpu.dll.b200_comm_info(h, C.byref(r), C.byref(w)) == 0 and (r.value, w.value) == (0, 1)
    assert gpu.dll.b200_comm_destroy(h) == 0
