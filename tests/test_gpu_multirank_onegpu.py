"""Multi-rank hot path on ONE GPU: P processes share cuda:0 (CUDA IPC maps every rank's peer window into the others,
kernels of different processes time-slice), so the fused pack+send / wait+unpack kernels, the in-kernel scalar
all-reduce of PCG and the three-launch operator split run exactly as on P GPUs - and are compared, rank by rank, with
dumps of the UNMODIFIED reference run on P ranks (tests/golden/mr_*.npz).  See tests/mr_gpu_worker.py for the checks."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name", ["mr_n3_e4x4x4_p2", "mr_n7_e4x2x2_p2", "mr_n2_e5x4x3_p4", "mr_n2_e4x4x4_periodic_p4",
                                  "mr_n7_e2x2x2_p8"])
def test_multirank_on_one_gpu_matches_multirank_reference(name):
    size = int(name.rsplit("_p", 1)[1])
    port = _free_port()
    procs = []
    for r in range(size):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(size), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), LIBP_MR_ONE_GPU="1", LIBP_P2P_TIMEOUT_MS="20000", LIBP_P2P_WINDOW_MB="8",
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mr_gpu_worker.py"), name], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o[-3000:]}"
    assert "MR_GPU_OK" in outs[0], outs[0][-2000:]
