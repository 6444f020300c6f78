"""Multi-rank GPU parity tests (one process per GPU over NCCL, torchrun): row stripes against the unstriped LP.
Skipped when the box has fewer GPUs than the world size.  The checks themselves live in tests/multirank_worker.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", ["pipelined", "halo", "solver"])
def test_row_stripes_on_several_gpus(case, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, f"{case} on {world} ranks failed:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}"
    assert f"multirank case {case} ok on {world} ranks" in r.stdout
