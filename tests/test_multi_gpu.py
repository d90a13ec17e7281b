"""N>1 on real GPUs (-m gpu; skipped with fewer than two devices): frame split + atx_allreduce_accum
(one kernel over NVLink peer memory; ncclAllReduce as the fallback, both exercised) equals the sequential render — sample counts exactly, radiance up to float reassociation."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_frame_split_nccl_sum_matches_sequential(built):
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mgpu_worker.py")]
    import os
    env = dict(os.environ, ATX_P2P_TIMEOUT_MS="20000")    # a rank that never joins a reduce fails the others instead of hanging them
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "ok=True" in proc.stdout
