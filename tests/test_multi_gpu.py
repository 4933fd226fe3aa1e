"""N>1 on real GPUs (skipped with fewer than 2): sharded envs + the single gather of SURVEY §8e, both
as peer writes over NVLink and as the NCCL collective, against the unsharded oracle."""
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
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("game,per,steps", [
    ("coinrun", 64, 120),
    ("bigfish,bossfight,caveflyer,chaser,climber,coinrun,dodgeball,fruitbot,heist,jumper,leaper,maze,miner,ninja,plunder,starpilot", 32, 80),
])
def test_sharded_gather_matches_unsharded_oracle(ref_lib, product_lib, game, per, steps):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), game, str(per), str(steps)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
