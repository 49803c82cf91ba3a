"""Multi-GPU: the x-slab decomposition against the single-GPU engine (needs >= 2 GPUs; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_slab.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,N,rho,sweeps,transport", [(2, 20000, 0.5, 9, "peer"), (2, 3000, 0.3, 13, "peer"), (2, 20000, 0.5, 9, "nccl"),
                                                          (4, 40000, 0.55, 9, "peer"), (8, 120000, 0.5, 5, "peer")])
def test_slab_matches_single_gpu(world, N, rho, sweeps, transport):
    """Bit-identical trajectories, with the halo refresh and the migration as stores into the neighbour GPU's memory
    (peer) and with the send/recv transport (nccl)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "slab_check.py"), str(N), str(rho), str(sweeps)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PP_SLAB_HALO=transport))
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stderr[-3000:]
