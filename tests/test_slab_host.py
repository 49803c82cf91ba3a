"""CPU (gloo) tests of the multi-process host logic of the slab decomposition."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("world", [2, 3])
def test_slab_exchange_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "slab_host_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert p.stdout.count("ok") == world


@pytest.mark.parametrize("world", [2, 3])
def test_slab_driver_gloo(world):
    """The sweep loop of SlabEngine against a stand-in engine: data directions of the send/recv transport, call
    sequence of the peer-memory transport (tests/slab_driver_check.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29620 + world), os.path.join(ROOT, "tests", "slab_driver_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert p.stdout.count("ok") == world


def test_layer_partition():
    from patchyparticles_b200.slab import layer_partition, neighbours
    assert layer_partition(170, 8) == [(0, 22), (22, 44), (44, 66), (66, 88), (88, 110), (110, 130), (130, 150), (150, 170)]
    assert layer_partition(44, 2) == [(0, 22), (22, 44)]
    with pytest.raises(ValueError):
        layer_partition(6, 4)
    assert neighbours(0, 4) == (3, 1) and neighbours(3, 4) == (2, 0) and neighbours(0, 2) == (1, 1)
