"""Multi-GPU path on real devices (NCCL): skipped unless the box has at least two GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        from mcevidence_b200 import _lib
        return _lib.load_library().mce_device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason='needs two GPUs')
def test_sharded_equals_single_gpu_nccl():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tools', 'dist_check.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert 'DIST_CHECK OK' in p.stdout
