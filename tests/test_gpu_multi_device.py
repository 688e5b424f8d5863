"""GPU: in-process index-range sharding (b200_init(G): one host thread per device, host sum of the partials —
the reference's own split of multi_exp, LFF/algebra/scalar_multiplication/multiexp.tcc:417-438).  Needs two
visible devices; runs tools/multi_gpu_check.py in a fresh process because the session's engine fixture owns
device 0 as a single-device engine."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _visible_devices():
    import torch
    return torch.cuda.device_count()


def test_two_devices_equal_one():
    if _visible_devices() < 2:
        pytest.skip("one visible device: the sharded path is covered by tests/test_host_partials.py (gloo) and by "
                    "the multi-GPU runs recorded under profiles/")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_check.py"), "2", "16"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU in-process check ok" in r.stdout
