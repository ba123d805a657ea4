"""GPU: the opt-in persistent dense-round kernel (MVS_DENSE_PERSISTENT=1, csrc/mvs_dense.cu) against the default
four-launch rounds: the same device bodies run in both, so a complete 4-stage fit with the SDF term must give the same
parameters BIT FOR BIT (and the same iteration / evaluation counts).  Each variant runs in its own process (the switch is
read once per process)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("B", [3, 37])
def test_persistent_rounds_equal_multikernel_rounds_bitwise(B, tmp_path):
    outs = []
    for name, env in (("multi", {}), ("persistent", {"MVS_DENSE_PERSISTENT": "1"})):
        out = str(tmp_path / (name + ".npz"))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dense_ab.py"), out, str(B)], capture_output=True, text=True,
                           timeout=300, env=dict(os.environ, **env))
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(np.load(out))
    a, b = outs
    assert int(a["it"]) == int(b["it"]) and int(a["ev"]) == int(b["ev"])
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["final"], b["final"])
