"""
The example scripts (device versions of the reference's examples/test_randomwalk.py and examples/test_changepoint.py)
and the ragged-shape sweep over every sampler family run to completion on a B200 and print sane summaries.
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rel, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, rel)], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_randomwalk_example():
    out = _run("examples/randomwalk_device.py")
    m = re.search(r"one chain: (\d+) samples kept, acceptance fraction ([0-9.]+), scale ([0-9.]+)", out)
    assert m and int(m.group(1)) == 9001 and 0.3 < float(m.group(2)) < 0.99                 # sampler.py:53-54 slicing
    m = re.search(r"65536 chains x 10000 steps: acceptance ([0-9.]+), max R-hat ([0-9.]+)", out)
    assert m and abs(float(m.group(1)) - 0.25) < 0.05 and float(m.group(2)) < 1.01           # randomwalk.py:36 target


def test_changepoint_example():
    out = _run("examples/changepoint_device.py")
    m = re.search(r"one chain: (\d+) states, acceptance fraction ([0-9.]+), k after burn-in: mean ([0-9.]+)", out)
    assert m and int(m.group(1)) == 20001 and 0.03 < float(m.group(2)) < 0.6 and 3.0 < float(m.group(3)) < 12.0
    m = re.search(r"65536 chains x 10000 steps: acceptance ([0-9.]+), overflows (\d+), mean sigma ([0-9.]+), mean k ([0-9.]+)", out)
    assert m and 0.05 < float(m.group(1)) < 0.5 and 0.05 < float(m.group(3)) < 0.2 and 4.0 < float(m.group(4)) < 10.0


def test_every_family_on_ragged_shapes():
    out = _run("scripts/sanitize_small.py", timeout=900)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) >= 10
    assert lines[-1].startswith("sanitize_small: all cases ran")
    ok = [l for l in lines if re.search(r"\bok\b", l)]
    bad = [l for l in lines[:-1] if l not in ok and "refused:" not in l]
    assert not bad, bad
    assert len(ok) >= 40, len(ok)
