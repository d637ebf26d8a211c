"""
Build-container check (skipped where /root/reference does not exist, e.g. on the GPU box): the committed fixtures are
what the UNMODIFIED reference produces today.  A sample of the generators of oracle/gen_golden.py is re-run into a
temporary directory -- the reference's own Sampler / proposals / PTSampler under np.random.seed(s) with the recording
RNG -- and every array must equal the committed tests/golden/*.npz bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import gen_golden, refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(),
                                reason="needs the reference tree (build container only)")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(tmp, names):
    for name in names:
        a = np.load(os.path.join(tmp, name + ".npz"))
        b = np.load(os.path.join(GOLDEN, name + ".npz"))
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert np.array_equal(a[k], b[k], equal_nan=True), "%s[%s] differs from the committed fixture" % (name, k)


@pytest.mark.parametrize("which,names", [
    ("_n1_dense_fixtures", ["hmc4_gauss12d", "adapthmc3_gauss12d", "pcn_gauss12d", "hmcmass3_gauss12d",
                            "adaptmalamass_gauss12d", "hmc5_gauss100d"]),
    ("_n3_pt_fixtures", ["pt_rw_gauss2d", "pt_rw_gauss5d", "pt_rw_gauss12d", "pt_rw_logistic"]),
])
def test_generators_reproduce_the_committed_fixtures(tmp_path, monkeypatch, which, names):
    monkeypatch.setattr(gen_golden, "GOLDEN_DIR", str(tmp_path))
    state = np.random.get_state()
    try:
        with refshim.quiet():
            R = refshim.load_reference()
            getattr(gen_golden, which)(R)
    finally:
        np.random.set_state(state)
    _same(str(tmp_path), names)
