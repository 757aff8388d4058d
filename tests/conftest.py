import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C CPU oracle (test infrastructure; compiled with gcc on first use)."""
    from oracle import klt_oracle
    klt_oracle.build()
    return klt_oracle


@pytest.fixture(scope="session")
def native_lib():
    """libklt_b200.so, built in-tree with nvcc if missing (cross-compiles without a GPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("klt_build", os.path.join(ROOT, "visual-odom-pipeline_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="session")
def klt(native_lib):
    import visual_odom_pipeline_b200 as K
    return K


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def assert_lk_equal(got, want, what=""):
    """Bit-exact nextPts / status, err where both status == 1 (cv2 leaves err undefined elsewhere,
    SURVEY.md A.6).  Logs every mismatching point."""
    import numpy as np
    q1, s1, e1 = want
    q2, s2, e2 = got
    q1 = np.asarray(q1, np.float32).reshape(-1, 2); q2 = np.asarray(q2, np.float32).reshape(-1, 2)
    s1 = np.asarray(s1).reshape(-1); s2 = np.asarray(s2).reshape(-1)
    e1 = np.asarray(e1, np.float32).reshape(-1); e2 = np.asarray(e2, np.float32).reshape(-1)
    dp = (q1.view(np.uint32) != q2.view(np.uint32)).any(-1)
    ds = s1 != s2
    both = (s1 == 1) & (s2 == 1)
    de = (e1.view(np.uint32) != e2.view(np.uint32)) & both
    badidx = np.nonzero(dp | ds | de)[0]
    lines = ["%s point %d: want %s st=%d err=%r got %s st=%d err=%r" % (what, i, q1[i], s1[i], e1[i], q2[i], s2[i], e2[i])
             for i in badidx[:20]]
    assert badidx.size == 0, "%d/%d points differ\n%s" % (badidx.size, s1.size, "\n".join(lines))
