"""CPU tier: the C oracle against the committed golden vectors and against live cv2 (the reference's
own implementation of the path, src/extractor/extractor.py:44).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import assert_lk_equal, load_golden
from visual_odom_pipeline_b200 import synth as S

GOLDEN_LK = ["lk_default", "lk_reference", "lk_hard", "lk_count_only", "lk_eps_only", "lk_mineig_flag", "lk_level0"]


@pytest.mark.parametrize("name", GOLDEN_LK)
def test_oracle_matches_golden_lk(oracle, name):
    g = load_golden(name)
    got = oracle.calc_optical_flow_pyr_lk(g["prev"], g["next"], g["prevPts"], None, tuple(int(v) for v in g["winSize"]),
                                          int(g["maxLevel"]), (int(g["criteria"][0]), int(g["criteria"][1]), float(g["criteria"][2])),
                                          flags=int(g["flags"]))
    assert got[0].shape == g["nextPts"].shape
    assert_lk_equal(got, (g["nextPts"], g["status"], g["err"]), name)


def test_oracle_matches_golden_pyramid(oracle):
    g = load_golden("pyramid")
    lv = g["img"]
    for k in ("l1", "l2", "l3", "l4"):
        lv = oracle.pyr_down(lv)
        assert np.array_equal(lv, g[k]), k
    assert oracle.pyr_max_level(241, 135, (21, 21), 8) == int(g["top_win21_max8"])
    d = oracle.scharr(g["img"])
    assert np.array_equal(d[..., 0], g["scharr_x"]) and np.array_equal(d[..., 1], g["scharr_y"])


cv2 = pytest.importorskip("cv2")

LIVE = [
    # h, w, n, win, maxLevel, criteria, motion, margin, extra
    (480, 640, 500, (21, 21), 3, (3, 30, 0.01), S.BENIGN, 0, {}),                       # BASELINE configs[0]
    (376, 1241, 700, (21, 21), 3, (3, 30, 0.01), S.BENIGN, 0, {}),                      # configs[1] shape
    (376, 1241, 500, (31, 31), 3, (3, 30, 0.03), S.HARD, 60, dict(noise_sigma=3.0, flat_cols=(400, 700))),
    (240, 320, 300, (20, 12), 3, (3, 30, 0.01), S.HARD, 30, {}),
    (240, 320, 300, (13, 29), 3, (3, 10, 0.01), S.HARD, 30, {}),
    (240, 320, 300, (3, 3), 5, (1, 10, 0.01), S.BENIGN, 30, {}),
    (240, 320, 300, (24, 24), 8, (2, 30, 0.05), S.BENIGN, 30, {}),
    (50, 70, 100, (21, 21), 3, (3, 30, 0.01), S.BENIGN, 30, {}),
]


@pytest.mark.parametrize("case", LIVE, ids=lambda c: "%dx%d_win%dx%d_L%d" % (c[1], c[0], c[3][0], c[3][1], c[4]))
def test_oracle_matches_live_cv2(oracle, case):
    h, w, n, win, lvl, crit, motion, margin, kw = case
    a, b = S.frame_pair(h, w, seed=7, motion=motion, **kw)
    p = S.uniform_points(n, h, w, seed=3, margin=margin)
    want = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=win, maxLevel=lvl, criteria=crit)
    got = oracle.calc_optical_flow_pyr_lk(a, b, p, None, win, lvl, crit)
    assert_lk_equal(got, want)


def test_oracle_pyrdown_live_cv2_odd_sizes(oracle):
    for (h, w) in [(376, 1241), (135, 241), (47, 156), (33, 17), (5, 7)]:
        img = S.texture(h, w, seed=h + w).astype(np.uint8)
        assert np.array_equal(oracle.pyr_down(img), cv2.pyrDown(img)), (h, w)


def test_oracle_special_points(oracle):
    """NaN / far-outside / border points: status and nextPts semantics of SURVEY.md A.6."""
    a, b = S.frame_pair(120, 160, seed=2)
    p = np.array([[np.nan, 10], [1e9, 1e9], [-500, 20], [0, 0], [159.9, 119.9], [80, 60], [-10.5, -10.5]], np.float32).reshape(-1, 1, 2)
    want = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    got = oracle.calc_optical_flow_pyr_lk(a, b, p, None, (21, 21), 3, (3, 30, 0.01))
    assert np.array_equal(got[1], want[1])
    fin = np.isfinite(want[0]).all(-1).ravel()
    assert np.array_equal(got[0].reshape(-1, 2)[fin].view(np.uint32), want[0].reshape(-1, 2)[fin].view(np.uint32))
    assert np.isnan(got[0].reshape(-1, 2)[0]).any()


def test_oracle_initial_flow_flag(oracle):
    a, b = S.frame_pair(120, 160, seed=4)
    p = S.uniform_points(50, 120, 160, seed=1)
    guess = (p + np.float32(2.0)).astype(np.float32)
    want = cv2.calcOpticalFlowPyrLK(a, b, p, guess.copy(), winSize=(21, 21), maxLevel=2, criteria=(3, 30, 0.01),
                                    flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    got = oracle.calc_optical_flow_pyr_lk(a, b, p, guess.copy(), (21, 21), 2, (3, 30, 0.01), flags=4)
    assert_lk_equal(got, want)
