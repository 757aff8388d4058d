"""Row f3, GPU tier: csrc/klt_bilateral.cu through the C ABI against the oracle (bit-exact: same
restatement, B.1-B.6), against OpenCV's own code path (identical except exact rounding ties: <= 1e-4 of the pixels, by 1) and
against the IPP-enabled wheel (within 1) -- reference call site src/loader/loader.py:16-20,86."""
import numpy as np
import pytest

from visual_odom_pipeline_b200 import synth as S

pytestmark = pytest.mark.gpu
TIE_FRACTION_MAX = 1e-4
REF = dict(d=5, sigmaColor=1.5, sigmaSpace=1.5)      # loader.py:16-20


def frames():
    rng = np.random.default_rng(3)
    for (h, w) in [(376, 1241), (480, 640), (768, 1024), (120, 167), (33, 17), (7, 9), (64, 70), (9, 300)]:
        smooth = S.texture(h, w, seed=h + 1).astype(np.uint8)
        yield smooth
        yield np.clip(smooth.astype(int) + rng.integers(-40, 40, smooth.shape), 0, 255).astype(np.uint8)


@pytest.fixture()
def cv2_generic():
    import cv2
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    yield cv2
    cv2.ipp.setUseIPP(was)


def test_bilateral_filter_equals_the_oracle_bit_for_bit(klt, oracle):
    for img in frames():
        for d, sc, ss in [(5, 1.5, 1.5), (3, 10.0, 2.0), (7, 25.0, 3.0), (15, 40.0, 5.0), (0, 4.0, 1.2)]:
            got = klt.bilateralFilter(img, d, sc, ss)
            want = oracle.bilateral_filter(img, d, sc, ss)
            assert got.dtype == np.uint8 and got.shape == img.shape
            assert np.array_equal(got, want), (img.shape, d, sc, ss, int((got != want).sum()))


def test_bilateral_filter_vs_opencv_generic_and_ipp_paths(klt, cv2_generic):
    cv2 = cv2_generic
    total = ties = 0
    for img in frames():
        got = klt.bilateralFilter(img, **REF)
        ref = cv2.bilateralFilter(img, **REF)
        diff = np.abs(got.astype(int) - ref.astype(int))
        assert diff.max() <= 1
        total += diff.size
        ties += int((diff != 0).sum())
        cv2.ipp.setUseIPP(True)
        ipp = cv2.bilateralFilter(img, **REF)
        cv2.ipp.setUseIPP(False)
        assert np.abs(got.astype(int) - ipp.astype(int)).max() <= 1
    assert ties <= TIE_FRACTION_MAX * total, "%d of %d pixels differ from OpenCV's own code path" % (ties, total)


def test_bilateral_filter_contract(klt, oracle):
    import cv2
    img = S.texture(60, 90, seed=2).astype(np.uint8)
    keep = img.copy()
    view = np.ascontiguousarray(np.pad(img, ((0, 0), (3, 5))))[:, 3:93]           # non-contiguous rows
    assert np.array_equal(klt.bilateralFilter(view, 5, 1.5, 1.5), oracle.bilateral_filter(img, 5, 1.5, 1.5))
    assert np.array_equal(img, keep)
    with pytest.raises(cv2.error):
        klt.bilateralFilter(img.astype(np.float32), 5, 1.5, 1.5)
    with pytest.raises(cv2.error):
        klt.bilateralFilter(img, 5, 1.5, 1.5, borderType=cv2.BORDER_CONSTANT)
    with pytest.raises(cv2.error):
        klt.bilateralFilter(img, 31, 1.5, 1.5)                                       # radius 15 > 7: outside the library's limits


def test_batched_device_filter_and_tracker_prefilter(klt, oracle):
    import cv2
    import torch
    from visual_odom_pipeline_b200 import filters as F, tracker as T
    seq = S.sequence(120, 167, 4, seed=8)
    batch = torch.from_numpy(np.stack(seq)).cuda()
    out = F.bilateral_filter(batch, 5, 1.5, 1.5)
    for i, f in enumerate(seq):
        assert np.array_equal(out[i].cpu().numpy(), oracle.bilateral_filter(f, 5, 1.5, 1.5))
    # a tracker fed RAW frames with the loader's filter fused into its frame copy == cv2 LK on frames filtered by the oracle
    lk = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 30, 0.01))
    filt = [oracle.bilateral_filter(f, 5, 1.5, 1.5) for f in seq]
    p = S.uniform_points(50, 120, 167, seed=1)
    trk = T.KLTTracker(prefilter=(5, 1.5, 1.5), **lk).reset(torch.from_numpy(seq[0]).cuda())
    cur = p.copy()
    for t in range(1, len(seq)):
        q, st, er = trk.track(torch.from_numpy(seq[t]).cuda(), torch.from_numpy(cur.reshape(1, -1, 2)).cuda())
        ref = cv2.calcOpticalFlowPyrLK(filt[t - 1], filt[t], cur, None, **lk)
        assert np.array_equal(q.cpu().numpy().reshape(-1, 2).view(np.uint32), ref[0].reshape(-1, 2).view(np.uint32))
        assert np.array_equal(st.cpu().numpy().ravel(), ref[1].ravel())
        cur = ref[0]
