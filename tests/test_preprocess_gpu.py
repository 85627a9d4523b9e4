"""Pre-processing on the GPU (csrc/preprocess.cuh) vs OpenCV through the reference's own lines
(software/realtime_detect.py:582-591): fixtures made with cv2 4.13.0 + the numpy restatement.  u8 outputs, must match exactly."""
import numpy as np
import pytest

import inputs
from oracle import np_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc(shipped_weights):
    import fpga_cnn_b200 as fc
    a = fc.CNNAccelerator()
    a.load_weights(shipped_weights)
    a.set_shifts(7, 10, 11)
    a.load_classifier(*inputs.make_fc())
    yield a
    a.close()


@pytest.mark.parametrize("case", inputs.PREP_CASES, ids=lambda c: c["name"])
def test_preprocess_fixtures(case, acc, prep_golden):
    frames = inputs.make_frames(case["frames"], case["n"], case["h"], case["w"])
    assert np.array_equal(acc.preprocess(frames), prep_golden[case["name"]])


@pytest.mark.parametrize("h,w", [(480, 640), (720, 1280), (129, 129), (333, 517), (1080, 1920), (512, 512), (640, 640), (1024, 1024)])
def test_preprocess_sizes_vs_oracle(h, w, acc):
    frames = inputs.make_frames(("rng", h * 7 + w), 2, h, w)
    frames[1] = inputs.make_frames(("edges", h + w), 1, h, w)[0]
    got = acc.preprocess(frames)
    for i in range(2):
        assert np.array_equal(got[i], np_oracle.preprocess_bgr(frames[i])), (h, w, i)


def test_preprocess_many_frames_and_device_pointers(acc):
    import torch
    frames = inputs.make_frames(("smooth", 91), 70, 480, 640)         # several 16 MiB staging chunks
    got = acc.preprocess(frames)
    for i in (0, 17, 35, 36, 69):
        assert np.array_equal(got[i], np_oracle.preprocess_bgr(frames[i])), i
    t = torch.from_numpy(frames[:5]).cuda()
    assert np.array_equal(acc.preprocess(t).cpu().numpy(), got[:5])


def test_detect_frames_equals_the_separate_steps(acc):
    frames = inputs.make_frames(("smooth", 92), 40, 480, 640)
    gray = acc.preprocess(frames)
    cls, probs, box = acc.infer_batch(gray)
    cls2, probs2, box2, gray2 = acc.detect_frames(frames, return_gray=True)
    assert np.array_equal(gray2, gray) and np.array_equal(cls2, cls) and np.array_equal(probs2, probs) and np.array_equal(box2, box)
    cls3, _, box3 = acc.detect_frames(frames, bbox="upsampled")
    assert np.array_equal(cls3, cls)
    assert np.array_equal(box3, acc.cam_bbox_batch(acc.run_batch(gray).reshape(40, 64, 256), cls))
    # against the CPU chain for a few frames: oracle preprocess -> oracle conv -> classify_vec / bbox_vec
    from oracle import load_port, port_infer
    port = load_port()
    fw, fb = inputs.make_fc()
    wt = np.fromfile(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "weights.bin"), dtype=np.uint8)
    for i in (0, 39):
        g = np_oracle.preprocess_bgr(frames[i])
        f = port_infer(port, g, wt, (7, 10, 11))
        c, p, _, _ = np_oracle.classify_vec(f, fw, fb)
        assert c == cls[i] and np.abs(p - probs[i]).max() <= 1e-5
        assert np_oracle.bbox_vec(f, c, fw)[0] == tuple(box[i])


def test_preprocess_rejects_bad_sizes(acc):
    with pytest.raises(ValueError):
        acc.preprocess(np.zeros((1, 100, 200, 3), dtype=np.uint8))      # crop side < 128
    with pytest.raises(ValueError):
        acc.preprocess(np.zeros((1, 128, 128), dtype=np.uint8))
