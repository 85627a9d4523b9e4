"""Pin the oracle: our C and numpy restatements vs the reference's own outputs.

The fixtures in tests/golden/ were produced by the reference itself (tests/golden/make_golden.py:
arm_cnn.c compiled in place + its numpy path).  When oracle/_ref/arm_cnn.so is present the C
restatement is additionally compared with that binary live on fresh seeds.
"""
import json
import os

import numpy as np
import pytest

import inputs
import oracle
from oracle import np_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def port():
    return oracle.load_port()


@pytest.mark.parametrize("case", inputs.CONV_CASES, ids=lambda c: c["name"])
def test_c_port_matches_reference_fixtures(case, port, shipped_weights, conv_golden):
    wt = inputs.make_weights(case["weights"], shipped_weights)
    imgs = inputs.make_images(case["images"], case["n"])
    want = conv_golden[case["name"]]
    for i in range(case["n"]):
        got = oracle.port_infer(port, imgs[i], wt, case["shifts"])
        assert np.array_equal(got, want[i]), f"{case['name']}[{i}]"
    if case.get("dump"):
        _, l0, l1 = oracle.port_infer(port, imgs[0], wt, case["shifts"], dump=True)
        assert np.array_equal(l0, conv_golden[case["name"] + "__l0"])
        assert np.array_equal(l1, conv_golden[case["name"] + "__l1"])


@pytest.mark.parametrize("case", inputs.CONV_CASES[:5], ids=lambda c: c["name"])
def test_numpy_port_matches_reference_fixtures(case, shipped_weights, conv_golden):
    wt = inputs.make_weights(case["weights"], shipped_weights)
    kern = np_oracle.unpack_weights(wt)
    imgs = inputs.make_images(case["images"], case["n"])
    for i in range(min(case["n"], 2)):
        got = np_oracle.infer(imgs[i], kern, case["shifts"])
        assert np.array_equal(got, conv_golden[case["name"]][i])


@pytest.mark.parametrize("case", inputs.HW_CASES, ids=lambda c: c["name"])
def test_generic_hw_matches_arm_benchmark_fixture(case, port, shipped_weights, conv_golden):
    wt = inputs.make_weights(case["weights"], shipped_weights)
    img = inputs.make_images(case["images"], 1, case["H"], case["W"])[0]
    got = oracle.port_infer(port, img, wt, case["shifts"], case["H"], case["W"])
    assert np.array_equal(got, conv_golden[case["name"]])


def test_sha_of_tb_case_matches_survey(conv_golden):
    import hashlib
    assert hashlib.sha256(conv_golden["tb_shipped_default"].tobytes()).hexdigest().startswith("ba9d1c552d775a83")


def test_tb_identity_is_maxpool_of_image(conv_golden):
    """sim/top/tb.v stimulus: identity centre tap on channel 0 => L0 ch0 = 2x2 max-pool of the image."""
    img = inputs.tb_image()
    l0 = conv_golden["tb_identity__l0"]
    assert np.array_equal(l0[0], img.reshape(64, 2, 64, 2).max(axis=(1, 3)))
    assert not l0[1:].any() and not conv_golden["tb_identity__l1"].any() and not conv_golden["tb_identity"].any()


def test_live_against_reference_binary(port, shipped_weights):
    ref = oracle.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/arm_cnn.so not built (no reference source on this box)")
    rng = np.random.default_rng(99)
    for trial in range(6):
        wt = shipped_weights if trial % 2 == 0 else inputs.make_weights(("rng", 100 + trial))
        sh = [int(s) for s in rng.integers(0, 16, 3)]
        img = inputs.make_images(("rng", 200 + trial), 1)[0]
        assert np.array_equal(oracle.port_infer(port, img, wt, sh), oracle.ref_infer(ref, img, wt, sh))


def test_scalar_kats(port):
    """relu_tb.v / accumulator_tb.v / conv_core_tb.v known answers through a 1-layer view of the oracle."""
    kats = json.load(open(os.path.join(GOLDEN, "kats.json")))
    for v, want in kats["relu"]:
        assert int(np.clip(np.int32(v) >> 0, 0, 255)) == want
    # conv_core_tb: window 10..90 x all-ones kernel = 450 -> centre pixel of an 8x8 map, shift 0 -> sat 255; shift 1 -> 225
    img = np.zeros((8, 8), dtype=np.uint8)
    img[2:5, 2:5] = np.array(kats["conv_core"]["window"], dtype=np.uint8).reshape(3, 3)
    kern = np.zeros((16, 1, 3, 3), dtype=np.int8)
    kern[0, 0] = 1
    out = np_oracle.conv_layer(img.reshape(1, 8, 8), kern, 1)
    assert out[0, 1, 1] == kats["conv_core"]["expect"] >> 1
    acc = kats["accumulator"]
    assert acc["overwrite"] + acc["add"] == acc["expect"]


def test_bad_arguments(port, shipped_weights):
    img = inputs.tb_image()
    with pytest.raises(ValueError):
        oracle.port_infer(port, img, shipped_weights, (2, 4, 32))
    with pytest.raises(ValueError):
        oracle.port_infer(port, img, shipped_weights, (-1, 4, 6))


def test_unpack_weights_index_formula(shipped_weights):
    """byte of k[o][i][dy][dx] = base + (((o/16)*ic + i)*16 + o%16)*9 + dy*3 + dx."""
    kern = np_oracle.unpack_weights(shipped_weights)
    base = [0, 144, 4752]
    rng = np.random.default_rng(0)
    for L, (ic, oc) in enumerate(np_oracle.LAYERS):
        for _ in range(50):
            o, i, dy, dx = rng.integers(oc), rng.integers(ic), rng.integers(3), rng.integers(3)
            b = shipped_weights[base[L] + (((o // 16) * ic + i) * 16 + o % 16) * 9 + dy * 3 + dx]
            assert kern[L][o, i, dy, dx] == np.int8(b.view(np.int8) if hasattr(b, "view") else b)


def test_tail_oracle_matches_reference_fixtures(conv_golden, tail_golden):
    fc_w, fc_b = inputs.make_fc()
    for case in inputs.CONV_CASES:
        if not case.get("tail"):
            continue
        feats = conv_golden[case["name"]]
        for i in range(feats.shape[0]):
            cls, p, _, _ = np_oracle.classify_vec(feats[i], fc_w, fc_b)
            assert cls == tail_golden[case["name"] + "__cls"][i]
            assert np.array_equal(p, tail_golden[case["name"] + "__probs"][i])
            box, _ = np_oracle.bbox_vec(feats[i], cls, fc_w)
            assert tuple(tail_golden[case["name"] + "__bbox"][i]) == box


# ---- Classifier.get_cam_bbox (pynq_inference.py:349-408): Pillow's bilinear resize restated ----------------------

def _cam_sets(conv_golden):
    sets = [(c["name"], inputs.make_features(c["features"], c["n"])) for c in inputs.CAM_CASES]
    for name in ("rng_shipped_mid", "smooth_shipped", "rng_random_mid"):
        sets.append(("cam_conv_" + name, conv_golden[name][:4]))
    return sets


def test_pil_bilinear_restatement_matches_pillow_fixtures(cam_golden):
    for src, dst in zip(cam_golden["pil__src"], cam_golden["pil__dst"]):
        assert np.array_equal(np_oracle.pil_resize_bilinear_u8(src, 128), dst)
    bounds, kk = np_oracle.pil_bilinear_coeffs(16, 128)
    assert kk.shape == (128, 3) and all(abs(int(r.sum()) - (1 << 22)) <= 2 for r in kk)      # 22-bit fixed point, sums to 1
    assert bounds[0] == (0, 1) and bounds[4] == (0, 2) and bounds[127] == (15, 1) and max(n for _, n in bounds) <= 3


def test_cam_bbox_oracle_matches_reference_fixtures(conv_golden, cam_golden):
    fc_w, _ = inputs.make_fc()
    for name, feats in _cam_sets(conv_golden):
        cls, box, cam = cam_golden[name + "__cls"], cam_golden[name + "__box"], cam_golden[name + "__cam"]
        for i in range(feats.shape[0]):
            for k in range(6):
                cam_u8, b = np_oracle.get_cam_bbox(feats[i], k, fc_w)
                assert b == tuple(box[i, k]), (name, i, k)
                assert np_oracle.get_cam_bbox_levels(cam_u8) == b, (name, i, k)
                if k == cls[i]:
                    assert np.array_equal(cam_u8, cam[i]), (name, i)


def test_cam_threshold_is_an_integer_level_rule():
    """np.percentile(cam_u8/255, 70) then max(., 0.2) then '>' == level > max(level of sorted[11468], 51), also when the
    two elements the percentile interpolates between differ, and around the 0.2 floor (51/255 == 0.2f)."""
    rng = np.random.default_rng(3)

    def float_rule(cam_u8):
        cam_full = cam_u8.astype(np.float32) / 255.0
        mask = cam_full > max(np.percentile(cam_full, 70), 0.2)
        if not mask.any():
            return (0, 0, 127, 127)
        rows, cols = np.any(mask, 1), np.any(mask, 0)
        y1, y2 = np.where(rows)[0][[0, -1]]
        x1, x2 = np.where(cols)[0][[0, -1]]
        return (int(max(0, x1 - 3)), int(max(0, y1 - 3)), int(min(127, x2 + 3)), int(min(127, y2 + 3)))

    cams = []
    for _ in range(150):
        a, b = sorted(rng.integers(0, 256, 2))
        n_a = 11469 + int(rng.integers(-1, 2))
        cams.append(np.concatenate([np.full(n_a, a, np.uint8), rng.integers(b, 256, 16384 - n_a).astype(np.uint8)]))
    for a in (49, 50, 51, 52, 53):
        for n_a in (11468, 11469, 11470):
            cams.append(np.concatenate([np.full(n_a, a, np.uint8), rng.integers(50, 56, 16384 - n_a).astype(np.uint8)]))
    for v in cams:
        rng.shuffle(v)
        assert float_rule(v.reshape(128, 128)) == np_oracle.get_cam_bbox_levels(v.reshape(128, 128))


# ---- pre-processing (realtime_detect.py:582-591): OpenCV's BGR2GRAY + INTER_AREA restated ------------------------

@pytest.mark.parametrize("case", inputs.PREP_CASES, ids=lambda c: c["name"])
def test_preprocess_oracle_matches_cv2_fixtures(case, prep_golden):
    frames = inputs.make_frames(case["frames"], case["n"], case["h"], case["w"])
    got = np.stack([np_oracle.preprocess_bgr(f) for f in frames])
    assert np.array_equal(got, prep_golden[case["name"]])


def test_area_table_properties():
    for S in (129, 200, 300, 480, 720, 1000):
        tab = np_oracle.area_tab(S, 128)
        w = np.zeros(128)
        for d, s, a in tab:
            assert 0 <= s < S and a > 0
            w[d] += a
        assert np.allclose(w, 1.0, atol=1e-6)                     # each output pixel's weights sum to 1
        src = [s for _, s, _ in tab]
        assert src[0] == 0 and src[-1] == S - 1 and all(b - a in (0, 1) for a, b in zip(src, src[1:]))


@pytest.mark.parametrize("case", inputs.ACC24_CASES, ids=lambda c: c["name"])
def test_acc24_restatement_matches_trainer_fixtures(case, shipped_weights, acc24_golden):
    """np_oracle.conv_layer(acc_bits=24) vs outputs of the reference's own fpga_conv_layer (train_cnn.py:101-116)."""
    wt = inputs.make_weights(case["weights"], shipped_weights)
    kern = np_oracle.unpack_weights(wt)
    if case.get("clamp127"):
        kern = [np.maximum(k, -127) for k in kern]
    imgs = inputs.make_images(case["images"], case["n"])
    for i in range(case["n"]):
        got = np_oracle.infer(imgs[i], kern, case["shifts"], acc_bits=24)
        assert np.array_equal(got, acc24_golden[case["name"]][i]), (case["name"], i)
        same = np.array_equal(got, np_oracle.infer(imgs[i], kern, case["shifts"]))
        assert same != case["wraps"]             # the wrap cases really leave the 24-bit range; the others never do


def test_acc24_scalar_kats():
    """accumulator_tb.v:28-61 (500 overwrite + 300 add = 800) stays 800 in 24 bits; 2^23 wraps to -2^23."""
    m = 1 << 23
    wrap = lambda v: ((v + m) % (2 * m)) - m
    assert wrap(500 + 300) == 800 and wrap(m) == -m and wrap(-m - 1) == m - 1 and wrap(288 * 127 * 255) == 288 * 127 * 255 - 2 * m


@pytest.mark.parametrize("case", inputs.PIL_CASES, ids=lambda c: c["name"])
def test_load_image_restatement_matches_reference_fixtures(case, pil_golden):
    """convert('L') + default-filter resize restated (np_oracle.load_image_array) vs pynq_inference.load_image_any's outputs."""
    assert np.array_equal(np_oracle.load_image_array(inputs.make_pil_image(case)), pil_golden[case["name"]])
