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
