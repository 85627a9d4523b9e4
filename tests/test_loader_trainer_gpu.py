"""The two remaining "next" rows of SURVEY.md 8(f): load_image_any's PIL path (pynq_inference.py:414-425) and
train_linear_classifier on GPU features (retrain_classifier.py:24-124).

load_image_any is integer arithmetic (Pillow's luma and fixed-point BICUBIC resampler): bit-exact against outputs of the
reference's own function (tests/golden/pil_cases.npz) and against the numpy restatement on more sizes.
The trainer is fp32 gradient descent: same start (numpy RandomState(42)), same operations, matrix products summed in a
different order -> weights within 2e-4 absolute of the reference's after hundreds of epochs, identical predictions.
"""
import os

import numpy as np
import pytest

import inputs
from oracle import np_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    import fpga_cnn_b200 as fc
    a = fc.CNNAccelerator()
    yield a
    a.close()


@pytest.mark.parametrize("case", inputs.PIL_CASES, ids=lambda c: c["name"])
def test_image_to_gray128_fixtures(case, acc, pil_golden):
    arr = inputs.make_pil_image(case)
    got = acc.image_to_gray128(arr[None])
    assert got.shape == (1, 128, 128) and np.array_equal(got.reshape(-1), pil_golden[case["name"]])


def test_image_to_gray128_batches_sizes_and_torch(acc):
    import torch
    rng = np.random.default_rng(5)
    for (h, w, c, n) in [(96, 160, 3, 5), (720, 1280, 3, 2), (131, 200, 1, 3), (128, 128, 4, 2), (2048, 1536, 1, 1), (64, 64, 3, 4)]:
        shape = (n, h, w) if c == 1 else (n, h, w, c)
        x = rng.integers(0, 256, shape, dtype=np.uint8)
        got = acc.image_to_gray128(x)
        for i in range(n):
            assert np.array_equal(got[i].reshape(-1), np_oracle.load_image_array(x[i])), (h, w, c, i)
        t = acc.image_to_gray128(torch.from_numpy(x).cuda())
        assert np.array_equal(t.cpu().numpy(), got)
    with pytest.raises(ValueError):
        acc.image_to_gray128(np.zeros((1, 10, 10, 2), np.uint8))


def test_load_image_any_files(tmp_path, acc, pil_golden):
    """The file-level surface: .bin passthrough with the size check, PNG / JPEG through PIL's decoder then the GPU."""
    import fpga_cnn_b200 as fc
    from PIL import Image
    raw = inputs.tb_image().reshape(-1)
    p = tmp_path / "img.bin"
    raw.tofile(p)
    assert np.array_equal(fc.load_image_any(str(p)), raw)
    (tmp_path / "short.bin").write_bytes(b"\x00" * 100)
    with pytest.raises(ValueError):
        fc.load_image_any(str(tmp_path / "short.bin"))
    for case in inputs.PIL_CASES[:4]:
        q = tmp_path / (case["name"] + ".png")
        Image.fromarray(inputs.make_pil_image(case), case["mode"]).save(q)
        assert np.array_equal(fc.load_image_any(str(q), acc), pil_golden[case["name"]])
    j = tmp_path / "photo.jpg"                      # lossy: compare with PIL itself on the same file
    Image.fromarray(inputs.make_pil_image(inputs.PIL_CASES[0]), "RGB").save(j, quality=90)
    want = np.array(Image.open(j).convert("L").resize((128, 128)), dtype=np.uint8).flatten()
    assert np.array_equal(fc.load_image_any(str(j), acc), want)
    pal = tmp_path / "pal.png"                      # palette mode: PIL converts to L first, the GPU resizes
    Image.fromarray(inputs.make_pil_image(inputs.PIL_CASES[0]), "RGB").convert("P").save(pal)
    want = np.array(Image.open(pal).convert("L").resize((128, 128)), dtype=np.uint8).flatten()
    assert np.array_equal(fc.load_image_any(str(pal), acc), want)


def test_train_linear_classifier_matches_reference_fixture(trainer_golden):
    import fpga_cnn_b200 as fc
    x, y = inputs.make_training_set()
    for tag, kw in (("e400", dict(lr=0.01, epochs=400)), ("e1000_lr05", dict(lr=0.05, epochs=1000))):
        W, b = fc.train_linear_classifier(x, y, 6, verbose=False, **kw)
        Wr, br = trainer_golden[tag + "_W"], trainer_golden[tag + "_b"]
        assert W.shape == (6, 1024) and W.dtype == np.float32 and b.shape == (6,)
        assert np.abs(W - Wr).max() <= 2e-4 and np.abs(b - br).max() <= 2e-4, (tag, np.abs(W - Wr).max(), np.abs(b - br).max())
        assert np.array_equal((x @ W.T + b).argmax(1), (x @ Wr.T + br).argmax(1))


def test_trained_classifier_runs_on_gpu_features(shipped_weights):
    """End of the reference's workflow (dump features -> retrain -> load fc): pooled GPU features of two kinds of images train
    a (2,1024) classifier that the tail kernel then applies to the same images."""
    import fpga_cnn_b200 as fc
    a = fc.CNNAccelerator()
    a.load_weights(shipped_weights)
    a.set_shifts(7, 10, 11)
    imgs = np.concatenate([inputs.make_images(("rng", 95), 60), inputs.make_images(("smooth", 96), 60)])
    labels = np.array([0] * 60 + [1] * 60)
    feats = a.run_batch(imgs).reshape(120, 64, 256)
    pooled = a.pool_features(feats)
    W, b = fc.train_linear_classifier(pooled, labels, 2, lr=0.05, epochs=200, verbose=False)
    a.load_classifier(W, b)
    cls, probs, _ = a.infer_batch(imgs)
    assert (cls == labels).mean() >= 0.95
    a.close()
