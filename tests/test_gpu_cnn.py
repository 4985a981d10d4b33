"""K4 parity on the GPU: the SfNeural forward pass against the fp32 oracle (tolerance stated below) and the decode
against the reference's golden vectors (bit-exact given the same softmax)."""
import numpy as np
import pytest
import torch

from camkifu_b200 import synth, weights

pytestmark = pytest.mark.gpu

# north_star: CNN softmax outputs within 1e-3 relative. Softmax vectors are compared relative to their scale
# (max |y| = the winning probability): |y_gpu - y_oracle| <= 1e-3 * max(y_oracle), and the label (argmax) exactly.
SOFTMAX_RTOL = 1e-3


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    e = StoneEngine(19)
    e.set_cnn_weights(weights.glorot_params(seed=0, bias_scale=0.5))
    return e


def boards(n, seed=1):
    import cv2
    frames, M, truth, _ = synth.make_clip(seed, n, 240, 320)
    return np.stack([cv2.warpPerspective(f, M, (380, 380)) for f in frames])


def check(out, goban, oracle, params):
    n = goban.shape[0]
    for k in range(n):
        xs = oracle.c_nn_gather(goban[k])
        y = oracle.c_cnn_forward(xs, params)
        yg = out["softmax"][k].cpu().numpy()
        err = np.abs(yg - y).max(axis=1) / y.max(axis=1)
        assert err.max() <= SOFTMAX_RTOL, "softmax error %.3g" % err.max()
        assert np.array_equal(yg.argmax(1), y.argmax(1))
        stones, conf, keep = oracle.c_nn_decode(yg)           # decode of the SAME softmax must be bit-exact
        assert np.array_equal(out["stones"][k].cpu().numpy(), stones)
        assert np.array_equal(out["conf"][k].cpu().numpy(), conf)
        assert np.array_equal(out["keep"][k].cpu().numpy().astype(bool), keep)
        s2, c2, k2 = oracle.c_nn_decode(y)                    # and agree with the oracle's own end result
        assert np.array_equal(stones, s2) and np.allclose(conf, c2, rtol=1e-3)
        far = np.abs(c2 - 0.6) > 1e-3
        assert np.array_equal(keep[far], k2[far])


def test_simt_forward_vs_oracle(engine, oracle):
    goban = boards(2)
    out = engine.cnn_forward(torch.from_numpy(goban).cuda(), simt=True)
    check(out, goban, oracle, weights.glorot_params(seed=0, bias_scale=0.5))


def test_decode_golden(engine, golden, oracle):
    """Feed the golden softmax through the device decode: NNCache.predict_all_stones / SfNeural.predict_all."""
    g = golden("neural_decode.npz")
    stones, conf, keep = oracle.c_nn_decode(g["y"])
    assert np.array_equal(stones, g["stones"]) and np.array_equal(conf, g["conf"])


def test_tc_layers_vs_oracle(engine, oracle):
    """Each tensor-core layer against the oracle's float32 activations (bf16 hi/lo split operands, 3 products):
    relative to the layer's largest activation the error must stay below 1e-4."""
    goban = boards(1, seed=5)
    params = weights.glorot_params(seed=0, bias_scale=0.5)
    xs = oracle.c_nn_gather(goban[0])
    y, acts = oracle.c_cnn_forward(xs, params, want_acts=True)
    sizes = [("conv1", 1, 36 * 36 * 32), ("pool2", 2, 16 * 16 * 32), ("conv3", 3, 14 * 14 * 90), ("pool4", 4, 3240),
             ("fc1", 5, 160)]
    engine.cnn_set_debug(True)      # conv1's activations normally stay in shared memory (fused front kernel)
    try:
        engine.cnn_forward(torch.from_numpy(goban).cuda())
        off = 0
        for name, layer, sz in sizes:
            ref = acts[:, off:off + sz]
            off += sz
            got = engine.cnn_debug_activation(1, layer).cpu().numpy().reshape(100, -1)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err < 1e-4, "%s: relative error %.3g" % (name, err)
    finally:
        engine.cnn_set_debug(False)


@pytest.mark.parametrize("n", [1, 3])
def test_tc_forward_vs_oracle(engine, oracle, n):
    goban = boards(n, seed=2 + n)
    out = engine.cnn_forward(torch.from_numpy(goban).cuda())
    check(out, goban, oracle, weights.glorot_params(seed=0, bias_scale=0.5))


def test_tc_matches_simt_on_noise(engine):
    """Tensor-core path vs the fp32 CUDA-core path on white-noise images (largest activations, no structure)."""
    rng = np.random.default_rng(11)
    goban = torch.from_numpy(rng.integers(0, 256, (2, 380, 380, 3), dtype=np.uint8)).cuda()
    a = engine.cnn_forward(goban)["softmax"].cpu().numpy()
    b = engine.cnn_forward(goban, simt=True)["softmax"].cpu().numpy()
    err = np.abs(a - b).max(axis=2) / b.max(axis=2)
    assert err.max() <= SOFTMAX_RTOL, "softmax error %.3g" % err.max()


@pytest.mark.parametrize("n", [5, 17, 64])
def test_tc_pool_fused_over_tile_ranges(engine, n):
    """conv4's epilogue pools across tile boundaries (every CTA takes a contiguous range of 128-pixel tiles and recomputes
    the tile in front of it): batch sizes that split the tiles over the CTAs differently — fewer tiles than CTAs + 1, an
    uneven split, the full 64-frame pass — against the fp32 CUDA-core path, every patch."""
    rng = np.random.default_rng(100 + n)
    base = boards(2, seed=9)
    goban = np.empty((n, 380, 380, 3), np.uint8)
    for k in range(n):
        noise = rng.integers(-40, 41, (380, 380, 3))
        goban[k] = np.clip(np.roll(base[k % 2], (7 * k) % 380, axis=(k % 2)).astype(np.int64) + noise, 0, 255)
    g = torch.from_numpy(goban).cuda()
    a = engine.cnn_forward(g)
    b = engine.cnn_forward(g, simt=True)
    ya, yb = a["softmax"].cpu().numpy(), b["softmax"].cpu().numpy()
    err = np.abs(ya - yb).max(axis=2) / yb.max(axis=2)
    assert err.max() <= SOFTMAX_RTOL, "softmax error %.3g" % err.max()
    top2 = np.sort(yb, axis=2)[:, :, -2:]
    clear = (top2[:, :, 1] - top2[:, :, 0]) > 2e-3 * top2[:, :, 1]      # ties within the tolerance may flip
    assert np.array_equal(ya.argmax(2)[clear], yb.argmax(2)[clear])
    assert clear.mean() > 0.99


def test_tc_forward_trained_weights(oracle, golden):
    """Realistic weights (tests/golden/sfneural_trained.npz: the reference architecture trained on synthetic boards by
    oracle/train_fixture.py — the reference's own model does not ship): peaked softmax outputs, large logits. Same bar:
    1e-3 relative on the softmax, identical labels and decode; and the network reads the synthetic boards correctly."""
    import cv2
    from camkifu_b200.engine import StoneEngine
    params = golden("sfneural_trained.npz")["params"]
    eng = StoneEngine(19)
    eng.set_cnn_weights(params)
    frames, M, truth, _ = synth.make_clip(12, 3, 360, 480)
    goban = np.stack([cv2.warpPerspective(f, M, (380, 380)) for f in frames])
    out = eng.cnn_forward(torch.from_numpy(goban).cuda())
    check(out, goban, oracle, params)
    assert np.array_equal(out["stones"].cpu().numpy(), truth)
    assert float(out["conf"].min()) > 0.9
    simt = eng.cnn_forward(torch.from_numpy(goban).cuda(), simt=True)
    assert np.array_equal(simt["stones"].cpu().numpy(), truth)
