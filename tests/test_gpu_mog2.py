"""Background model on the GPU (ckb_mog2_apply, ckb_zone_fg_counts) against the oracle's cv2-pinned restatement and the
golden stream recorded from the reference's StonesFinder._learn_bg (stonesfinder.py:113-115,171-176)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    return StoneEngine(19)


def test_golden_background_stream(engine, golden):
    """warp + MOG2 + zone counts on the device, one frame per call as the plugin does: bit-identical masks."""
    g = golden("background_stream.npz")
    frames, mtx, init = g["frames"], g["mtx"], int(g["bg_init_frames"])
    state = engine.mog2_new_state()
    for i in range(frames.shape[0]):
        goban = engine.warp(torch.from_numpy(frames[i:i + 1]).cuda(), mtx)
        fg = engine.mog2_apply(goban, state, i, [0.01 if i < init else 0.005])
        assert np.array_equal(np.packbits(fg[0].cpu().numpy() > 0), g["masks"][i]), "frame %d" % i
        assert set(np.unique(fg.cpu().numpy())) <= {0, 255}
        assert np.array_equal(engine.zone_fg_counts(fg)[0].cpu().numpy(), g["zone_fg"][i])


def test_golden_background_batched(engine, golden):
    """The same stream in one call (all frames applied in order inside the kernel) and split 7 + 23."""
    g = golden("background_stream.npz")
    frames, mtx, init = g["frames"], g["mtx"], int(g["bg_init_frames"])
    n = frames.shape[0]
    goban = engine.warp(torch.from_numpy(frames).cuda(), mtx)
    rates = [0.01 if i < init else 0.005 for i in range(n)]
    state = engine.mog2_new_state()
    fg = engine.mog2_apply(goban, state, 0, rates).cpu().numpy()
    state2 = engine.mog2_new_state()
    fg2 = torch.cat([engine.mog2_apply(goban[:7], state2, 0, rates[:7]),
                     engine.mog2_apply(goban[7:], state2, 7, rates[7:])]).cpu().numpy()
    for i in range(n):
        assert np.array_equal(np.packbits(fg[i] > 0), g["masks"][i]), "frame %d" % i
    assert np.array_equal(fg, fg2)
    assert torch.equal(state, state2)
    assert np.array_equal(engine.zone_fg_counts(torch.from_numpy(fg).cuda()).cpu().numpy(), g["zone_fg"])


@pytest.mark.parametrize("rates", ["auto", "fast", "mixed"])
def test_noise_video_vs_oracle(engine, oracle, rates):
    """150 frames (more than one 64-frame launch) of a noisy canonical-size scene with objects appearing, moving and
    leaving; automatic (negative), large and mixed learning rates; state carried across calls."""
    rng = np.random.default_rng(17)
    S, n = 380, 150
    bg = rng.integers(0, 256, (S, S, 3)).astype(np.int16)
    frames = np.empty((n, S, S, 3), np.uint8)
    for i in range(n):
        f = bg + rng.integers(-10, 11, (S, S, 3))
        if i > 15:
            x = (i * 7) % (S - 60)
            f[100:160, x:x + 60] = rng.integers(0, 256, 3)
        if 40 < i < 100:
            f[200:260, 30:120] += 70
        frames[i] = np.clip(f, 0, 255).astype(np.uint8)
    lr = {"auto": [-1.0] * n, "fast": [0.25] * n, "mixed": [(-1.0, 0.01, 0.3, 0.0)[i % 4] for i in range(n)]}[rates]
    model = oracle.CMog2((S, S))
    want = np.stack([model.apply(frames[i], lr[i]) for i in range(n)])
    state = engine.mog2_new_state()
    d = torch.from_numpy(frames).cuda()
    got = torch.cat([engine.mog2_apply(d[:140], state, 0, lr[:140]), engine.mog2_apply(d[140:], state, 140, lr[140:])])
    got = got.cpu().numpy()
    assert np.array_equal(got, want), "%d pixels differ" % int((got != want).sum())
    # the device state equals the oracle's (planar on the device: [25][pixels] floats + [pixels] mode counts)
    st = state.cpu().numpy()
    planes = st[:25 * S * S * 4].view(np.float32).reshape(25, S * S)
    nm = st[25 * S * S * 4:25 * S * S * 4 + S * S]
    assert np.array_equal(nm, model.nmodes)
    live = np.arange(5)[None, :] < model.nmodes[:, None]                      # only modes in use are defined
    for k in range(5):
        assert np.array_equal(planes[k][live[:, k]], model.state[live[:, k], k])              # weights
        assert np.array_equal(planes[5 + k][live[:, k]], model.state[live[:, k], 5 + k])      # variances
        for c in range(3):
            assert np.array_equal(planes[10 + 3 * k + c][live[:, k]], model.state[live[:, k], 10 + 3 * k + c])


def test_arguments_and_reset(engine):
    from camkifu_b200._lib import CkbError
    S = 380
    img = torch.zeros((1, S, S, 3), dtype=torch.uint8, device="cuda")
    state = engine.mog2_new_state()
    with pytest.raises(CkbError):
        engine.mog2_apply(img, state, 0, [1.0])        # OpenCV re-initialises at rates >= 1: ask for a reset instead
    a = engine.mog2_apply(img, state, 0, [0.01])
    assert int(a.min()) == 255                          # a fresh model calls everything foreground
    b = engine.mog2_apply(img, state, 1, [0.01])
    assert int(b.max()) == 0
    assert engine.mog2_apply(img[:0], state, 2, []).shape[0] == 0
