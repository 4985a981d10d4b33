"""The drop-in plugins on the GPU, driven the way VidProcessor.execute drives a finder, against the reference's golden
vectors (recorded by running the unmodified reference, oracle/gen_golden.py) and against the oracle."""
import numpy as np
import pytest
import torch

from camkifu_b200 import plugins, synth, weights
from camkifu_b200.harness import HeadlessVManager, run_frames

pytestmark = pytest.mark.gpu
CODE = {'E': 0, 'B': 1, 'W': 2}


def codes(arr):
    return np.vectorize(CODE.get)(arr).astype(np.uint8)


def test_sfclustering_stream_matches_reference(golden):
    """7 frames through _doframe -> _find: canonical image, running average, the bulk moves the controller received and
    the final board are those of the reference's SfClustering (clustering_stream.npz)."""
    g = golden("clustering_stream.npz")
    vm = HeadlessVManager(g["mtx"])
    sf = plugins.SfClusteringB200(vm)
    for i in range(7):
        sf.set_rng_seed(int(g["seeds"][i]))          # the golden run did cv2.setRNGSeed(seed) before every frame
        n0 = len(vm.controller.piped)
        sf._doframe(g["frames"][i].copy())
        sf.total_f_processed += 1
        assert np.array_equal(sf.goban_img, g["goban"][i])
        bulk = [a for ins, a in vm.controller.piped[n0:] if ins == "bulk"]
        got = np.array([(CODE[m.color], m.y, m.x) for m in bulk[0][0]], np.int32) if bulk else np.zeros((0, 3), np.int32)
        assert np.array_equal(got, g["moves_%d" % i]), "frame %d" % i
        if i == 1:
            assert np.array_equal(sf.accu, g["accu_1"])
    assert np.array_equal(sf.accu, g["accu_last"])
    assert np.array_equal(codes(vm.controller.stones), g["board"])


def test_sfclustering_find_stones_delegate(golden):
    """The SfMeta delegate contract: find_stones(img, rs, re, cs, ce, **kwargs) -> object array of E/B/W, or None."""
    g = golden("clustering_full.npz")
    sf = plugins.SfClusteringB200(None)              # SfMeta passes vmanager=None (sf_meta.py:52)
    for k in range(3):
        sf.set_rng_seed(int(g["seeds"][k]))
        st = sf.find_stones(g["goban_%d" % k])
        assert st.dtype == object and st.shape == (19, 19)
        assert np.array_equal(codes(st), g["stones_%d" % k])
        assert all(v is plugins.E or v is plugins.B or v is plugins.W for v in st.ravel())
    sf.set_rng_seed(int(g["seeds"][3]))
    assert sf.find_stones(g["goban_sparse"]) is None  # density check failed in the reference too
    sf.set_rng_seed(int(g["seeds"][4]))
    st = sf.find_stones(g["goban_0"], rs=6, re=13, cs=12, ce=19, canvas=None)
    if g["stones_region"][0, 0] != 255:
        assert np.array_equal(codes(st), g["stones_region"])
    # float32 input (what SfMeta hands over after its own averaging) takes the same path
    sf.set_rng_seed(int(g["seeds"][0]))
    st = sf.find_stones(g["goban_0"].astype(np.float32))
    assert np.array_equal(codes(st), g["stones_0"])


def test_sfneural_predict_all_matches_oracle(oracle):
    """Still image (bg_init_frames = 0): frame 0 loads the net, frame 1 runs predict_all -> one bulk update with every
    non-empty intersection seen at confidence > 0.6 (sf_neural.py:57-70)."""
    frames, mtx, truth, _ = synth.make_clip(8, 3, 240, 320, new_board_every=100)
    params = weights.glorot_params(seed=0, bias_scale=0.5)
    vm = HeadlessVManager(mtx, video="snapshot.png")
    plugins.SfNeuralB200.cnn_params = params
    try:
        sf = plugins.SfNeuralB200(vm)
        ctl = run_frames(sf, frames[:2])
    finally:
        plugins.SfNeuralB200.cnn_params = None
    assert sf.has_sampled and len(ctl.bulk_moves()) == 1
    y = oracle.c_cnn_forward(oracle.c_nn_gather(sf.goban_img), params)
    stones, conf, keep = oracle.c_nn_decode(y)
    sure = np.abs(conf - 0.6) > 1e-3                  # intersections whose keep flag is not within tolerance of the rule
    got = {(r, c): col for col, r, c in ctl.bulk_moves()[0]}
    for r in range(19):
        for c in range(19):
            if sure[r, c]:
                assert ((r, c) in got) == bool(keep[r, c])
                if keep[r, c]:
                    assert CODE[got[(r, c)]] == stones[r, c]
    # cache API of the reference's NNCache
    sq, cf = sf.cache.predict_4_stones(3, 4)
    assert np.array_equal(codes(sq), stones[6:8, 8:10]) and abs(cf - conf[6, 8]) < 1e-3
    # steady state: a third frame showing the same position adds nothing contradictory
    run_frames(sf, frames[2:3])
    board = codes(ctl.stones)
    assert np.array_equal(board[sure & keep], stones[sure & keep])


def test_detect_pipeline_matches_engine(oracle):
    from camkifu_b200.engine import StoneEngine, rng_seed, rng_advance
    from camkifu_b200.pipeline import DetectPipeline, pinned_frames
    H, W, n = 360, 480, 21
    frames, mtx, truth, _ = synth.make_clip(3, n, H, W)
    params = weights.glorot_params(seed=0)
    pipe = DetectPipeline(H, W, mode="both", sub_batch=8, cnn_params=params)
    host = pinned_frames(n, H, W)
    host.copy_(torch.from_numpy(frames))
    st0 = rng_seed(9)
    res = {k: v.copy() for k, v in pipe.detect(host, mtx, rng_state=st0).items()}
    assert pipe.h2d_bytes < n * H * W * 3            # only the board's bounding box travels
    res_full = pipe.detect(frames, mtx, rng_state=st0, crop=False)   # pageable numpy input, whole frames
    for k in res:
        assert np.array_equal(res[k], res_full[k]), k
    eng = pipe.eng
    goban = eng.warp(torch.from_numpy(frames).cuda(), mtx)
    ref = eng.cnn_forward(goban)
    km = eng.find_stones(goban, [rng_advance(st0, i) for i in range(n)])
    assert np.array_equal(res["stones"], ref["stones"].cpu().numpy())
    assert np.array_equal(res["keep"], ref["keep"].cpu().numpy())
    assert np.array_equal(res["km_stones"], km["stones"].cpu().numpy())
    assert np.array_equal(res["km_stones"], truth)   # the k-means path reads these synthetic boards perfectly
    g0 = goban[0].cpu().numpy()
    assert np.array_equal(g0, oracle.c_warp(frames[0], mtx, 380))


def test_detect_stream_matches_synchronous_calls():
    """submit/collect pipelining (batch k+1 uploads while batch k computes) returns the same board states, in order,
    as one synchronous detect() per batch; different segments (homographies) may follow each other."""
    from camkifu_b200.pipeline import DetectPipeline, pinned_frames
    H, W = 360, 480
    params = weights.glorot_params(seed=0)
    pipe = DetectPipeline(H, W, mode="neural", sub_batch=4, cnn_params=params)
    batches = []
    for seed, n in ((4, 9), (5, 4), (6, 13), (7, 1), (8, 8)):
        frames, mtx, _, _ = synth.make_clip(seed, n, H, W)
        host = pinned_frames(n, H, W)
        host.copy_(torch.from_numpy(frames))
        batches.append((host, mtx))
    sync = [{k: v.copy() for k, v in pipe.detect(h, m).items()} for h, m in batches]
    got = [{k: v.copy() for k, v in r.items()} for r in pipe.detect_stream(iter(batches), depth=2)]
    assert len(got) == len(sync)
    for a, b in zip(got, sync):
        for k in b:
            assert np.array_equal(a[k], b[k]), k
    with pytest.raises(ValueError):
        t0 = pipe.submit(*batches[0])
        for h, m in batches[1:4]:
            pipe.submit(h, m)
        pipe.collect(t0)           # its slot has been reused: at most DEPTH batches may be outstanding


def test_sfneural_stream_matches_reference(golden):
    """neural_stream.npz: the unmodified SfNeural._find over a 72-frame 'game' clip (background sampling, initial
    assessment, foreground-driven targeting, look-back), with net.predict = the oracle's float32 forward on the trained
    fixture weights. The plugin — MOG2 and the CNN on the device — must send the controller the same stone updates on the
    same frames (the initial assessment, the look-back cancellations, the three stones a hand plays during the clip),
    and carry the same targets / heat-map state. (Within one frame the reference iterates a Python set: order is not
    defined.)"""
    g = golden("neural_stream.npz")
    frames, mtx, log = g["frames"], g["mtx"], g["log"]
    vm = HeadlessVManager(mtx)
    plugins.SfNeuralB200.cnn_params = golden("sfneural_trained.npz")["params"]
    try:
        sf = plugins.SfNeuralB200(vm)
        sf.bg_init_frames = int(g["bg_init_frames"])
        for i in range(frames.shape[0]):
            n0 = len(vm.controller.piped)
            sf._doframe(frames[i].copy())
            sf.total_f_processed += 1
            got = set()
            for ins, a in vm.controller.piped[n0:]:
                if ins == "bulk":
                    got |= {(0, CODE[m.color], m.y, m.x) for m in a[0]}
                elif ins == "append":
                    got.add((1, CODE[a[0].color], a[0].y, a[0].x))
                elif ins == "delete":
                    got.add((2, 0, a[1], a[0]))
            want = {tuple(int(v) for v in row[1:]) for row in log[log[:, 0] == i]}
            assert got == want, "frame %d: %s vs %s" % (i, sorted(got ^ want)[:6], len(want))
            assert np.array_equal(sf.targets, g["targets"][i]), "targets after frame %d" % i
            heat = np.array([[0 if h is None else h.energy + 100 for h in row] for row in sf.heatmap], np.int32)
            assert np.array_equal(heat, g["heat"][i]), "heat map after frame %d" % i
    finally:
        plugins.SfNeuralB200.cnn_params = None
    assert np.array_equal(codes(vm.controller.stones), g["board"])
    assert len(log) > 100 and g["targets"].max() > plugins.TARGET_THRESH     # the clip exercised the steady state
    assert (log[:, 1] == 1).sum() == 3                                       # three stones were suggested one by one
    for (_, color, r, c) in g["events"]:
        assert codes(vm.controller.stones)[r, c] == color                    # ... the ones the hand played


def test_process_video_file_sharded(tmp_path):
    """Offline video path: a lossless file is decoded into the pinned ring, streamed through the pipeline ("both" modes),
    and the concatenation of two ranks' shards equals the single-process result and the direct engine calls."""
    import cv2
    from camkifu_b200.engine import StoneEngine, rng_seed, rng_advance
    from camkifu_b200.video import process_video
    n, H, W = 41, 240, 320
    frames, mtx, truth, _ = synth.make_clip(31, n, H, W)
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30, (W, H))
    for f in frames:
        wr.write(f)
    wr.release()
    eng = StoneEngine(19)
    eng.set_cnn_weights(weights.glorot_params(seed=0))
    st0 = rng_seed(4)
    whole = process_video(path, mtx, mode="both", batch=8, engine=eng, rng_state=st0)
    assert whole["stones"].shape == (n, 19, 19) and whole["km_trusted"].shape == (n,)
    assert np.array_equal(whole["km_stones"], truth)
    goban = eng.warp(torch.from_numpy(frames).cuda(), mtx)
    nn = eng.cnn_forward(goban)
    km = eng.find_stones(goban, [rng_advance(st0, i) for i in range(n)])
    assert np.array_equal(whole["stones"], nn["stones"].cpu().numpy())
    assert np.array_equal(whole["conf"], nn["conf"].cpu().numpy())
    assert np.array_equal(whole["km_stones"], km["stones"].cpu().numpy())
    parts = [process_video(path, mtx, mode="both", batch=8, engine=eng, rng_state=st0, rank=r, world=3, gather=False)
             for r in range(3)]
    for k in whole:
        assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), k
    # an in-memory array is a valid source too
    mem = process_video(frames, mtx, mode="neural", batch=16, engine=eng)
    assert np.array_equal(mem["stones"], whole["stones"])
    # several decoder threads per rank: batches arrive out of order, the result does not change
    multi = process_video(path, mtx, mode="both", batch=4, engine=eng, rng_state=st0, decoders=3)
    for k in whole:
        assert np.array_equal(multi[k], whole[k]), k
    # a pinned ring standing for a long video (zero-copy source)
    from camkifu_b200.video import RingClip
    from camkifu_b200.pipeline import pinned_frames
    ring = pinned_frames(8, H, W)
    ring.copy_(torch.from_numpy(frames[:8]))
    long = process_video(RingClip(ring, 29), mtx, mode="neural", batch=8, engine=eng)
    assert np.array_equal(long["stones"], whole["stones"][np.arange(29) % 8])


def test_sfclustering_keeps_a_background_model(golden):
    """Like the reference's SfClustering (StonesFinder.__init__ default learn_bg=True), the plugin maintains the MOG2
    model on every frame: get_foreground() returns the mask cv2 would have produced, bg_init_frames exists."""
    g = golden("background_stream.npz")
    vm = HeadlessVManager(g["mtx"])
    sf = plugins.SfClusteringB200(vm)
    assert sf.bg_init_frames == 50
    sf.bg_init_frames = int(g["bg_init_frames"])
    for i in range(12):
        sf._doframe(g["frames"][i].copy())
        sf.total_f_processed += 1
        fg = sf.get_foreground()
        assert fg.shape == (380, 380) and np.array_equal(np.packbits(fg > 0), g["masks"][i]), "frame %d" % i
    assert sf.is_agitated(*np.unravel_index(int(g["zone_fg"][11].argmax()), (19, 19)))


def test_sfclustering_regions_delegate_equals_serial_calls(golden):
    """SfMeta's nine regions on one frame: the batched delegate call returns what nine find_stones calls return, in the
    same RNG order, and leaves the finder's RNG state where the serial calls leave it."""
    from camkifu_b200 import meta
    g = golden("clustering_full.npz")
    regions = meta.subregions()
    a, b = plugins.SfClusteringB200(None), plugins.SfClusteringB200(None)
    a.set_rng_seed(77)
    b.set_rng_seed(77)
    serial = [a.find_stones(g["goban_0"], *r) for r in regions]
    batched = b.find_stones_regions(g["goban_0"], regions)
    assert a.rng_state == b.rng_state
    for s, t in zip(serial, batched):
        assert (s is None) == (t is None)
        if s is not None:
            assert np.array_equal(codes(s), codes(t))
    assert any(s is not None for s in serial)


def test_nvjpeg_ingest_matches_host_decode(tmp_path):
    """Motion-JPEG ingest on the device (process_video(ingest="nvjpeg"), csrc/jpeg_ingest.cu): the frames nvJPEG decodes
    differ from OpenCV / FFmpeg's decode of the same file by a few levels at most, and the k-means board states of the
    two ingests are identical."""
    import cv2
    from camkifu_b200.engine import StoneEngine, rng_seed
    from camkifu_b200.video import MjpegAvi, process_video
    n, H, W = 24, 480, 640
    frames, mtx, truth, _ = synth.make_clip(13, n, H, W)
    path = str(tmp_path / "clip_mjpg.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30, (W, H))
    for f in frames:
        wr.write(f)
    wr.release()
    avi = MjpegAvi(path)
    assert (len(avi), avi.H, avi.W) == (n, H, W)
    eng = StoneEngine(19)
    if eng.jpeg_backend() == "unavailable":
        pytest.skip("libnvjpeg is not present on this machine")
    d = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    eng.jpeg_decode(avi.base_address, avi.offsets, avi.sizes, d)
    cap = cv2.VideoCapture(path)
    host = np.stack([cap.read()[1] for _ in range(n)])
    cap.release()
    diff = np.abs(d.cpu().numpy().astype(np.int16) - host.astype(np.int16))
    assert diff.max() <= 6 and diff.mean() < 0.5
    st0 = rng_seed(2)
    a = process_video(path, mtx, mode="clustering", batch=8, engine=eng, rng_state=st0)
    b = process_video(path, mtx, mode="clustering", batch=8, engine=eng, rng_state=st0, ingest="nvjpeg")
    assert np.array_equal(a["km_stones"], b["km_stones"]) and np.array_equal(a["km_stones"], truth)
    assert np.array_equal(a["km_trusted"], b["km_trusted"])
    c = process_video(avi.repeat(2), mtx, mode="clustering", batch=16, engine=eng, rng_state=st0, ingest="nvjpeg")
    assert c["km_stones"].shape[0] == 2 * n and np.array_equal(c["km_stones"][:n], b["km_stones"])
    # several decoder lanes (threads + streams) and more batches than ring slots: same results, in frame order
    eng.set_cnn_weights(weights.glorot_params(seed=0))
    e = process_video(avi.repeat(4), mtx, mode="both", batch=5, engine=eng, rng_state=st0, ingest="nvjpeg", decoders=3)
    one = process_video(avi.repeat(4), mtx, mode="both", batch=32, engine=eng, rng_state=st0, ingest="nvjpeg", decoders=1)
    for k in ("km_stones", "km_trusted", "stones", "keep"):
        assert np.array_equal(e[k], one[k]), k
    assert e["km_stones"].shape[0] == 4 * n and np.array_equal(e["km_stones"][n:2 * n], b["km_stones"])
    d2 = torch.empty_like(d)
    eng.jpeg_decode(avi.base_address, avi.offsets, avi.sizes, d2, lane=5)
    assert torch.equal(d, d2)
    with pytest.raises(Exception):
        eng.jpeg_decode(avi.base_address, avi.offsets, avi.sizes, d2, lane=eng.JPEG_LANES)
