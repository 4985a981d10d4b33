"""Seeded synthetic goban frames (host-side test/bench data; not on the product path).

Follows the generator spec of SURVEY.md §8(d): a low-frequency "room" background, a wooden board with dark grid lines
and black/white stones rendered in canonical space at 4x, projected into the camera frame by a known homography whose
four corners are an inset rectangle plus per-corner jitter. The board->frame homography is what a BoardFinder would
publish as `mtx` (reference: src/camkifu/board/boardfinder.py:30-31,43-45 — `cv2.getPerspectiveTransform(hull, dst)`
with dst = [(0,0),(S,0),(S,S),(0,S)] float32).
"""
import cv2
import numpy as np

E, B, W = 0, 1, 2  # uint8 codes used across this repo for empty / black / white


def random_stones(rng: np.random.Generator, gsize: int = 19, p=(0.5, 0.25, 0.25)) -> np.ndarray:
    """(gsize, gsize) uint8 codes, i.i.d. with P(E,B,W)=p (dense enough for sf_clustering.py:170-178)."""
    return rng.choice(np.array([E, B, W], dtype=np.uint8), size=(gsize, gsize), p=p)


def render_canonical(rng: np.random.Generator, stones: np.ndarray, scale: int = 4) -> np.ndarray:
    """Top-down board image of side 20*gsize*scale, uint8 BGR."""
    gsize = stones.shape[0]
    pitch = 20 * scale
    side = pitch * gsize
    wood = np.array([90, 160, 210], dtype=np.float32)
    tex = cv2.GaussianBlur(rng.normal(0.0, 20.0, (side // 4, side // 4)).astype(np.float32), (0, 0), 3)
    tex = cv2.resize(tex, (side, side), interpolation=cv2.INTER_LINEAR)
    img = wood[None, None, :] + tex[:, :, None]
    img = np.clip(img, 0, 255).astype(np.uint8)
    half = pitch // 2
    for k in range(gsize):
        p = half + k * pitch
        cv2.line(img, (half, p), (side - half, p), (30, 40, 50), max(1, scale // 2))
        cv2.line(img, (p, half), (p, side - half), (30, 40, 50), max(1, scale // 2))
    rad = int(0.45 * pitch)
    for r in range(gsize):
        for c in range(gsize):
            s = stones[r, c]
            if s == E:
                continue
            ctr = (half + c * pitch, half + r * pitch)
            if s == B:
                cv2.circle(img, ctr, rad, (25, 25, 25), -1, cv2.LINE_AA)
                cv2.circle(img, (ctr[0] - rad // 3, ctr[1] - rad // 3), rad // 5, (70, 70, 70), -1, cv2.LINE_AA)
            else:
                cv2.circle(img, ctr, rad, (235, 235, 235), -1, cv2.LINE_AA)
                cv2.circle(img, (ctr[0] - rad // 3, ctr[1] - rad // 3), rad // 5, (255, 255, 255), -1, cv2.LINE_AA)
    return img


def random_corners(rng: np.random.Generator, H: int, W: int, margin: float = 0.15, jitter: float = 0.03) -> np.ndarray:
    """4x2 float32 (x, y) corners ordered TL, TR, BR, BL: inset rectangle + N(0, jitter*min(H,W)) per corner."""
    x0, x1 = margin * W, (1 - margin) * W
    y0, y1 = margin * H, (1 - margin) * H
    base = np.array([(x0, y0), (x1, y0), (x1, y1), (x0, y1)], dtype=np.float64)
    base += rng.normal(0.0, jitter * min(H, W), base.shape)
    return base.astype(np.float32)


def board_homography(corners: np.ndarray, canonical_size: int) -> np.ndarray:
    """The `mtx` a BoardFinder publishes: frame -> canonical, 3x3 float64 (boardfinder.py:43-45)."""
    S = canonical_size
    dst = np.array([(0, 0), (S, 0), (S, S), (0, S)], dtype=np.float32)
    return cv2.getPerspectiveTransform(np.asarray(corners, dtype=np.float32), dst)


def make_background(rng: np.random.Generator, H: int, W: int) -> np.ndarray:
    small = rng.normal(110.0, 60.0, (H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    small = cv2.GaussianBlur(small, (0, 0), 2.0)
    bg = cv2.resize(small, (W, H), interpolation=cv2.INTER_CUBIC)
    return bg


def render_frame(rng: np.random.Generator, H: int, W: int, stones: np.ndarray, corners: np.ndarray,
                 noise: int = 8, background: np.ndarray = None) -> np.ndarray:
    """Camera frame (H, W, 3) uint8 BGR, C-contiguous."""
    gsize = stones.shape[0]
    scale = 4
    board = render_canonical(rng, stones, scale)
    side = board.shape[0]
    src = np.array([(0, 0), (side, 0), (side, side), (0, side)], dtype=np.float32)
    Mb = cv2.getPerspectiveTransform(src, np.asarray(corners, dtype=np.float32))
    bg = make_background(rng, H, W) if background is None else background
    proj = cv2.warpPerspective(board, Mb, (W, H), flags=cv2.INTER_AREA)
    mask = cv2.warpPerspective(np.full((side, side), 255, np.uint8), Mb, (W, H), flags=cv2.INTER_NEAREST)
    frame = np.where(mask[:, :, None] > 0, proj.astype(np.float32), bg)
    if noise:
        frame = frame + rng.integers(-noise, noise + 1, frame.shape).astype(np.float32)
    return np.ascontiguousarray(np.clip(frame, 0, 255).astype(np.uint8))


def make_clip(seed: int, n: int, H: int, W: int, gsize: int = 19, new_board_every: int = 1):
    """n frames of one 'video': fixed corners, a new random position every `new_board_every` frames.

    Returns (frames uint8 [n,H,W,3], mtx float64 [3,3], truth uint8 [n,gsize,gsize], corners float32 [4,2]).
    """
    rng = np.random.default_rng(seed)
    corners = random_corners(rng, H, W)
    mtx = board_homography(corners, 20 * gsize)
    frames = np.empty((n, H, W, 3), dtype=np.uint8)
    truth = np.empty((n, gsize, gsize), dtype=np.uint8)
    bg = make_background(rng, H, W)
    stones = None
    for i in range(n):
        if stones is None or i % new_board_every == 0:
            stones = random_stones(rng, gsize)
        frames[i] = render_frame(rng, H, W, stones, corners, background=bg)
        truth[i] = stones
    return frames, mtx, truth, corners


def make_clip_parallel(seed: int, n: int, H: int, W: int, gsize: int = 19, workers: int = 0):
    """Same kind of clip as make_clip (fixed corners and background, a new random position every frame), rendered by a
    thread pool with one child generator per frame. Returns (frames, mtx, truth, corners)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    root = np.random.SeedSequence(seed)
    rng0 = np.random.default_rng(root.spawn(1)[0])
    corners = random_corners(rng0, H, W)
    mtx = board_homography(corners, 20 * gsize)
    bg = make_background(rng0, H, W)
    frames = np.empty((n, H, W, 3), dtype=np.uint8)
    truth = np.empty((n, gsize, gsize), dtype=np.uint8)
    children = root.spawn(n + 1)[1:]

    def one(i):
        rng = np.random.default_rng(children[i])
        stones = random_stones(rng, gsize)
        frames[i] = render_frame(rng, H, W, stones, corners, background=bg)
        truth[i] = stones

    with ThreadPoolExecutor(max_workers=workers or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(n)))
    return frames, mtx, truth, corners


def make_game_clip(seed: int, n: int, H: int, W: int, events=(), gsize: int = 19, p=(0.7, 0.15, 0.15), reach: int = 4):
    """A fixed-camera 'game' clip: one starting position, and for every event (frame, colour code, r, c) a hand-coloured
    blob that travels from the lower board edge to intersection (r, c) during `reach` frames, leaves a stone there and
    withdraws during the next `reach` frames. This is the kind of input the background model (MOG2) and SfNeural's
    foreground-driven targeting react to. Returns (frames [n,H,W,3], mtx, truth [n,g,g], corners)."""
    rng = np.random.default_rng(seed)
    corners = random_corners(rng, H, W)
    S = 20 * gsize
    mtx = board_homography(corners, S)
    inv = np.linalg.inv(mtx)
    bg = make_background(rng, H, W)
    stones = random_stones(rng, gsize, p)
    for (_, _, r, c) in events:
        stones[r, c] = E
    frames = np.empty((n, H, W, 3), dtype=np.uint8)
    truth = np.empty((n, gsize, gsize), dtype=np.uint8)

    def to_frame(x, y):       # canonical (x = column, y = row) -> frame pixel
        v = inv @ np.array([x, y, 1.0])
        return v[0] / v[2], v[1] / v[2]

    board_rng_state = rng.bit_generator.state
    for i in range(n):
        hand = None
        for (f0, color, r, c) in events:
            if i >= f0 + reach:
                stones[r, c] = color
            if f0 <= i < f0 + 2 * reach:
                t = (i - f0 + 1) / reach if i < f0 + reach else (f0 + 2 * reach - i - 1) / reach
                tx, ty = 10 + 20 * c, 10 + 20 * r
                sx, sy = tx, S + 60.0           # starts below the board
                hand = (sx + (tx - sx) * t, sy + (ty - sy) * t)
        # the board texture is part of the scene, not of the frame: render it from the same generator state each time
        rng.bit_generator.state = board_rng_state
        frame = render_frame(rng, H, W, stones, corners, noise=0, background=bg).astype(np.float32)
        if hand is not None:
            cx, cy = to_frame(*hand)
            ex, ey = to_frame(hand[0], hand[1] + 140.0)
            blob = np.zeros((H, W), np.uint8)
            rad = max(3, int(0.035 * min(H, W)))
            cv2.line(blob, (int(cx), int(cy)), (int(ex), int(ey)), 255, 2 * rad, cv2.LINE_AA)
            cv2.circle(blob, (int(cx), int(cy)), rad, 255, -1, cv2.LINE_AA)
            a = (blob.astype(np.float32) / 255.0)[:, :, None]
            frame = frame * (1 - a) + a * np.array([120, 150, 205], np.float32)
        noise_rng = np.random.default_rng([seed, i])
        frame += noise_rng.integers(-6, 7, frame.shape).astype(np.float32)
        frames[i] = np.clip(frame, 0, 255).astype(np.uint8)
        truth[i] = stones
    return frames, mtx, truth, corners


def wild_homographies(rng, n):
    """Homographies far from a board finder's: the horizon inside the canonical square (the projective denominator
    changes sign), singular matrices (cv::invert yields zeros), scales that saturate the fixed-point coordinates,
    random dense matrices."""
    out = []
    for t in range(n):
        kind = t % 4
        if kind == 0:
            M = np.array([[1.0, 0.1, 5], [0.05, 1.2, -3], [rng.normal(0, 4e-3), rng.normal(0, 4e-3), 1.0]])
        elif kind == 1:
            M = np.array([[1.0, 2, 3], [2, 4, 6], [0.5, 1, 1.5]]) * rng.normal()
        elif kind == 2:
            M = np.diag([1e-4, 1e-4, 1.0]) @ np.array([[1, 0.2, 0], [0.1, 1, 0], [0, 0, 1.0]])
        else:
            M = rng.normal(size=(3, 3)) * np.array([[1, 1, 100], [1, 1, 100], [1e-3, 1e-3, 1]])
        out.append(M)
    return out
