"""Host side of SfMeta's region bookkeeping on top of the device statistics.

SfMeta (src/camkifu/stone/sf_meta.py) splits the goban into 3 x 3 Regions; each frame every Region (i) looks at the
foreground mask around and inside itself (`check_foreground` :342-381), (ii) runs one of two delegates on its part of the
image — `SfClustering.find_stones(img, rs, re, cs, ce)` :245-262 or `SfContours.find_stones` — and (iii) submits the
stones that recur in a short history (`commit` :305-340). The arithmetic of all three is on the device here:

  * per-zone foreground counts: ckb_zone_fg_counts (one launch per frame batch) -> `check_foreground` below needs only
    those 361 integers, no mask pixels;
  * the nine k-means of a frame: ckb_find_stones_regions, one set of launches for all regions x frames
    (`StoneEngine.find_stones_regions`), instead of nine serial find_stones calls;
  * SfContours' per-zone mean colours: ckb_zone_means (the contour extraction that produces its mask stays on the CPU);
  * the history vote: ckb_history_vote.

Contours, constraint checks (check_against, check_flow, check_lines, ...) and the Region state machine are host logic of
the reference and are not rebuilt (SURVEY.md section 2 marks them out of scope).
"""
import math

import numpy as np


def subregions(gsize: int = 19, split: int = 3):
    """SfMeta.subregion (sf_meta.py:101-125) for every (row, col), row-major: [(rs, re, cs, ce)]."""
    step = int(gsize / split)
    out = []
    for row in range(split):
        for col in range(split):
            re, ce = (row + 1) * step, (col + 1) * step
            if gsize - re < step:
                re = gsize
            if gsize - ce < step:
                ce = gsize
            out.append((row * step, re, col * step, ce))
    return out


def outer_border(rs: int, re: int, cs: int, ce: int, gsize: int = 19):
    """Region.outer_border (sf_meta.py:444-463): zones (row, col) around the region, in the reference's order (a zone may
    appear twice at the goban's edges, as in the reference: it then counts twice)."""
    out = []
    x = max(0, cs - 1)
    out += [(y, x) for y in range(max(0, rs - 1), min(gsize, re + 1))]
    y = min(gsize - 1, re)
    out += [(y, x) for x in range(max(1, cs), min(gsize, ce + 1))]
    x = min(gsize - 1, ce)
    out += [(y, x) for y in range(min(gsize - 2, re - 1), max(-1, rs - 2), -1)]
    y = max(0, rs - 1)
    out += [(y, x) for x in range(min(gsize - 2, ce - 1), max(0, cs - 1), -1)]
    return out


def check_foreground(zone_fg: np.ndarray, rects: np.ndarray, rs: int, re: int, cs: int, ce: int) -> bool:
    """Region.check_foreground (sf_meta.py:342-381) from the per-zone foreground pixel counts (ckb_zone_fg_counts:
    int32 [g, g]) and the zone rectangles (int [g, g, 4]): True = calm. The zones tile the region's sub-image exactly, so
    its foreground sum is the sum of its zones' counts."""
    g = zone_fg.shape[0]
    S = 20 * g
    moving, border_threshold = 0, 2
    for (r, c) in outer_border(rs, re, cs, ce, g):
        a0, b0, a1, b1 = (int(v) for v in rects[r, c])
        if (a1 - a0) * (b1 - b0) * 0.7 < zone_fg[r, c]:
            if (a0 == 0 or a1 == S - 1) and (b0 == 0 or b1 == S - 1):
                moving = border_threshold      # moved at a corner: agitated right away
            moving += 1
            if border_threshold <= moving:
                return False
    threshold = 3 * ((S / g / 2) ** 2) * math.pi      # 3 times the expected area of a stone
    return not (threshold < int(zone_fg[rs:re, cs:ce].sum()))
