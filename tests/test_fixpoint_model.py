"""CPU model of the iteration behind nms_fixpoint_pub_kernel / proposal_fixpoint_kernel (maskrcnn_b200/csrc/nms.cu, proposal.cu):
greedy NMS as the fixed point of "box i survives iff no SURVIVING earlier box suppresses it", iterated over all chunks of 64 boxes at
once (exact inside a chunk, previous pass across chunks).  Checks the two claims the kernels rest on - the fixed point is the greedy
answer, and chunk c is final after pass c + 1 (so W + 1 passes bound the loop) - on clustered boxes, on a chain in which every box
only overlaps its successor (the worst case), and with the asynchronous reads the kernels allow (a chunk may see a newer word)."""
import numpy as np

import oracle
from maskrcnn_b200 import synth


def _suppression(d, thr):
    y1, x1, y2, x2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    area = (y2 - y1 + 1) * (x2 - x1 + 1)
    h = np.maximum(0, np.minimum(y2[:, None], y2[None]) - np.maximum(y1[:, None], y1[None]) + 1)
    w = np.maximum(0, np.minimum(x2[:, None], x2[None]) - np.maximum(x1[:, None], x1[None]) + 1)
    inter = (w * h).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / ((area[:, None] + area[None]) - inter) >= np.float32(thr)


def _block_fixed_point(M, chunk=64, newer=None):
    """Returns (keep, passes, passes after which each chunk last changed).  M[j, i]: box j suppresses box i (score order).
    newer: optional rng - a chunk then reads, per earlier chunk, either the previous pass's word or this pass's (asynchrony)."""
    n = len(M)
    L = np.tril(M.T, -1)                       # L[i, j]: an earlier box j suppresses box i
    W = (n + chunk - 1) // chunk
    keep = np.ones(n, bool)
    last_change = np.zeros(W, int)
    for p in range(1, W + 3):
        new = keep.copy()
        for c in range(W):
            lo, hi = c * chunk, min(n, (c + 1) * chunk)
            seen = keep.copy()
            if newer is not None:               # words of earlier chunks already published in this pass
                for e in range(c):
                    if newer.random() < 0.5:
                        seen[e * chunk:(e + 1) * chunk] = new[e * chunk:(e + 1) * chunk]
            cand = ~(L[lo:hi, :lo].astype(np.int32) @ seen[:lo].astype(np.int32) > 0)
            alive = cand.copy()
            D = L[lo:hi, lo:hi]
            while True:                          # the chunk's own triangular system, Jacobi to its fixed point
                nxt = cand & ~((D.astype(np.int32) @ alive.astype(np.int32)) > 0)
                if (nxt == alive).all():
                    break
                alive = nxt
            if (alive != keep[lo:hi]).any():
                last_change[c] = p
            new[lo:hi] = alive
        if (new == keep).all():
            return keep, p, last_change
        keep = new
    raise AssertionError("no fixed point within W + 2 passes")


def _dets(n, seed, thr_cluster=True):
    rng = np.random.default_rng(seed)
    b = synth.random_rois(n, seed, image=512.0, min_size=12, max_size=200) * 512.0
    if thr_cluster:
        b[n // 2:] = b[: n - n // 2] + rng.uniform(-5, 5, (n - n // 2, 4)).astype(np.float32)
    s = np.sort(synth.unique_scores(n, seed))[::-1]
    return np.concatenate([b, s[:, None]], 1).astype(np.float32)


def test_fixed_point_is_the_greedy_answer_and_settles_fast():
    for n, thr, seed in ((700, 0.7, 1), (700, 0.3, 2), (450, 0.5, 3)):
        d = _dets(n, seed)
        keep, passes, last = _block_fixed_point(_suppression(d, thr))
        np.testing.assert_array_equal(np.nonzero(keep)[0], oracle.nms(d, thr))     # scores descending: score order = index order
        assert passes <= 12
        assert all(last[c] <= c + 1 for c in range(len(last)))


def test_chain_needs_one_pass_per_chunk_and_no_more():
    n = 640
    x = np.arange(n, dtype=np.float32) * 20
    d = np.stack([np.zeros(n, np.float32), x, np.full(n, 100, np.float32), x + 100, np.linspace(0.99, 0.01, n, dtype=np.float32)], 1)
    keep, passes, last = _block_fixed_point(_suppression(d, 0.5))
    np.testing.assert_array_equal(np.nonzero(keep)[0], oracle.nms(d, 0.5))
    W = n // 64
    assert passes <= W + 1 and all(last[c] <= c + 1 for c in range(W))


def test_asynchronous_reads_reach_the_same_fixed_point():
    d = _dets(600, 7)
    M = _suppression(d, 0.6)
    want = oracle.nms(d, 0.6)
    for seed in range(4):
        keep, _, _ = _block_fixed_point(M, newer=np.random.default_rng(seed))
        np.testing.assert_array_equal(np.nonzero(keep)[0], want)
