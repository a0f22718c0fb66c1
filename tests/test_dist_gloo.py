"""World-size-2/3 tests of the image-sharding + all-gather host logic on CPU (gloo, 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maskrcnn_b200 import dist as mdist


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mdist.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_detections(n_images, d):
    rng = np.random.default_rng(123)
    counts = rng.integers(0, d + 1, n_images).astype(np.int32)
    dets = np.zeros((n_images, d, 6), np.float32)
    for i, c in enumerate(counts):
        dets[i, :c] = rng.standard_normal((c, 6)).astype(np.float32) + i
    return dets, counts


def _worker(rank, world, port, n_images, d, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dets, counts = _fake_detections(n_images, d)
        b, e = mdist.shard_range(n_images, rank, world)
        got_d, got_c = mdist.gather_detections(torch.from_numpy(dets[b:e]), torch.from_numpy(counts[b:e]), n_images=n_images)
        ok = (got_d.shape == (n_images, d, 6) and np.array_equal(got_d.numpy(), dets) and np.array_equal(got_c.numpy(), counts))
        # without the global image count (ranks agree on the padding through a max-reduce)
        got_d2, got_c2 = mdist.gather_detections(torch.from_numpy(dets[b:e]), torch.from_numpy(counts[b:e]))
        ok = ok and np.array_equal(got_d2.numpy(), dets) and np.array_equal(got_c2.numpy(), counts)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_images", [(2, 8), (2, 7), (3, 8)])
def test_gather_detections_matches_single_process(tmp_path, world, n_images):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, 5, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "ok%d" % r)).read() == "1"


def test_gather_is_identity_without_process_group():
    d, c = torch.zeros(2, 3, 6), torch.zeros(2, dtype=torch.int32)
    a, b = mdist.gather_detections(d, c)
    assert a is d and b is c


def test_owner_rank_is_the_inverse_of_shard_range():
    """detection_collect_kernel finds the rank that owns an image with this formula; it must invert shard_range()."""
    from maskrcnn_b200 import dist as mdist
    for n in (1, 7, 8, 63, 64, 65):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                b, e = mdist.shard_range(n, r, world)
                for i in range(b, e):
                    assert mdist.owner_rank(i, n, world) == r
