"""N > 1 host logic on CPU: frame sharding + gather of detections over a world-size-2 gloo group.
The per-frame detector here is the CPU oracle (checker side only; the GPU kernels are covered by
the `-m gpu` tests) -- what is under test is the partition, the no-collective data path and the
frame-order gather of aruco_slam_b200/shard.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_shard_bounds_cover_everything():
    from aruco_slam_b200 import shard
    for n in (0, 1, 5, 32, 33, 64):
        for w in (1, 2, 3, 4, 8):
            b = shard.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            sizes = [b[i + 1] - b[i] for i in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_slice(4, 2, 2)


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from aruco_slam_b200 import shard, synth, dictionaries as D
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    dic = D.getPredefinedDictionary(D.DICT_4X4_50)
    frames = [synth.render_config("C1", s).image for s in range(n_frames)]
    seen = []

    def detect(block):
        seen.append(len(block))
        return [O.detect(f, dic) for f in block]

    res = shard.detect_sharded(detect, frames, rank, world)
    if rank == 0:
        q.put(("result", [(c.tolist(), i.tolist(), r.tolist()) for c, i, r in res]))
    q.put(("seen", rank, seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 1])
def test_two_rank_gloo_gather_matches_single_process(n_frames):
    import torch.multiprocessing as mp
    from aruco_slam_b200 import synth, dictionaries as D
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=180) for _ in range(3)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    result = next(m[1] for m in msgs if m[0] == "result")
    seen = {m[1]: m[2] for m in msgs if m[0] == "seen"}
    assert sum(sum(v) for v in seen.values()) == n_frames           # every frame processed exactly once
    dic = D.getPredefinedDictionary(D.DICT_4X4_50)
    assert len(result) == n_frames
    for s, (c, i, r) in enumerate(result):
        oc, oi, orj = O.detect(synth.render_config("C1", s).image, dic)
        assert np.array_equal(np.array(i).reshape(-1), oi.reshape(-1))
        assert np.array_equal(np.array(c, np.float32).reshape(-1), oc.reshape(-1))
        assert np.array_equal(np.array(r, np.float32).reshape(-1), orj.reshape(-1))
