"""CPU: the multi-GPU host logic with world size 2 over gloo.  Each rank takes its frame range (with halo and
look-ahead frame) or its services from the plan, encodes with the ORACLE standing in for the GPU encoder (this is a
test of the partitioning, not of the kernels), and rank 0 checks that the concatenation equals the one-shot stream."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import cases
    import oracle
    from odr_audioenc_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs, mode, br, pcm, _, _ = cases.make_case("Bj", "S8", n_frames)
    c = oracle.configure(fs, mode, br)
    rng = sharding.time_shards(n_frames, world)[rank]
    first, end = sharding.pcm_slice(rng)
    mine = pcm[first:end]  # all this rank is given
    # the encoder sees a stream that starts at `first`; frames before f0 only provide history
    lead = rng.history_samples // 1152
    assert rng.history_samples in (0, 1152)
    part, _ = oracle.encode(c, mine, lead, lead + (rng.f1 - rng.f0))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([part.size], dtype=torch.int64))
    buf = torch.zeros(int(max(s.item() for s in sizes)), dtype=torch.uint8)
    buf[:part.size] = torch.from_numpy(part)
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    ok = torch.ones(1, dtype=torch.int64)
    if rank == 0:
        whole = np.concatenate([g.numpy()[:int(s.item())] for g, s in zip(gathered, sizes)])
        want, _ = oracle.encode(c, pcm)
        ok[0] = int(np.array_equal(whole, want))
    dist.broadcast(ok, src=0)
    dist.destroy_process_group()
    assert ok.item() == 1


@pytest.mark.parametrize("n_frames", [31, 40])
def test_time_shards_concatenate_to_the_one_shot_stream(n_frames):
    tmp.spawn(_worker, args=(2, _free_port(), n_frames), nprocs=2, join=True)


def test_time_shards_cover_every_frame_once():
    from odr_audioenc_b200 import sharding
    for n, w in ((1500000, 8), (7, 8), (100, 3), (0, 2)):
        r = sharding.time_shards(n, w)
        assert r[0].f0 == 0 and r[-1].f1 == n and all(a.f1 == b.f0 for a, b in zip(r, r[1:]))
        assert max(x.f1 - x.f0 for x in r) - min(x.f1 - x.f0 for x in r) <= 1
        assert all(x.history_samples == (0 if x.f0 == 0 else 1152) for x in r)
        assert all(x.has_next == (x.f1 < n and x.f1 > x.f0) for x in r)


def test_service_shards_balance_an_ensemble():
    from odr_audioenc_b200 import sharding
    # BASELINE config D: 18 services x 1 h, mixed 96-192 kbit/s
    sv = [(48000, 2, 192, 150000)] * 6 + [(48000, 2, 160, 150000)] * 4 + [(48000, 2, 128, 150000)] * 4 + \
         [(48000, 2, 112, 150000)] * 2 + [(48000, 1, 96, 150000)] * 2
    plan = sharding.service_shards(sv, 8)
    assert sorted(i for p in plan for i in p) == list(range(18))
    load = [sum(sharding.service_cost(*sv[i]) for i in p) for p in plan]
    assert max(load) / (sum(load) / 8) < 1.25
    assert plan == sharding.service_shards(sv, 8)  # deterministic: every rank derives the same plan
