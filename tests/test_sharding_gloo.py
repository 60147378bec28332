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


def _worker(rank, world, port, n_frames, psy=1):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import cases
    import oracle
    from odr_audioenc_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs, mode, br, pcm, _, _ = cases.make_case("Bj", "S8", n_frames)
    c = oracle.configure(fs, mode, br, psy)
    rng = sharding.time_shards(n_frames, world, psy_model=psy)[rank]
    first, end = sharding.pcm_slice(rng)
    mine = pcm[first:end]  # all this rank is given
    # the encoder sees a stream that starts at `first`; frames before f0 only provide history
    lead = rng.history_samples // 1152
    assert rng.history_samples in (0, 1152 if psy != 2 else 2304)
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


@pytest.mark.parametrize("n_frames,psy", [(31, 1), (40, 1), (36, 2)])
def test_time_shards_concatenate_to_the_one_shot_stream(n_frames, psy):
    """psy model 2 looks 1632 samples back: its shards carry two frames of history (tlb_info.halo_samples rounded up)"""
    tmp.spawn(_worker, args=(2, _free_port(), n_frames, psy), nprocs=2, join=True)


def _ensemble_worker(rank, world, port):
    """three services, two ranks: one service is cut in time; rank 0 reassembles every service from the pieces"""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle
    import signals
    from odr_audioenc_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ens = [(48000, "j", 192, 150), (48000, "j", 192, 150), (48000, "m", 96, 150)]
    services = [(fs, 1 if m == "m" else 2, br, n) for fs, m, br, n in ens]
    plan = sharding.ensemble_shards(services, world, min_piece_frames=8)
    assert any(len({q.service for q in pieces}) < len(pieces) or any(q.f0 > 0 or q.f1 < 150 for q in pieces) for pieces in plan)
    ok = torch.ones(1, dtype=torch.int64)
    for owner in range(world):
        for p in plan[owner]:
            fs, mode, br, n = ens[p.service]
            nch = 1 if mode == "m" else 2
            c = oracle.configure(fs, mode, br)
            pcm = signals.make("S8", n, nch, fs)
            nbytes = (p.f1 - p.f0) * c.lg_frame
            buf = torch.zeros(nbytes, dtype=torch.uint8)
            if rank == owner:
                first, end = sharding.pcm_slice(p)
                lead = p.history_samples // 1152
                part, _ = oracle.encode(c, pcm[first:end], lead, lead + (p.f1 - p.f0))
                buf = torch.from_numpy(part.copy())
            if owner != 0:
                if rank == owner:
                    dist.send(buf, dst=0)
                elif rank == 0:
                    dist.recv(buf, src=owner)
            if rank == 0:
                want, _ = oracle.encode(c, pcm)
                ok[0] &= int(np.array_equal(buf.numpy(), want[p.f0 * c.lg_frame:p.f1 * c.lg_frame]))
    dist.broadcast(ok, src=0)
    dist.destroy_process_group()
    assert ok.item() == 1


def test_ensemble_pieces_reassemble_every_service():
    tmp.spawn(_ensemble_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_time_shards_cover_every_frame_once():
    from odr_audioenc_b200 import sharding
    for n, w in ((1500000, 8), (7, 8), (100, 3), (0, 2)):
        r = sharding.time_shards(n, w)
        assert r[0].f0 == 0 and r[-1].f1 == n and all(a.f1 == b.f0 for a, b in zip(r, r[1:]))
        assert max(x.f1 - x.f0 for x in r) - min(x.f1 - x.f0 for x in r) <= 1
        assert all(x.history_samples == (0 if x.f0 == 0 else 1152) for x in r)
        assert all(x.has_next == (x.f1 < n and x.f1 > x.f0) for x in r)
        r2 = sharding.time_shards(n, w, psy_model=2)
        assert all(x.history_samples == min(x.f0 * 1152, 2304) and x.history_samples in (0, 1152, 2304) for x in r2)
    assert sharding.halo_for(1) == 1152 and sharding.halo_for(2) == 2304 and sharding.halo_for(halo_samples=1632) == 2304
    with pytest.raises(ValueError):
        sharding.time_shards(100, 2, halo=1152, psy_model=2)   # shorter than what psy model 2 looks back on
    with pytest.raises(ValueError):
        sharding.halo_for(halo_samples=100)


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


def test_ensemble_shards_split_services_in_time():
    """18 equal-length services on 8 GPUs: whole services cap the speed-up at 6.0x (3 on one rank); cut in time every
    rank gets the same cost, each frame appears once, pieces carry halo and look-ahead like time shards"""
    from odr_audioenc_b200 import sharding
    sv = [(48000, 2, 192, 150000)] * 6 + [(48000, 2, 160, 150000)] * 4 + [(48000, 2, 128, 150000)] * 4 + \
         [(48000, 2, 112, 150000)] * 2 + [(48000, 1, 96, 150000)] * 2
    for w in (1, 2, 3, 4, 8):
        plan = sharding.ensemble_shards(sv, w)
        assert plan == sharding.ensemble_shards(sv, w)
        cost = sharding.plan_cost(plan, sv)
        assert max(cost) / (sum(cost) / w) < 1.001, (w, cost)
        cover = {}
        for pieces in plan:
            for q in pieces:
                cover.setdefault(q.service, []).append(q)
                assert q.history_samples == min(q.f0 * 1152, 1152) and q.has_next == (q.f1 < 150000)
        assert sorted(cover) == list(range(18))
        for segs in cover.values():
            segs.sort(key=lambda q: q.f0)
            assert segs[0].f0 == 0 and segs[-1].f1 == 150000 and all(a.f1 == b.f0 for a, b in zip(segs, segs[1:]))
        assert sum(len(v) - 1 for v in cover.values()) <= w - 1   # at most world-1 services are cut
    # degenerate inputs: more ranks than work, empty ensemble
    assert sharding.ensemble_shards([], 4) == [[], [], [], []]
    tiny = sharding.ensemble_shards([(48000, 2, 192, 100)], 8)
    assert sum(q.f1 - q.f0 for p in tiny for q in p) == 100
