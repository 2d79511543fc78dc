"""Multi-rank host logic on CPU with the gloo backend, world_size 2: picture -> rank assignment
and the reference-picture broadcast every rank's DPB copy depends on."""
import hashlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from openvvc_b200 import dist as ovd, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, nb = 256, 128, 6
    mine = ovd.local_pictures(nb, rank, world)
    dpb = [ovd.SharedPicture(w, h, "cpu") for _ in range(nb)]
    digests = []
    for k in range(nb):
        owner = ovd.owner_of(k, world)
        if owner == rank:                      # "reconstruct" picture k: only its owner has the samples
            dpb[k].set_planes(synth.planes(w, h, 100 + k))
        ovd.publish_reference(dpb[k], owner, world)
        digests.append(hashlib.md5(dpb[k].buf.numpy().tobytes()).hexdigest())
    q.put((rank, mine, digests))
    dist.destroy_process_group()


def test_two_rank_reference_exchange_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, d0), (r1, mine1, d1) = res
    assert mine0 == [0, 2, 4] and mine1 == [1, 3, 5]          # picture k -> rank k mod G
    assert d0 == d1                                           # both DPB copies hold identical pictures
    # and they are the owner's samples, not zeros
    from openvvc_b200 import dist as ovd, synth
    ref = ovd.SharedPicture(256, 128, "cpu")
    ref.set_planes(synth.planes(256, 128, 103))
    assert hashlib.md5(ref.buf.numpy().tobytes()).hexdigest() == d0[3]


def test_plane_layout_matches_picture_create_rule():
    from openvvc_b200 import dist as ovd
    layout, total = ovd.plane_layout(3840, 2160)
    assert [p for (_, p, _) in layout] == [7680, 3840, 3840]
    assert total == 7680 * 2160 + 2 * 3840 * 1080
    layout, total = ovd.plane_layout(416, 240)
    assert [p for (_, p, _) in layout] == [1024, 512, 512] and layout[1][0] == 1024 * 240


# ---- random-access GOP as a dependency graph: pictures round-robin over the ranks, references sent to exactly the ranks that need them
def _fake_recon(poc, a, b):
    """stand-in for the reconstruction of a picture from its two references (any deterministic function of them)"""
    return ((a.astype(np.int64) * 3 + b.astype(np.int64) * 5 + poc * 7919) % 1021).astype(np.int16)


def _gop_serial(nb_gops, shape, seed):
    from openvvc_b200 import dist as ovd
    pics = {0: np.random.default_rng(seed).integers(0, 1021, size=shape).astype(np.int16)}
    for poc, (pa, pb) in ovd.stream_pictures(nb_gops):
        pics[poc] = _fake_recon(poc, pics[pa], pics[pb])
    return pics


def _gop_worker(rank, world, port, q, nb_gops, wave):
    from openvvc_b200 import dist as ovd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = (16, 24)
    sched = ovd.GopSchedule(nb_gops, world)
    dpb = {0: torch.from_numpy(np.random.default_rng(5).integers(0, 1021, size=shape).astype(np.int16))}
    sent = [0]

    def reconstruct(poc, parents):
        assert all(p in dpb for p in parents), "rank %d: picture %d reconstructed before its references arrived" % (rank, poc)
        dpb[poc] = torch.from_numpy(_fake_recon(poc, dpb[parents[0]].numpy(), dpb[parents[1]].numpy()))

    def publish(poc, consumers):
        for c in consumers:
            dist.send(dpb[poc], dst=c)
            sent[0] += 1

    def consume(poc, owner):
        dpb[poc] = torch.empty(shape, dtype=torch.int16)
        dist.recv(dpb[poc], src=owner)

    ovd.run_gop_stream(sched, rank, reconstruct, publish, consume, order=sched.wave_order() if wave else None)
    q.put((rank, {poc: hashlib.md5(t.numpy().tobytes()).hexdigest() for poc, t in dpb.items()}, sent[0]))
    dist.destroy_process_group()


def test_gop_structure():
    from openvvc_b200 import dist as ovd
    order = ovd.gop_decode_order()
    assert order[:10] == [32, 16, 8, 4, 2, 1, 3, 6, 5, 7] and sorted(order) == list(range(1, 33))
    assert ovd.gop_parents(32) == (0, 0) and ovd.gop_parents(16) == (0, 32) and ovd.gop_parents(24) == (16, 32)
    assert ovd.gop_parents(12) == (8, 16) and ovd.gop_parents(13) == (12, 14)
    seen = {0}
    for poc, parents in ovd.stream_pictures(3):                # every reference is decoded before its dependents
        assert all(p in seen for p in parents), (poc, parents)
        seen.add(poc)
    assert len(seen) == 97
    s1 = ovd.GopSchedule(2, 1)
    assert s1.exchange_bytes(100) == 0
    s8 = ovd.GopSchedule(2, 8)
    assert all(s8.owner[poc] not in c for poc, c in s8.consumers.items())
    assert s8.owner[32] == 0 and s8.owner[16] == 1 and 1 in s8.consumers[32]
    wo = ovd.GopSchedule(4, 8).wave_order()
    assert sorted(p for p, _ in wo) == sorted(p for p, _ in ovd.stream_pictures(4))
    seen = {0}
    for poc, parents in wo:                                    # the issue order is a valid order too
        assert all(p in seen for p in parents), (poc, parents)
        seen.add(poc)
    assert [p for p, _ in wo[:6]] == [32, 16, 8, 24, 64, 4]       # the next GOP starts while the first one is two levels deep


import pytest


@pytest.mark.parametrize("wave", [0, 1], ids=["decode_order", "wave_order"])
def test_two_rank_gop_stream_gloo(wave):
    """the GOP walk on 2 ranks (gloo): every rank ends with the pictures it owns or needs, identical to a serial walk, and only
    pictures with consumers on the other rank travel"""
    from openvvc_b200 import dist as ovd
    world, port, nb_gops = 2, _free_port(), 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gop_worker, args=(r, world, port, q, nb_gops, wave)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = {poc: hashlib.md5(a.tobytes()).hexdigest() for poc, a in _gop_serial(nb_gops, (16, 24), 5).items()}
    sched = ovd.GopSchedule(nb_gops, world)
    total_sent = 0
    for rank, have, sent in res:
        total_sent += sent
        for poc, dg in have.items():
            assert dg == want[poc], (rank, poc)
        for poc, _ in sched.pictures:
            assert (poc in have) == (sched.owner[poc] == rank or sched.needs(rank, poc)), (rank, poc)
    assert total_sent == sched.exchange_bytes(1)
