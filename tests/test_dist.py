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
