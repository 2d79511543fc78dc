"""Frame-parallel sharding of the reconstruction path over the GPUs of one node.

The path shards by PICTURE, like the reference's frame threads (ovdec.c:189-248): picture k (decode order) is
reconstructed by rank k mod G.  The only data-path exchange is the reference-picture hand-off: when a picture that
later pictures reference is finished (post-ALF), its owner broadcasts the three planes to the other ranks' DPB copies
(NCCL over NVLink on GPUs; gloo in the CPU tests).  torch.distributed is plumbing only.
"""
import numpy as np
import torch
import torch.distributed as dist


def owner_of(picture_idx, world):
    """rank that reconstructs picture `picture_idx` (decode order)"""
    return picture_idx % world


def local_pictures(nb_pictures, rank, world):
    """decode-order indices of the pictures this rank reconstructs"""
    return [k for k in range(nb_pictures) if owner_of(k, world) == rank]


def plane_layout(w, h, align=256):
    """(offset, pitch, rows) in bytes of the three 16-bit planes inside one flat picture buffer; the same layout
    ovgpu_picture_create uses (pitch = width * 2 rounded up to 256 bytes)."""
    out, off = [], 0
    for (pw, ph) in ((w, h), (w // 2, h // 2), (w // 2, h // 2)):
        pitch = (pw * 2 + align - 1) // align * align
        out.append((off, pitch, ph))
        off += pitch * ph
    return out, off


class SharedPicture:
    """One DPB picture as a flat torch uint8 tensor (so torch.distributed can move it) + its plane pointers/pitches
    for ovgpu_picture_wrap."""

    def __init__(self, w, h, device):
        self.w, self.h = w, h
        self.layout, nbytes = plane_layout(w, h)
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)

    def plane_ptrs(self):
        base = self.buf.data_ptr()
        return [base + off for (off, _, _) in self.layout], [pitch for (_, pitch, _) in self.layout]

    def plane_view(self, c):
        off, pitch, rows = self.layout[c]
        pw = self.w if c == 0 else self.w // 2
        return self.buf[off:off + pitch * rows].view(rows, pitch).view(torch.int16)[:, :pw]

    def set_planes(self, planes):
        for c, p in enumerate(planes):
            self.plane_view(c).copy_(torch.from_numpy(np.ascontiguousarray(p).view(np.int16)).to(self.buf.device))


def publish_reference(pic, src_rank, world):
    """Owner -> everybody hand-off of a finished reference picture (ncclBroadcast of P bytes; no-op for one rank)."""
    if world > 1:
        dist.broadcast(pic.buf, src=src_rank)


# ---- random-access GOP as a dependency graph ("33-picture GOP") ---------------------------------------
# The common-test-conditions random-access configuration codes a GOP of 32 pictures as a binary hierarchy: POC 32 first (from the
# previous GOP's POC 32), then 16 from (0, 32), 8 from (0, 16), ... down to the odd pictures from their two neighbours.  That
# shape -- a chain of 6, then 1-1-2-4-8-16 pictures wide -- is what frame-parallel decoding has to live with.

def gop_decode_order(gop=32):
    """POC offsets 1..gop of one GOP in decoding order: gop, then the hierarchy depth first (32, 16, 8, 4, 2, 1, 3, 6, 5, 7, ...)"""
    out = [gop]

    def rec(lo, hi):
        if hi - lo < 2:
            return
        mid = (lo + hi) // 2
        out.append(mid)
        rec(lo, mid)
        rec(mid, hi)
    rec(0, gop)
    return out


def gop_parents(x, gop=32):
    """POC offsets (relative to the GOP's first POC, 0 = the previous GOP's last picture) of the two reference pictures of the
    picture at offset x in 1..gop"""
    if x == gop:
        return (0, 0)
    s = x & -x                      # largest power of two dividing x
    return (x - s, x + s)


def stream_pictures(nb_gops, gop=32):
    """The pictures of nb_gops GOPs in decoding order (after POC 0, which is given): list of (poc, (parent poc, parent poc))"""
    out = []
    for g in range(nb_gops):
        base = g * gop
        for x in gop_decode_order(gop):
            a, b = gop_parents(x, gop)
            out.append((base + x, (base + a, base + b)))
    return out


class GopSchedule:
    """Who reconstructs what and who needs it: picture k of the decoding order -> rank k mod world; the consumers of a picture are
    the other ranks that own a picture referencing it.  POC 0 is resident on every rank."""

    def __init__(self, nb_gops, world, gop=32):
        self.world = world
        self.pictures = stream_pictures(nb_gops, gop)
        self.owner = {0: None}
        for k, (poc, _) in enumerate(self.pictures):
            self.owner[poc] = owner_of(k, world)
        cons = {poc: set() for poc in self.owner}
        for (poc, parents) in self.pictures:
            for p in set(parents):
                if p != 0 and self.owner[p] != self.owner[poc]:
                    cons[p].add(self.owner[poc])
        self.consumers = {poc: sorted(s) for poc, s in cons.items()}
        last = {}
        for k, (poc, parents) in enumerate(self.pictures):
            for p in parents:
                last[p] = k
        self.last_use = last            # decode index of the last picture that references a picture
        self.gop = gop

    def wave_order(self, stride=2):
        """ISSUE order for frame-parallel reconstruction: pictures sorted by (GOP * stride + level in the hierarchy).  Decoding
        order walks one GOP's dependency chain (32, 16, 8, 4, 2, 1: six pictures that can only run one after the other) before it
        reaches anything else; issued in that order, in-order queues (a stream, a communicator) idle behind the chain.  In wave
        order a new GOP starts every `stride` levels, so three GOPs' chains are in flight together.  It is a valid order: a
        picture's references are on lower levels of its GOP, or the level-0 picture of the previous GOP.  Ownership stays by
        DECODING-order index."""
        def level(x):
            return 0 if x == self.gop else (self.gop.bit_length() - 1) - ((x & -x).bit_length() - 1)
        keyed = []
        for k, (poc, parents) in enumerate(self.pictures):
            g, x = (poc - 1) // self.gop, (poc - 1) % self.gop + 1
            keyed.append((g * stride + level(x), k, poc, parents))
        keyed.sort()
        return [(poc, parents) for (_, _, poc, parents) in keyed]

    def needs(self, rank, poc):
        return rank in self.consumers[poc]

    def exchange_bytes(self, picture_bytes):
        """bytes that cross NVLink for the whole stream (one copy per consumer rank)"""
        return sum(len(c) for c in self.consumers.values()) * picture_bytes


def run_gop_stream(sched, rank, reconstruct, publish, consume, first=0, last=None, order=None):
    """Walks the stream in decoding order on one rank -- the SAME order on every rank, which is what keeps the grouped send /
    receive calls matched.  reconstruct(poc, parents) for pictures this rank owns, then publish(poc, consumers) when other ranks
    need it; consume(poc, owner) for pictures it needs from others.  first / last: a slice of the order; order: an issue order other than the decoding order (GopSchedule.wave_order), the same on every rank."""
    for (poc, parents) in (order if order is not None else sched.pictures)[first:last]:
        o = sched.owner[poc]
        if o == rank:
            reconstruct(poc, parents)
            if sched.consumers[poc]:
                publish(poc, sched.consumers[poc])
        elif sched.needs(rank, poc):
            consume(poc, o)
