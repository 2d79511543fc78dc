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
