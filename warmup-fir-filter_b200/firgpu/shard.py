"""Partition planner for running the FIR path on the 1/2/4/8 GPUs of one box.

The path shards without any data-path collective: rows (channels) are
independent, and a long signal or a tall image is cut into contiguous pieces
that carry a duplicated halo of ``taps - 1`` samples (rows) copied at shard
time.  One process drives one GPU (``torch.distributed`` is used only for the
barrier / timing reduction in the bench and for optional result gathering).

Halo sizes follow the reference's centre alignment
(fir_1d/model/python/fir_1d_fixed_ref.py:86,99-104): output n reads inputs
``n - (k - L//2) * dilation`` for ``k = 0..L-1``, i.e. ``(L - 1 - L//2) *
dilation`` samples before and ``(L//2) * dilation`` after.
"""
from __future__ import annotations

from dataclasses import dataclass


def taps_reach(n_taps: int, dilation: int = 1) -> tuple[int, int]:
    """(samples read before the output position, samples read after it)."""
    centre = n_taps // 2
    return (n_taps - 1 - centre) * dilation, centre * dilation


def split_even(total: int, parts: int, align: int = 1) -> list[tuple[int, int]]:
    """``parts`` contiguous [start, stop) ranges covering [0, total); every cut
    is a multiple of ``align`` (the last range absorbs the remainder).  Ranges
    may be empty when ``total`` is small."""
    if parts < 1 or align < 1 or total < 0:
        raise ValueError("split_even needs parts >= 1, align >= 1, total >= 0")
    units = -(-total // align)
    base, extra = divmod(units, parts)
    out, cursor = [], 0
    for p in range(parts):
        n_units = base + (1 if p < extra else 0)
        start = min(cursor * align, total)
        cursor += n_units
        stop = min(cursor * align, total)
        out.append((start, stop))
    return out


@dataclass(frozen=True)
class Piece:
    """One rank's share of an axis: ``[start, stop)`` plus how many real
    neighbour elements travel with it on each side."""
    start: int
    stop: int
    before: int
    after: int

    @property
    def size(self) -> int:
        return self.stop - self.start

    @property
    def lo(self) -> int:          # first element copied to the rank
        return self.start - self.before

    @property
    def hi(self) -> int:          # one past the last element copied
        return self.stop + self.after


def plan_segments(n: int, parts: int, n_taps: int, dilation: int = 1, align: int = 4096) -> list[Piece]:
    """Contiguous signal segments (configs 2 and 5).  ``align`` keeps every
    cut on a tile boundary and, for interleaved I/Q, on an even scalar."""
    if dilation == 2 and align % 2:
        raise ValueError("I/Q segments must be cut on an even scalar index")
    before, after = taps_reach(n_taps, dilation)
    return [Piece(a, b, min(before, a), min(after, n - b)) if b > a else Piece(a, b, 0, 0)
            for a, b in split_even(n, parts, align)]


def plan_rows(rows: int, parts: int) -> list[Piece]:
    """Channel / row sharding (config 3): no halo, rows are zero-padded
    separately (fir_1d/sim/vector/gen_fixed_output.py:44-59)."""
    return [Piece(a, b, 0, 0) for a, b in split_even(rows, parts)]


def plan_bands(height: int, parts: int, kh: int, align: int = 1) -> list[Piece]:
    """Row bands of an image (config 4) with ``kh - 1`` duplicated halo rows."""
    above, below = taps_reach(kh, 1)
    return [Piece(a, b, min(above, a), min(below, height - b)) if b > a else Piece(a, b, 0, 0)
            for a, b in split_even(height, parts, align)]


def aligned_lead(before: int, itemsize: int) -> int:
    """Offset (in elements) at which a segment body is placed inside its local
    buffer, with the halo right before it.  The bulk-copy kernel needs the body
    16-byte aligned; it is placed on a 256-byte boundary, because 32 KiB bulk
    copies that start 16 bytes into a 128-byte line straddle one more line each
    (measured: 0.170 instead of 0.1655 ms per 2^28-sample step)."""
    q = max(256 // itemsize, 1)
    return -(-before // q) * q


def scatter_segment(full, piece: Piece, device, pinned: bool = False):
    """Copy ``full[piece.lo:piece.hi]`` (a host NumPy array or CPU tensor, 1-D)
    into a fresh device buffer and return ``(buffer, body_view)``; pass
    ``body_view`` with ``halo_left_valid=piece.before``,
    ``halo_right_valid=piece.after`` to ``ops.fir1d_fixed``."""
    import torch
    src = torch.as_tensor(full)
    lead = aligned_lead(piece.before, src.element_size())
    buf = torch.empty(lead + piece.size + piece.after, dtype=src.dtype, device=device)
    chunk = src[piece.lo:piece.hi]
    if pinned:
        chunk = chunk.pin_memory()
    buf[lead - piece.before:lead + piece.size + piece.after].copy_(chunk, non_blocking=pinned)
    return buf, buf[lead:lead + piece.size]


def scatter_band(full, piece: Piece, device):
    """2-D analogue: ``full`` is [..., H, W]; returns ``(buffer, body_view)``
    where ``body_view`` is the band without its halo rows."""
    import torch
    src = torch.as_tensor(full)
    buf = src[..., piece.lo:piece.hi, :].to(device).contiguous()
    return buf, buf[..., piece.before:piece.before + piece.size, :]


def exchange_halos(buf, lead: int, pieces: list[Piece], rank: int, dist) -> None:
    """Shard-time halo fill for data that is already distributed: ``buf`` is
    this rank's ``[lead pad | halo before | body | halo after]`` buffer with
    the body in place; the halo slots receive the neighbours' edge samples by
    point-to-point send/recv (NCCL on GPUs, gloo in the CPU tests).  This is
    the only inter-rank traffic the path ever needs, and it happens once,
    outside the filtering loop."""
    import torch
    me = pieces[rank]
    body = buf[lead:lead + me.size]

    def raw(t):                       # NCCL has no int16: move the bytes
        return t.contiguous().view(torch.uint8)

    ops, keep = [], []
    if me.after:
        ops.append(dist.P2POp(dist.irecv, raw(buf[lead + me.size:lead + me.size + me.after]), rank + 1))
    if me.before:
        ops.append(dist.P2POp(dist.irecv, raw(buf[lead - me.before:lead]), rank - 1))
    if rank + 1 < len(pieces) and pieces[rank + 1].before:
        keep.append(raw(body[me.size - pieces[rank + 1].before:].clone()))
        ops.append(dist.P2POp(dist.isend, keep[-1], rank + 1))
    if rank > 0 and pieces[rank - 1].after:
        keep.append(raw(body[:pieces[rank - 1].after].clone()))
        ops.append(dist.P2POp(dist.isend, keep[-1], rank - 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_band_halos(buf, pieces: list[Piece], rank: int, dist) -> None:
    """Row-band analogue of :func:`exchange_halos` for images that are already
    distributed: ``buf`` is ``[..., before + size + after, W]`` with this
    rank's band rows in place; the halo rows receive the neighbours' edge rows
    (send/recv of contiguous copies, once, at shard time)."""
    import torch
    me = pieces[rank]
    ops, recvs, keep = [], [], []

    def raw(t):
        return t.contiguous().view(torch.uint8)

    if me.before:
        tmp = torch.empty_like(buf[..., :me.before, :]).contiguous()
        recvs.append((tmp, slice(0, me.before)))
        ops.append(dist.P2POp(dist.irecv, raw(tmp), rank - 1))
    if me.after:
        tmp = torch.empty_like(buf[..., me.before + me.size:, :]).contiguous()
        recvs.append((tmp, slice(me.before + me.size, me.before + me.size + me.after)))
        ops.append(dist.P2POp(dist.irecv, raw(tmp), rank + 1))
    body = buf[..., me.before:me.before + me.size, :]
    if rank + 1 < len(pieces) and pieces[rank + 1].before:
        keep.append(raw(body[..., me.size - pieces[rank + 1].before:, :]))
        ops.append(dist.P2POp(dist.isend, keep[-1], rank + 1))
    if rank > 0 and pieces[rank - 1].after:
        keep.append(raw(body[..., :pieces[rank - 1].after, :]))
        ops.append(dist.P2POp(dist.isend, keep[-1], rank - 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for tmp, rows in recvs:
        buf[..., rows, :] = tmp
