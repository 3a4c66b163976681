"""Keyframe-sharded loop check across GPUs: one process per GPU, torch.distributed for the exchange.

Partitioning: keyframe k lives on rank k % world (append-time round robin), never
replicated.  A loop check broadcasts nothing but the 64 KB query (every rank already holds it), runs
the local scan, then does the path's one real exchange: every rank's [per-keyframe records | per-query
packed bests] go round in ONE all_gather (NCCL over NVLink/NVSwitch on GPUs), and one kernel
(lc_merge_shards_device) re-interleaves the records into global keyframe order and takes the unsigned
minimum of the packed bests (dist << 32 | global_row) over the ranks - an integer min, so the merge is
bit-exact and order-independent.
global_row counts rows of all keyframes in insertion order, so the merged best reproduces OpenCV's
multi-image tie-break (lowest image, then lowest row) whatever the world size.

With CUDA IPC available (one node, NVLink / NVSwitch) the exchange needs no collective at all: every rank writes its
contribution into a slot of its own exchange buffer, which the peers have mapped, and ONE kernel per rank
(lc_merge_shards_peer_device) signals the peers, waits for their signals and merges while it reads their slots directly
over peer memory.  The NCCL all_gather + merge kernel stays as the fallback (and for shard doubles on the CPU).

torch is plumbing here (device buffers, the current CUDA stream, the process group).  The compute is
liblc_b200's lc_loop_check_batch_device; a `shard` object with the same three methods can be swapped in
(the CPU tests of the merge logic do that with the oracle over gloo).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _native

_SENTINEL = (-1, 0, -1, -1)  # padding record for ranks with one keyframe fewer
_SIGN = -(1 << 63)


class NativeShard:
    """This rank's shard: an lc_ctx bound to the rank's GPU, outputs in torch CUDA tensors."""

    def __init__(self, rank, world, device, **cfg):
        self.device = torch.device("cuda", device)
        self.ctx = _native.Context(device=device, shard_rank=rank, shard_world=world, **cfg)
        self._sentinel = torch.tensor(_SENTINEL, dtype=torch.int32, device=self.device)
        self._pin = None  # pinned staging for host queries, grown on demand

    def append(self, frame_id, desc, n):
        if desc is None:
            self.ctx.db_append_meta(frame_id, n)
        else:
            self.ctx.db_append(frame_id, desc)

    def local_keyframes(self):
        return self.ctx.db_size()["local_keyframes"]

    def to_device(self, q_rows: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(q_rows, np.uint8).reshape(-1, 32))
        if not t.numel():
            return t.to(self.device)
        if self._pin is None or self._pin.shape[0] < t.shape[0]:
            self._pin = torch.empty((max(t.shape[0], 4096), 32), dtype=torch.uint8).pin_memory()
        else:
            torch.cuda.current_stream(self.device).synchronize()  # the previous upload has left the staging buffer
        self._pin[:t.shape[0]].copy_(t)
        return self._pin[:t.shape[0]].to(self.device, non_blocking=True)

    def new_buffer(self, nbytes):
        return torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)

    def loop_check(self, d_q: torch.Tensor, nq, cur_ids, slots, out=None):
        """d_q: (sum nq, 32) uint8 CUDA tensor.  Returns (records int32 [F, slots, 4], best int64 [sum nq]); `out` =
        (records, best) tensors to fill instead of new ones (the two halves of a contribution buffer)."""
        F = len(nq)
        rec = out[0] if out is not None else torch.empty((F, slots, 4), dtype=torch.int32, device=self.device)
        nloc = self.local_keyframes()
        if nloc != slots:
            rec[:, nloc:] = self._sentinel  # this rank holds one keyframe fewer: pad
        best = out[1] if out is not None else torch.empty(int(np.sum(nq)), dtype=torch.int64, device=self.device)
        # one frame: the library's [nloc] records are a prefix of rec[0]; several: it writes [F, nloc] contiguously
        local = torch.empty((F, nloc, 4), dtype=torch.int32, device=self.device) if (nloc != slots and F > 1) else rec
        # NULL means "the context's own stream" in the C ABI; torch's default stream is handle 0, so name it
        # explicitly (cudaStreamLegacy == 0x1) to stay ordered with the torch ops around this call
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        self.ctx.loop_check_batch_device(d_q.data_ptr() if d_q.numel() else None, nq, cur_ids,
                                         local.data_ptr() if nloc else None, best.data_ptr() if best.numel() else None,
                                         stream)
        if local is not rec and nloc:
            rec[:, :nloc] = local
        return rec, best

    # ---- peer-memory exchange (lc_exchange_*): this rank's buffer, mapped by the peers
    def exchange_setup(self, slot_bytes, group):
        """Collective.  Returns True when every rank has mapped every peer's exchange buffer."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ok, handle = 1, b"\0" * 64
        try:
            handle, base = self.ctx.exchange_create(slot_bytes)
        except _native.LcError:
            ok = 0
        handles = [None] * world
        dist.all_gather_object(handles, (ok, handle), group=group)
        if all(o for o, _ in handles):
            try:
                self.ctx.exchange_connect(world, rank, handles=[h for _, h in handles])
            except _native.LcError:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            try:
                self.ctx.exchange_close()
            except _native.LcError:
                pass
            return False
        self._slots = [_wrap_device_memory(self.ctx.exchange_slot(seq)[0], slot_bytes, self.device) for seq in (2, 1)]  # index = seq & 1
        return True

    def exchange_slot_tensor(self, seq):
        return self._slots[seq & 1]

    def merge_peers(self, seq, F, slots, world, rows):
        """One kernel: signal the peers, wait for theirs, merge straight out of their slots."""
        rec = torch.empty((F, slots * world, 4), dtype=torch.int32, device=self.device)
        best = torch.empty(rows, dtype=torch.int64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        self.ctx.merge_shards_peer_device(seq, F, slots, rows, rec.data_ptr() if rec.numel() else None, best.data_ptr() if rows else None, 3, stream)
        return rec, best

    def merge(self, gathered: torch.Tensor, contrib_bytes, world, F, slots, rows):
        """gathered: the ranks' contribution buffers back to back (uint8).  One kernel: records in global keyframe order
        [F, slots * world, 4] and the unsigned min of the packed bests [rows]."""
        rec = torch.empty((F, slots * world, 4), dtype=torch.int32, device=self.device)
        best = torch.empty(rows, dtype=torch.int64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        self.ctx.merge_shards_device(gathered.data_ptr(), contrib_bytes, world, F, slots, rows, rec.data_ptr() if rec.numel() else None,
                                     best.data_ptr() if rows else None, stream)
        return rec, best


class _DeviceMemory:
    """Library-owned device memory as torch sees it (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _wrap_device_memory(ptr, nbytes, device):
    return torch.as_tensor(_DeviceMemory(ptr, nbytes), device=device)


def merge_torch(gathered: torch.Tensor, contrib_bytes, world, F, slots, rows):
    """The merge of NativeShard.merge in torch ops, for shard doubles without the library (CPU tests over gloo)."""
    g = gathered.view(world, contrib_bytes)
    rec_bytes = F * slots * 16
    rec = g[:, :rec_bytes].contiguous().view(torch.int32).view(world, F, slots, 4)
    # [world, F, slots, 4] -> [F, slots, world, 4]: global keyframe k = slot * world + rank
    merged = rec.permute(1, 2, 0, 3).reshape(F, slots * world, 4)
    # the packed values are uint64 (UINT64_MAX = "nothing compared on this shard") carried in int64 tensors:
    # flipping the sign bit makes the signed min the unsigned one
    best = g[:, rec_bytes:rec_bytes + rows * 8].contiguous().view(torch.int64).view(world, rows)
    best = ((best ^ _SIGN).min(dim=0).values ^ _SIGN) if rows else best.new_empty(0)
    return merged, best


class ShardedLoopDB:
    """Every rank calls every method with the same arguments (SPMD)."""

    def __init__(self, group=None, device=None, shard=None, exchange=None, **cfg):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if shard is None:
            device = device if device is not None else torch.cuda.current_device()
            shard = NativeShard(self.rank, self.world, device, **cfg)
        self.shard = shard
        self.n_keyframes = 0
        # how the shards' results meet: "p2p" = one kernel over peer memory (CUDA IPC), "nccl" = all_gather + merge kernel,
        # "auto" = p2p when the group runs on NCCL and every rank can map every peer, else nccl
        self.exchange = exchange or os.environ.get("LC_SHARD_EXCHANGE", "auto")
        if self.exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(f"exchange = {self.exchange!r}: expected auto, p2p or nccl")
        self._xchg_bytes = 0   # slot size of the live peer exchange (0: none)
        self._xchg_failed = False
        self._seq = 0

    def owns(self, k):
        return k % self.world == self.rank

    def _peer_exchange(self, contrib_bytes):
        """Collective decision (every rank sees the same sizes): is the peer-memory exchange usable for this contribution?"""
        if self.world == 1 or self.exchange == "nccl" or self._xchg_failed or not hasattr(self.shard, "exchange_setup"):
            return False
        if self.exchange == "auto" and dist.get_backend(self.group) != "nccl":
            return False
        if contrib_bytes == 0:
            return False  # nothing to exchange (no query frames): the fallback path handles the empty case
        if contrib_bytes <= self._xchg_bytes:
            return True
        # (re)create with room to grow; peers may still be reading the old buffers: drain every rank first
        torch.cuda.synchronize(self.shard.device)
        dist.barrier(group=self.group)
        want = max(2 * contrib_bytes, int(os.environ.get("LC_SHARD_SLOT_MIN", 1 << 22)))  # (the knob lets tests force regrowth)
        want = (want + 255) & ~255
        if self.shard.exchange_setup(want, self.group):
            self._xchg_bytes, self._seq = want, 0
            return True
        if self.exchange == "p2p":
            raise RuntimeError("exchange='p2p' was asked for but CUDA IPC between the ranks' GPUs is not available")
        self._xchg_failed, self._xchg_bytes = True, 0
        return False

    def append(self, frame_id, desc, n=None):
        """desc may be None on ranks that do not own this keyframe (then n is required)."""
        k = self.n_keyframes
        own = self.owns(k)
        if own and desc is None:
            raise ValueError(f"rank {self.rank} owns keyframe {k} but got no descriptors")
        self.shard.append(frame_id, desc if own else None, len(desc) if desc is not None else int(n))
        self.n_keyframes += 1
        return k

    def loop_check(self, queries, current_frame_ids, copy=True):
        """queries: list of (nq,32) uint8 arrays (or one CUDA uint8 tensor of concatenated rows with
        `nq` given via a tuple (tensor, nq_list)).  Returns (records [F][K] KF_RESULT in global keyframe
        order, best list of uint64 arrays).  copy=False: the arrays are views of the pinned download buffers (two sets,
        used in turn), valid until the next-but-one call - no second pass over 16 bytes per (frame, keyframe)."""
        if isinstance(queries, tuple):
            d_q, nq = queries
            nq = np.asarray(nq, np.int32)
        else:
            nq = np.array([len(q) for q in queries], np.int32)
            rows = np.concatenate([np.asarray(q, np.uint8).reshape(-1, 32) for q in queries]) if len(queries) else \
                np.zeros((0, 32), np.uint8)
            d_q = self.shard.to_device(rows)
        cur = np.asarray(current_frame_ids, np.int32)
        rec, best = self.merged_device(d_q, nq, cur)
        K = self.n_keyframes
        rec_h, best_h = self._to_host(rec.reshape(-1), best)  # both downloads queued, one synchronisation
        # a merged record is four int32 in lc_kf_result order: reinterpret instead of copying field by field
        out = rec_h.view(np.int32).reshape(len(nq), -1, 4)[:, :K]
        out = (out.copy() if copy else np.ascontiguousarray(out)).view(_native.KF_RESULT).reshape(len(nq), K)
        b = best_h.view(np.uint64)
        offs = np.concatenate([[0], np.cumsum(nq)])
        return out, [b[offs[i]:offs[i + 1]].copy() if copy else b[offs[i]:offs[i + 1]] for i in range(len(nq))]

    def _to_host(self, rec_flat, best):
        """Device results -> numpy through pinned staging (CUDA tensors) or directly (the CPU test double).  Two staging
        sets take turns, so that what the previous call returned (copy=False) survives this one."""
        if not rec_flat.is_cuda:
            return rec_flat.numpy(), best.numpy()
        need = (rec_flat.numel(), best.numel())
        pool = getattr(self, "_pin_pool", None)
        if pool is None:
            pool = self._pin_pool = [None, None]
            self._pin_turn = 0
        self._pin_turn ^= 1
        cur = pool[self._pin_turn]
        if cur is None or cur[0].numel() < need[0] or cur[1].numel() < need[1]:
            cur = pool[self._pin_turn] = (torch.empty(max(need[0], 1), dtype=torch.int32).pin_memory(),
                                          torch.empty(max(need[1], 1), dtype=torch.int64).pin_memory())
        r, b = cur[0][:need[0]], cur[1][:need[1]]
        r.copy_(rec_flat, non_blocking=True)
        b.copy_(best, non_blocking=True)
        torch.cuda.current_stream(rec_flat.device).synchronize()
        return r.numpy(), b.numpy()

    def merged_device(self, d_q, nq, cur):
        """Device-resident result of the sharded check: (records int32 [F, slots*world, 4] in global keyframe
        order (padded with sentinels past K), best int64 [sum nq]).  The exchange happens here: one all_gather of
        every rank's [records | bests] contribution, then one merge kernel."""
        slots = (self.n_keyframes + self.world - 1) // self.world
        if self.world == 1:
            return self.shard.loop_check(d_q, nq, cur, slots)
        F, rows = len(nq), int(np.sum(nq))
        rec_bytes = F * slots * 16
        contrib_bytes = (rec_bytes + rows * 8 + 15) & ~15
        if self._peer_exchange(contrib_bytes):
            # no collective: the contribution goes into this rank's exchange slot, one kernel signals, waits and merges
            # while it reads the peers' slots over NVLink
            self._seq += 1
            contrib = self.shard.exchange_slot_tensor(self._seq)
            rec = contrib[:rec_bytes].view(torch.int32).view(F, slots, 4)
            best = contrib[rec_bytes:rec_bytes + rows * 8].view(torch.int64)
            self.shard.loop_check(d_q, nq, cur, slots, out=(rec, best))
            return self.shard.merge_peers(self._seq, F, slots, self.world, rows)
        new_buffer = getattr(self.shard, "new_buffer", lambda n: torch.empty(int(n), dtype=torch.uint8))
        contrib = new_buffer(contrib_bytes)
        rec = contrib[:rec_bytes].view(torch.int32).view(F, slots, 4)
        best = contrib[rec_bytes:rec_bytes + rows * 8].view(torch.int64)
        self.shard.loop_check(d_q, nq, cur, slots, out=(rec, best))
        gathered = new_buffer(self.world * contrib_bytes)
        dist.all_gather_into_tensor(gathered, contrib, group=self.group)
        merge = getattr(self.shard, "merge", merge_torch)
        return merge(gathered, contrib_bytes, self.world, F, slots, rows)


# ---------------------------------------------------------------------------------------------------------------------
# The shipped pipeline's SIFT / L2 loop search (src/main.cpp:1375-1392) over the same partitioning: view k lives on rank
# k % world; a check is independent per view, so the only exchange is an all_gather of the 16-byte per-view records.

_L2_SENTINEL = (-1, 0, -1, 0)


class NativeL2Shard:
    """This rank's views in an lc_ctx (lc_l2_* entry points)."""

    def __init__(self, rank, world, device, **cfg):
        self.device = torch.device("cuda", device)
        self.ctx = _native.Context(device=device, shard_rank=rank, shard_world=world, **cfg)

    def append(self, frame_id, desc, n):
        if desc is None:
            self.ctx.l2_db_append_meta(frame_id, n)
        else:
            self.ctx.l2_db_append(frame_id, desc)

    def loop_check(self, q, current_frame_id, loop_gap, ratio, min_rows):
        return self.ctx.l2_loop_check(q, current_frame_id, loop_gap, ratio, min_rows)


class ShardedL2DB:
    """Every rank calls every method with the same arguments (SPMD)."""

    def __init__(self, group=None, device=None, shard=None, **cfg):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if shard is None:
            device = device if device is not None else torch.cuda.current_device()
            shard = NativeL2Shard(self.rank, self.world, device, **cfg)
        self.shard = shard
        self.n_views = 0

    def owns(self, k):
        return k % self.world == self.rank

    def append(self, frame_id, desc, n=None):
        """desc ((n, 128) SIFT descriptors, u8 or f32) may be None on ranks that do not own this view (then n is required)."""
        k = self.n_views
        own = self.owns(k)
        if own and desc is None:
            raise ValueError(f"rank {self.rank} owns view {k} but got no descriptors")
        self.shard.append(frame_id, desc if own else None, len(desc) if desc is not None else int(n))
        self.n_views += 1
        return k

    def loop_check(self, q, current_frame_id, loop_gap, ratio=0.7, min_rows=100):
        """L2_RESULT records of all views in global order: n_good = matches that survive the ratio test, -1 where the
        pair is skipped (too recent, or fewer than min_rows rows on either side)."""
        local = np.ascontiguousarray(self.shard.loop_check(q, current_frame_id, loop_gap, ratio, min_rows), _native.L2_RESULT)
        if self.world == 1:
            return local
        slots = (self.n_views + self.world - 1) // self.world
        dev = getattr(self.shard, "device", None)
        mine = torch.tensor(_L2_SENTINEL, dtype=torch.int32).repeat(slots, 1)
        mine[:len(local)] = torch.from_numpy(local.view(np.int32).reshape(-1, 4))
        if dev is not None:
            mine = mine.to(dev)
        gathered = torch.empty((self.world * slots, 4), dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        # [world, slots, 4] -> [slots, world, 4]: global view k = slot * world + rank
        merged = gathered.view(self.world, slots, 4).permute(1, 0, 2).reshape(slots * self.world, 4)[:self.n_views].contiguous()
        return merged.cpu().numpy().view(_native.L2_RESULT).reshape(-1)
