"""Multi-GPU orchestration of the CT map/reduce path: one process per GPU, torch.distributed for
the plumbing (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).

The map half is embarrassingly parallel over entry index.  The reduce half has ONE real exchange
step: equal keys can land on different GPUs, so key records are routed to the
owner of their Redis set -- owner = hash(exp_hour, issuer) mod world, i.e. a whole
"serials::<expDate>::<issuer>" set lives on one GPU exactly like a Redis-cluster key -- the owner
resolves lowest-index-wins, membership bits travel back, and a single all-reduce(sum) merges the
per-issuer histograms and status counters at the end of the chunk.  Ownership is disjoint, so the
summed histogram is exact.

`ops` abstracts the five device operations so that the orchestration itself can be exercised on CPU
(tests/test_sharded_gloo.py plugs an emulation; production uses GpuOps over libctmr).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import capi

KEY_BYTES = capi.KEY_BYTES


# ---------------------------------------------------------------------------------------------------------------
# One process per GPU, owners' tables attached over CUDA IPC (include/ctmr.h "ctmr_peer_*"): the exact multi-GPU
# path.  torch.distributed only carries the 128-byte handles at start-up (any backend) and the bench's timing
# reductions; there is no collective on the data path -- K_map inserts into the owner's table over NVLink and the
# ranks meet at barriers kept in peer memory.
# ---------------------------------------------------------------------------------------------------------------
def attach_peers(db, group=None):
    """All ranks: exchange peer handles and attach.  Afterwards db.process_device / store_batch / reset_device /
    peer_allreduce_histogram_device are collective calls (same sequence on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return rank, world
    handles = [None] * world
    dist.all_gather_object(handles, db.peer_export(world), group=group)
    db.peer_attach(rank, world, handles)
    return rank, world


def round_entries(n: int, rounds: int = 0) -> int:
    """E of the collective ctmr_process_device: entries per rank and round (every rank passes the same n)."""
    return -(-n // (rounds or capi.peer_rounds()))


def call_index_span(n: int, world: int, rounds: int = 0) -> int:
    """How far a collective ctmr_process_device call advances the global index (its next call's first_index)."""
    rounds = rounds or capi.peer_rounds()
    return world * rounds * round_entries(n, rounds)


def sequential_order(n: int, world: int, rounds: int = 0):
    """The order in which the sequential reference (numThreads=1) would have to see the entries of one collective
    call for its result to equal the group's: rounds in order, inside a round the ranks in order, inside a rank's
    slice the entries in order.  Yields (rank, lo, hi) slices of each rank's local batch [0, n)."""
    rounds = rounds or capi.peer_rounds()
    e = round_entries(n, rounds)
    for k in range(rounds):
        lo, hi = min(n, k * e), min(n, (k + 1) * e)
        for r in range(world):
            if hi > lo:
                yield r, lo, hi


def host_batch_order(ns, stage_entries: int):
    """The same for the collective HOST-buffer call (ctmr_process_batch with peers attached): rank r passes ns[r]
    entries, a round takes `stage_entries` per rank, every rank runs max(rounds) rounds.  (Byte-budget shrinking of a
    round is not modelled: callers keep entries below config.max_batch_bytes / stage_entries on average.)"""
    rounds = max(1, max(-(-n // stage_entries) for n in ns))
    for k in range(rounds):
        for r, n in enumerate(ns):
            lo, hi = min(n, k * stage_entries), min(n, (k + 1) * stage_entries)
            if hi > lo:
                yield r, lo, hi


class GpuOps:
    """The five device operations, straight onto the C ABI (no arithmetic in Python)."""

    def __init__(self, db):
        self.db = db
        self.device = torch.device("cuda", db.device)

    def _stream(self):
        # torch's default stream has handle 0, which the C ABI reads as "use the ctx stream":
        # pass cudaStreamLegacy (0x1) instead so that the kernels really run on torch's stream
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    def map(self, batch: capi.DevBatch, out: capi.DevOut):
        self.db.map_device(batch, out, self._stream())

    def partition(self, keys, n, world, keys_by_owner, src_pos, owner_counts):
        self.db.partition_keys_device(keys, n, world, keys_by_owner, src_pos, owner_counts, self._stream())

    def partition_fixed(self, keys, n, world, capacity, keys_by_owner, src_pos, overflow):
        self.db.partition_keys_fixed_device(keys, n, world, capacity, keys_by_owner, src_pos, overflow, self._stream())

    def reduce(self, keys, m, was_unknown, first):
        self.db.reduce_device(keys, m, was_unknown, first, self._stream())

    def scatter(self, was_unknown, first, src_pos, m, was_unknown_dst, first_dst):
        self.db.scatter_bits_device(was_unknown, first, src_pos, m, was_unknown_dst, first_dst, self._stream())

    def read_histogram(self, counts_dst, n_slots, status_dst):
        self.db.read_histogram_device(counts_dst, n_slots, status_dst, self._stream())


class ShardedReducer:
    """Reduce half across ranks for one chunk of key records produced by this rank's map half."""

    def __init__(self, ops, device, n_issuer_slots: int, group=None, max_keys: int = 0, fixed_capacity: bool = False,
                 slack: float = 1.25, min_slots: int = 1024):
        """fixed_capacity=True selects the sync-free exchange: every rank sends the same number of key slots
        (`slack` x its fair share, padded with invalid records) to every peer, so that no split size has to
        travel through the host; `check_overflow()` tells afterwards whether some owner got more than that."""
        self.ops, self.device, self.group = ops, torch.device(device), group
        self.fixed_capacity, self.slack, self.min_slots = fixed_capacity, slack, min_slots
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_issuer_slots = n_issuer_slots
        self.hist = torch.zeros(n_issuer_slots + capi.ST_COUNT, dtype=torch.int64, device=self.device)
        self._cap = 0
        self._recv_cap = 0
        if max_keys:
            self._ensure(max_keys, max_keys * 2)

    def _ensure(self, n, recv):
        if n > self._cap:
            self.keys_by_owner = torch.empty((n, KEY_BYTES), dtype=torch.uint8, device=self.device)
            self.src_pos = torch.empty(n, dtype=torch.int32, device=self.device)
            self.bits_back = torch.empty((2, n), dtype=torch.uint8, device=self.device)
            self._cap = n
        if recv > self._recv_cap:
            self.recv_keys = torch.empty((recv, KEY_BYTES), dtype=torch.uint8, device=self.device)
            self.recv_bits = torch.empty((2, recv), dtype=torch.uint8, device=self.device)
            self._recv_cap = recv
        if not hasattr(self, "owner_counts"):
            self.owner_counts = torch.zeros(max(self.world, 1), dtype=torch.int64, device=self.device)

    def reduce_chunk(self, keys: torch.Tensor, n: int, was_unknown: torch.Tensor, first: torch.Tensor):
        """keys: [n, 64] uint8 key records of this rank's entries.  Fills was_unknown/first ([n] uint8)."""
        if self.world == 1:
            self.ops.reduce(keys, n, was_unknown, first)
            return
        if self.fixed_capacity:
            return self._reduce_chunk_fixed(keys, n, was_unknown, first)
        self._ensure(n, self._recv_cap)
        # 1. counting-sort the valid key records by owner rank
        self.ops.partition(keys, n, self.world, self.keys_by_owner, self.src_pos, self.owner_counts)
        recv_dev = torch.empty_like(self.owner_counts)
        dist.all_to_all_single(recv_dev, self.owner_counts, group=self.group)
        send = [int(x) for x in self.owner_counts.tolist()]      # the one host sync per chunk
        recv = [int(x) for x in recv_dev.tolist()]
        m_send, m_recv = sum(send), sum(recv)
        self._ensure(n, max(m_recv, 1))
        # 2. key exchange: 64-byte records to their owners (all-to-all over NVSwitch)
        dist.all_to_all_single(self.recv_keys[:m_recv], self.keys_by_owner[:m_send], recv, send, group=self.group)
        # 3. owners insert + resolve lowest-index-wins + count
        rb_unknown, rb_first = self.recv_bits[0, :max(m_recv, 1)], self.recv_bits[1, :max(m_recv, 1)]
        self.ops.reduce(self.recv_keys, m_recv, rb_unknown, rb_first)
        # 4. membership bits back to the entries' home ranks
        bu, bf = self.bits_back[0, :max(m_send, 1)], self.bits_back[1, :max(m_send, 1)]
        dist.all_to_all_single(bu[:m_send], rb_unknown[:m_recv], send, recv, group=self.group)
        dist.all_to_all_single(bf[:m_send], rb_first[:m_recv], send, recv, group=self.group)
        was_unknown.zero_()
        first.zero_()
        self.ops.scatter(bu, bf, self.src_pos, m_send, was_unknown, first)

    def _reduce_chunk_fixed(self, keys, n, was_unknown, first):
        """The same routing without a host round trip: equal-split all-to-alls over fixed-capacity buckets."""
        W = self.world
        cap = max(64, -(-int(n * self.slack / W + self.min_slots) // 64) * 64)   # slots per (source, owner) pair
        tot = W * cap
        if getattr(self, "_fixed_cap", 0) < tot:
            self.f_send = torch.empty((tot, KEY_BYTES), dtype=torch.uint8, device=self.device)
            self.f_recv = torch.empty((tot, KEY_BYTES), dtype=torch.uint8, device=self.device)
            self.f_src = torch.empty(tot, dtype=torch.int32, device=self.device)
            self.f_bits = torch.empty((4, tot), dtype=torch.uint8, device=self.device)  # unknown/first at the owner, then back home
            self._fixed_cap = tot
        if not hasattr(self, "overflow"):
            self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        send, recv, src = self.f_send[:tot], self.f_recv[:tot], self.f_src[:tot]
        self.ops.partition_fixed(keys, n, W, cap, send, src, self.overflow)
        dist.all_to_all_single(recv, send, group=self.group)                 # bucket w of every rank -> rank w
        ou, of_, hu, hf = (self.f_bits[k, :tot] for k in range(4))
        self.ops.reduce(recv, tot, ou, of_)                                  # invalid (padding) records are skipped
        dist.all_to_all_single(hu, ou, group=self.group)                     # bits back, same geometry
        dist.all_to_all_single(hf, of_, group=self.group)
        was_unknown.zero_()
        first.zero_()
        self.ops.scatter(hu, hf, src, tot, was_unknown, first)               # padding slots carry src = 0xFFFFFFFF

    def check_overflow(self):
        """True if any fixed-capacity bucket overflowed on any rank since the last check (one host sync; call it once
        per step or per run, not per chunk).  The affected chunks must be re-routed with the variable-size path."""
        if not self.fixed_capacity or not hasattr(self, "overflow"):
            return False
        flag = self.overflow.clone()
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        self.overflow.zero_()
        return bool(int(flag.item()))

    def merged_histogram(self):
        """One all-reduce(sum) of [per-issuer unique counts || status counters] (north_star: the
        single NCCL allreduce at the end of each chunk).  Returns (counts[n_slots], status[8])."""
        self.ops.read_histogram(self.hist[: self.n_issuer_slots], self.n_issuer_slots, self.hist[self.n_issuer_slots:])
        if self.world > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM, group=self.group)
        return self.hist[: self.n_issuer_slots], self.hist[self.n_issuer_slots:]
