"""The N>1 path on CPU: world_size-2 gloo run of ct_mapreduce_b200.sharded.ShardedReducer.

The orchestration under test is the production code (key routing by owner, all-to-all with ragged
splits, membership bits travelling back, scatter to entry order, one all-reduce of the histograms).
The five device operations are replaced by a small numpy emulation (this file) so that no GPU is
needed; the expected result is the oracle run sequentially over the WHOLE corpus, i.e. what the
reference's single Redis would have answered.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import NOW_NS, README_FILTER, ROOT

M64 = (1 << 64) - 1


def mix64(z):
    z &= M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def key_owner(exp_hour, issuer, world):
    """ctmr::key_owner (csrc/ctmr_device.cuh): owner of the Redis set serials::<expDate>::<issuer>."""
    return mix64(((issuer & 0xFFFFFFFF) << 32) | (exp_hour & 0xFFFFFFFF)) % world


class EmulatedOps:
    """numpy stand-in for GpuOps: same contracts as the CUDA kernels behind the C ABI."""

    def __init__(self, n_issuers):
        from ct_mapreduce_b200 import capi
        self.kd = capi.KEY_DTYPE
        self.table = {}   # key body bytes -> lowest index
        self.pairs = {}   # (issuer, hour) -> lowest index among was-unknown entries
        self.counts = np.zeros(n_issuers, np.int64)
        self.status = np.zeros(8, np.int64)

    def _recs(self, t, n):
        return t.numpy()[:n].reshape(-1).view(self.kd)[:n]

    def partition(self, keys, n, world, keys_by_owner, src_pos, owner_counts):
        recs = self._recs(keys, n)
        owners = np.array([key_owner(int(r["exp_hour"]), int(r["issuer"]), world) if r["valid"] else -1 for r in recs])
        out = keys_by_owner.numpy().reshape(-1).view(self.kd)
        sp = src_pos.numpy()
        cnt = np.zeros(world, np.int64)
        pos = 0
        for w in range(world):
            idx = np.nonzero(owners == w)[0]
            out[pos:pos + idx.size] = recs[idx]
            sp[pos:pos + idx.size] = idx
            cnt[w] = idx.size
            pos += idx.size
        owner_counts.copy_(torch.from_numpy(cnt))

    def partition_fixed(self, keys, n, world, capacity, keys_by_owner, src_pos, overflow):
        """ctmr_partition_keys_fixed_device: bucket w = slots [w*capacity, (w+1)*capacity), padding invalid / src -1."""
        recs = self._recs(keys, n)
        owners = np.array([key_owner(int(r["exp_hour"]), int(r["issuer"]), world) if r["valid"] else -1 for r in recs])
        out = keys_by_owner.numpy().reshape(-1).view(self.kd)
        out[:world * capacity] = np.zeros(1, self.kd)
        sp = src_pos.numpy()
        sp[:world * capacity] = -1
        for w in range(world):
            idx = np.nonzero(owners == w)[0]
            if idx.size > capacity:
                overflow.fill_(1)
                idx = idx[:capacity]
            out[w * capacity:w * capacity + idx.size] = recs[idx]
            sp[w * capacity:w * capacity + idx.size] = idx

    def reduce(self, keys, m, was_unknown, first):
        recs = self._recs(keys, m)
        bodies = [r.tobytes()[8:56] for r in recs]
        for r, b in zip(recs, bodies):   # insert phase: lowest index wins
            if r["valid"]:
                self.table[b] = min(self.table.get(b, 1 << 63), int(r["index"]))
        wu, fi = was_unknown.numpy(), first.numpy()
        for j, (r, b) in enumerate(zip(recs, bodies)):  # resolve phase
            u = bool(r["valid"]) and self.table[b] == int(r["index"])
            wu[j] = u
            if u:
                self.counts[int(r["issuer"])] += 1
                pk = (int(r["issuer"]), int(r["exp_hour"]))
                self.pairs[pk] = min(self.pairs.get(pk, 1 << 63), int(r["index"]))
        for j, r in enumerate(recs):
            fi[j] = bool(wu[j]) and self.pairs[(int(r["issuer"]), int(r["exp_hour"]))] == int(r["index"])

    def scatter(self, was_unknown, first, src_pos, m, was_unknown_dst, first_dst):
        sp = src_pos.numpy()[:m].astype(np.int64)
        live = sp >= 0  # 0xFFFFFFFF (= -1 as int32) marks an unused slot of the fixed-capacity layout
        was_unknown_dst.numpy()[sp[live]] = was_unknown.numpy()[:m][live]
        first_dst.numpy()[sp[live]] = first.numpy()[:m][live]

    def read_histogram(self, counts_dst, n_slots, status_dst):
        counts_dst.copy_(torch.from_numpy(self.counts[:n_slots]))
        status_dst.copy_(torch.from_numpy(self.status))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, chunks, q, fixed=False, slack=1.25, min_slots=1024):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ct_mapreduce_b200 import capi, sharded
    from oracle import oracle
    cfg = oracle.synth_cfg(n, len_mode=1, len_lo=512, len_hi=2048, dup_mode=1)
    iblob, ioffs = oracle.synth_issuers(cfg)
    ops = EmulatedOps(cfg.n_issuers)
    red = sharded.ShardedReducer(ops, "cpu", n_issuer_slots=cfg.n_issuers, fixed_capacity=fixed, slack=slack, min_slots=min_slots)
    per = n // chunks
    res = {}
    for ch in range(chunks):
        # chunk ch of the log, split contiguously over the ranks (SURVEY §8(e) partitioning)
        lo = ch * per + rank * (per // world)
        cnt = per // world
        blob, offs, idx = oracle.synth_corpus(cfg, lo, cnt)
        # "map half": what K_map would emit for these entries (status, exp_hour, serial), via the oracle
        m = oracle.DB(README_FILTER, False).process(blob, offs, iblob, ioffs, idx, NOW_NS)
        recs = np.zeros(cnt, capi.KEY_DTYPE)
        recs["index"] = np.arange(lo, lo + cnt)
        recs["exp_hour"] = m.exp_hour
        recs["issuer"] = idx
        recs["valid"] = m.status == 0
        for i in range(cnt):
            sl = int(m.serial_len[i])
            recs["serial_len"][i] = sl
            a = int(offs[i]) + int(m.serial_off[i])
            recs["serial"][i, :sl] = blob[a:a + sl]
        ops.status += np.bincount(m.status, minlength=8)
        keys = torch.from_numpy(recs.view(np.uint8).reshape(cnt, 64).copy())
        wu = torch.zeros(cnt, dtype=torch.uint8)
        fi = torch.zeros(cnt, dtype=torch.uint8)
        red.reduce_chunk(keys, cnt, wu, fi)
        res[ch] = (lo, wu.numpy().copy(), fi.numpy().copy())
    counts, status = red.merged_histogram()
    q.put((rank, res, counts.numpy().copy(), status.numpy().copy(), red.check_overflow()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fixed", [False, True], ids=["ragged_all_to_all", "fixed_capacity_all_to_all"])
def test_two_rank_reduce_matches_sequential_oracle(ora, fixed):
    n, chunks, world = 2400, 3, 2
    cfg = ora.synth_cfg(n, len_mode=1, len_lo=512, len_hi=2048, dup_mode=1)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    odb = ora.DB(README_FILTER, False)
    want = odb.process(blob, offs, iblob, ioffs, idx, NOW_NS)   # the single-Redis answer

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunks, q, fixed)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not any(o[4] for o in outs)  # no bucket overflowed (never set on the ragged path)
    got_wu = np.zeros(n, np.uint8)
    got_fi = np.zeros(n, np.uint8)
    for rank, res, counts, status, _ in outs:
        for ch, (lo, wu, fi) in res.items():
            got_wu[lo:lo + wu.size] = wu
            got_fi[lo:lo + fi.size] = fi
    assert np.array_equal(got_wu, want.was_unknown)
    assert np.array_equal(got_fi, want.first_issuer_hour)
    # duplicates straddle ranks AND chunks, yet every kept certificate is unknown exactly once
    assert int(got_wu.sum()) * 2 == int((want.status == 0).sum())
    # merged histogram (identical on both ranks after the all-reduce) = Count()-sum per issuer
    dense_digest = {}
    for k in range(cfg.n_issuers):
        der = iblob[ioffs[k]:ioffs[k + 1]].tobytes()
        rc, c = ora.parse_cert(der)
        dense_digest[k] = ora.issuer_id(der[c.spki_off:c.spki_off + c.spki_len])[0]
    oc = odb.issuer_counts()
    for rank, res, counts, status, _ in outs:
        assert {dense_digest[k]: int(v) for k, v in enumerate(counts) if v} == oc
        assert np.array_equal(status, odb.filter_counters().astype(np.int64))


@pytest.mark.timeout(300)
def test_fixed_capacity_overflow_is_reported_on_every_rank():
    """Buckets sized below the fair share must overflow; the flag is all-reduced so that every rank re-routes."""
    n, chunks, world = 1200, 1, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunks, q, True, 0.05, 0)) for r in range(world)]  # 64 slots per bucket
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[4] for o in outs)


def test_key_owner_is_a_function_of_the_redis_set():
    # all serials of one (expDate, issuer) set land on one rank; sets spread evenly
    owners = [key_owner(h, k, 8) for h in range(490000, 490400) for k in range(16)]
    hist = np.bincount(owners, minlength=8)
    assert hist.min() > 0.8 * hist.mean() and hist.max() < 1.2 * hist.mean()
    assert key_owner(-5, 7, 8) == key_owner(-5 & 0xFFFFFFFF, 7, 8)
