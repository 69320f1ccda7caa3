"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the full CUDA path sharded over
ranks with NCCL -- K_map per rank, key routing, owner-side reduce, bits back, all-reduced histogram --
against the oracle run sequentially over the whole corpus."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import NOW_NS, README_FILTER, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, chunks, q, fixed=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from ct_mapreduce_b200 import capi, engine, sharded
    cfg = capi.synth_cfg(n, len_mode=1, len_lo=512, len_hi=4096, dup_mode=1)
    iblob, ioffs = engine.synth_issuers(cfg)
    db = engine.GpuCertDatabase(device=rank, table_capacity=1 << 18, issuer_cn_filter=README_FILTER, max_issuers=1024)
    dense = db.register_issuers(iblob, ioffs)
    assert (dense == np.arange(cfg.n_issuers)).all()
    ops = sharded.GpuOps(db)
    red = sharded.ShardedReducer(ops, dev, n_issuer_slots=cfg.n_issuers, fixed_capacity=fixed)
    per = n // chunks
    cnt = per // world
    res = {}
    for ch in range(chunks):
        lo = ch * per + rank * cnt
        blob, offsets, idx, total = engine.synth_corpus_device(cfg, lo, cnt, dev)
        status = torch.empty(cnt, dtype=torch.uint8, device=dev)
        sha = torch.empty((cnt, 32), dtype=torch.uint8, device=dev)
        exp_hour = torch.empty(cnt, dtype=torch.int64, device=dev)
        keys = torch.empty((cnt, 64), dtype=torch.uint8, device=dev)
        wu = torch.empty(cnt, dtype=torch.uint8, device=dev)
        fi = torch.empty(cnt, dtype=torch.uint8, device=dev)
        b = capi.DevBatch()
        b.blob, b.blob_bytes, b.offsets, b.n = blob.data_ptr(), total, offsets.data_ptr(), cnt
        b.issuer_idx, b.issuer_map, b.issuer_map_len = idx.data_ptr(), None, 0
        b.first_index, b.now_unix_ns = lo, NOW_NS
        o = capi.DevOut(status.data_ptr(), sha.data_ptr(), exp_hour.data_ptr(), None, None, None, None, keys.data_ptr())
        ops.map(b, o)
        red.reduce_chunk(keys, cnt, wu, fi)
        torch.cuda.synchronize(dev)
        res[ch] = (lo, status.cpu().numpy(), sha.cpu().numpy(), wu.cpu().numpy(), fi.cpu().numpy())
    counts, stat = red.merged_histogram()
    assert not red.check_overflow()
    db.check_device()
    q.put((rank, res, counts.cpu().numpy(), stat.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("fixed", [False, True], ids=["ragged_all_to_all", "fixed_capacity_all_to_all"])
def test_sharded_gpu_path_matches_sequential_oracle(ora, fixed):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    n, chunks = 24000, 3
    cfg = ora.synth_cfg(n, len_mode=1, len_lo=512, len_hi=4096, dup_mode=1)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    odb = ora.DB(README_FILTER, False)
    want = odb.process(blob, offs, iblob, ioffs, idx, NOW_NS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunks, q, fixed)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = {f: np.zeros_like(getattr(want, f)) for f in ("status", "sha256", "was_unknown", "first_issuer_hour")}
    for rank, res, counts, stat in outs:
        for ch, (lo, st, sha, wu, fi) in res.items():
            got["status"][lo:lo + st.size] = st
            got["sha256"][lo:lo + st.size] = sha
            got["was_unknown"][lo:lo + st.size] = wu
            got["first_issuer_hour"][lo:lo + st.size] = fi
    for f in got:
        a, b = got[f], getattr(want, f)
        bad = np.nonzero((a != b).reshape(n, -1).any(axis=1))[0]
        assert bad.size == 0, (f, bad.size, bad[:12], a[bad[:6]], b[bad[:6]])
    oc = odb.issuer_counts()
    for rank, res, counts, stat in outs:
        assert int(counts.sum()) == sum(oc.values()) == int(want.was_unknown.sum())
        assert np.array_equal(stat.astype(np.uint64), odb.filter_counters())
