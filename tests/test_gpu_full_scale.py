"""BASELINE.json sizes (10 M entries on one GPU) through size-independent properties: the oracle cannot
run 10 M sequential entries in test time, so the checks here are the ones the domain offers --
generator-side truth for every entry's status / expiry hour, "every kept certificate is unknown
exactly once" under 50 % duplicates, counts summing to membership bits, idempotence of a replay,
and hashlib fingerprints on a random sample."""
import hashlib

import numpy as np
import pytest

from conftest import NOW_NS, NOW_SEC, README_FILTER

pytestmark = pytest.mark.gpu


def _run(n, kw, flt, log_expired, table_cap):
    import torch
    from ct_mapreduce_b200 import capi, engine
    dev = torch.device("cuda:0")
    cfg = capi.synth_cfg(n, **kw)
    blob, offsets, idx, total = engine.synth_corpus_device(cfg, 0, n, dev)
    iblob, ioffs = engine.synth_issuers(cfg)
    db = engine.GpuCertDatabase(table_capacity=table_cap, issuer_cn_filter=flt, log_expired_entries=log_expired, max_issuers=1024)
    db.register_issuers(iblob, ioffs)
    t = dict(status=torch.empty(n, dtype=torch.uint8, device=dev), sha=torch.empty((n, 32), dtype=torch.uint8, device=dev),
             exp_hour=torch.empty(n, dtype=torch.int64, device=dev), wu=torch.empty(n, dtype=torch.uint8, device=dev),
             fi=torch.empty(n, dtype=torch.uint8, device=dev))

    def go(first_index):
        b = capi.DevBatch()
        b.blob, b.blob_bytes, b.offsets, b.n = blob.data_ptr(), total, offsets.data_ptr(), n
        b.issuer_idx, b.issuer_map, b.issuer_map_len = idx.data_ptr(), None, 0
        b.first_index, b.now_unix_ns = first_index, NOW_NS
        o = capi.DevOut(t["status"].data_ptr(), t["sha"].data_ptr(), t["exp_hour"].data_ptr(), None, None,
                        t["wu"].data_ptr(), t["fi"].data_ptr(), None)
        db.process_device(b, o)
        db.check_device()
        torch.cuda.synchronize()

    # generator-side truth
    cert_id = torch.empty(n, dtype=torch.int64, device=dev)
    not_after = torch.empty(n, dtype=torch.int64, device=dev)
    bc_mode = torch.empty(n, dtype=torch.uint8, device=dev)
    import ctypes as C
    rc = capi.load().ctmr_synth_truth_device(C.byref(cfg), 0, n, cert_id.data_ptr(), not_after.data_ptr(), bc_mode.data_ptr(), 1)
    assert rc == 0
    torch.cuda.synchronize()
    return db, cfg, t, go, (blob, offsets, idx, total), (cert_id, not_after, bc_mode)


@pytest.mark.timeout(600)
def test_config2_10M_fingerprint_and_dedup():
    """configs[1]: 10 M x ~1.5 KB, SHA-256 + KnownCertificates dedup (no CN filter, expired kept)."""
    import torch
    n = 10_000_000
    db, cfg, t, go, (blob, offsets, idx, total), (cert_id, not_after, bc_mode) = _run(n, {}, b"", True, 1 << 25)
    go(0)
    status, wu, fi = t["status"], t["wu"], t["fi"]
    # status truth: only CA:TRUE certificates are filtered in this configuration
    exp_status = torch.where(bc_mode == 2, torch.tensor(2, dtype=torch.uint8, device=status.device),
                             torch.tensor(0, dtype=torch.uint8, device=status.device))
    assert torch.equal(status, exp_status)
    assert torch.equal(t["exp_hour"], torch.div(not_after, 3600, rounding_mode="floor"))
    ok = status == 0
    assert torch.equal(wu.bool(), ok)                        # all certificates distinct -> every stored one is unknown
    assert int(fi.sum()) <= int(wu.sum()) and int((fi.bool() & ~wu.bool()).sum()) == 0
    counts = db.issuer_counts()
    assert sum(counts.values()) == int(wu.sum())
    used, cap = db.table_stats()
    assert used == int(wu.sum())
    # per-issuer counts against a bincount of the generator's issuer index over stored entries
    want = torch.bincount(idx[ok].long(), minlength=cfg.n_issuers).cpu().numpy()
    dig2k = {db.issuer_digest(k): k for k in range(cfg.n_issuers)}
    got = np.zeros(cfg.n_issuers, np.int64)
    for d, c in counts.items():
        got[dig2k[d]] = c
    assert np.array_equal(got, want)
    # first (issuer, hour) bits: exactly one per distinct pair among stored entries
    pairs = (idx[ok].long() << 32) | (t["exp_hour"][ok] & 0xFFFFFFFF)
    assert int(fi.sum()) == int(torch.unique(pairs).numel())
    # fingerprints of a random sample against hashlib
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(n, 2000, replace=False))
    offs_h = offsets.cpu().numpy()
    for i in sample[:2000]:
        a, b = int(offs_h[i]), int(offs_h[i + 1])
        assert t["sha"][i].cpu().numpy().tobytes() == hashlib.sha256(blob[a:b].cpu().numpy().tobytes()).digest()
    # idempotence: replaying the same batch (higher indices) finds everything already known
    prev_counts = counts
    go(n)
    assert int(t["wu"].sum()) == 0 and int(t["fi"].sum()) == 0
    assert db.issuer_counts() == prev_counts
    db.close()


@pytest.mark.timeout(600)
def test_mixed_sizes_50pct_duplicates_4M():
    """configs[4] shape on one GPU: 512 B..8 KB, every certificate exactly twice, README filter."""
    import torch
    n = 4_000_000
    kw = dict(len_mode=1, len_lo=512, len_hi=8192, dup_mode=1)
    db, cfg, t, go, (blob, offsets, idx, total), (cert_id, not_after, bc_mode) = _run(n, kw, README_FILTER, False, 1 << 23)
    go(0)
    status, wu = t["status"], t["wu"]
    # generator-side truth for the filter: CA -> expired -> CN class (k % 4 in {0, 1} passes the README filter)
    k = idx.long()
    exp = torch.zeros(n, dtype=torch.uint8, device=status.device)
    exp[(k % 4) >= 2] = 4
    exp[not_after < NOW_SEC] = 3
    exp[bc_mode == 2] = 2
    assert torch.equal(status, exp)
    ok = status == 0
    # every stored certificate has exactly one twin: exactly half of the stored entries are unknown,
    # and the unknown one is the twin with the LOWER index
    assert int(wu.sum()) * 2 == int(ok.sum())
    ids = cert_id[ok]
    order = torch.argsort(ids, stable=True)
    first_of_pair = torch.zeros(ids.numel(), dtype=torch.bool, device=ids.device)
    first_of_pair[order[0::2]] = True   # stable sort keeps entry order inside a pair
    assert torch.equal(wu[ok].bool(), first_of_pair)
    assert sum(db.issuer_counts().values()) == int(wu.sum())
    db.close()


@pytest.mark.timeout(600)
def test_group_host_call_1M_entries_with_duplicates():
    """The Go-facing multi-GPU entry point at a batch size that takes dozens of rounds: 1 M entries (every certificate
    twice) through ctmr_group_process_batch on every GPU of the box -- or, on a 1-GPU box, on two shards of cuda:0 --
    checked by properties: exactly the EARLIER twin of every kept pair is unknown (twins mostly live on different shards and
    in different rounds), the summed histogram equals the bits, a replay finds everything known, and the group's answers
    equal a single GPU's on the same batch."""
    import torch
    from ct_mapreduce_b200 import capi, engine
    n = 1_000_000
    cfg = capi.synth_cfg(n, dup_mode=1)
    dev = torch.device("cuda:0")
    blob_d, offs_d, idx_d, total = engine.synth_corpus_device(cfg, 0, n, dev)
    blob = blob_d[:total].cpu().numpy()
    offs = offs_d.cpu().numpy().astype(np.uint64)
    idx = idx_d.cpu().numpy().astype(np.uint32)
    cert_id = torch.empty(n, dtype=torch.int64, device=dev)
    na = torch.empty(n, dtype=torch.int64, device=dev)
    bc = torch.empty(n, dtype=torch.uint8, device=dev)
    import ctypes as C
    assert capi.load().ctmr_synth_truth_device(C.byref(cfg), 0, n, cert_id.data_ptr(), na.data_ptr(), bc.data_ptr(), 1) == 0
    torch.cuda.synchronize()
    ids = cert_id.cpu().numpy()
    del blob_d, offs_d, idx_d, cert_id, na, bc
    iblob, ioffs = engine.synth_issuers(cfg)
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() > 1 else [0, 0]
    with engine.GpuCertGroup(devices, table_capacity=1 << 21, issuer_cn_filter=README_FILTER, max_issuers=1024,
                             max_batch_entries=16384) as g:
        r = g.store_batch(blob, offs, iblob, ioffs, idx, NOW_NS)
        ok = r.status == 0
        assert 0 < int(r.was_unknown.sum()) * 2 == int(ok.sum())
        kept = np.nonzero(ok)[0]
        order = kept[np.argsort(ids[kept], kind="stable")]
        first_of_pair = np.zeros(n, bool)
        first_of_pair[order[0::2]] = True           # stable: entry order inside a pair
        assert np.array_equal(r.was_unknown.astype(bool), first_of_pair)
        assert sum(g.issuer_counts().values()) == int(r.was_unknown.sum())
        assert int(g.status_counters()[0]) == int(ok.sum())
        per = [m.table_stats()[0] for m in g.members]
        assert sum(per) == int(r.was_unknown.sum()) and min(per) > 0.3 * max(per)   # the owner hash spreads the sets
        r2 = g.store_batch(blob, offs, iblob, ioffs, idx, NOW_NS, want_sha=False)
        assert not r2.was_unknown.any() and not r2.first_issuer_hour.any()
    with engine.GpuCertDatabase(table_capacity=1 << 21, issuer_cn_filter=README_FILTER, max_issuers=1024) as db:
        s = db.store_batch(blob, offs, iblob, ioffs, idx, NOW_NS)
    for f in ("status", "sha256", "exp_hour", "was_unknown", "first_issuer_hour"):
        assert np.array_equal(getattr(r, f), getattr(s, f)), f
