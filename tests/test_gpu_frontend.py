"""CT wire-format front end on the GPU (ctmr_process_raw, include/ctmr_frontend.h) against the oracle's
restatement of GetRawEntries' JSON/base64 decode + ct.LogEntryFromLeaf + insertCTWorker + Store
(cmd/ct-fetch/ct-fetch.go:424,446-484,191-245): bit-exact on every output.  Marked gpu."""
import base64
import hashlib
import os
import struct

import numpy as np
import pytest

from conftest import NOW_NS, README_FILTER

pytestmark = pytest.mark.gpu

PATH_FIELDS = ("status", "exp_hour", "serial_off", "serial_len", "was_unknown", "first_issuer_hour")
TS0 = 1_690_000_000_000


@pytest.fixture(scope="module")
def eng():
    from ct_mapreduce_b200 import build, engine
    build.build()
    return engine


def synth_pages(ora, n, seed=7, page=1000, malformed=True, **cfg_kw):
    """n synthetic entries as get-entries bodies: x509 and precert entries, chains of 0-2 certificates,
    duplicates (dup_mode) and, optionally, one of every malformation the front end classifies."""
    from ct_mapreduce_b200 import frontend as fe
    cfg = ora.synth_cfg(n, **cfg_kw)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    issuers = [iblob[int(ioffs[k]):int(ioffs[k + 1])].tobytes() for k in range(ioffs.size - 1)]
    rng = np.random.default_rng(seed)
    kinds = rng.integers(0, 100, n)
    entries = []
    for i in range(n):
        leaf = blob[int(offs[i]):int(offs[i + 1])].tobytes()
        ca = issuers[int(idx[i])]
        chain = [ca, issuers[(int(idx[i]) + 1) % len(issuers)]][: 1 + int(kinds[i]) % 2]
        k = int(kinds[i])
        precert = k % 3 == 0
        if precert:
            li = fe.merkle_tree_leaf_precert(TS0 + i, hashlib.sha256(ca).digest(), fe.tbs_of(leaf), b"" if k % 5 else b"\x00\x01\x02")
            ed = fe.precert_chain_entry(leaf, chain)
        else:
            li = fe.merkle_tree_leaf_x509(TS0 + i, leaf, b"" if k % 5 else b"\xaa")
            ed = fe.certificate_chain(chain)
        if malformed and i % 97 == 13:
            m = (i // 97) % 12
            if m == 0: li = li + b"\x00"                                        # trailing data
            elif m == 1: li = li[:-2]                                           # truncated
            elif m == 2: li = li[:10] + struct.pack(">H", 2 + (i % 5) * 0x2000) + li[12:]  # unknown entry types (incl. 0x8002...)
            elif m == 3: li = b"\x00\x01" + li[2:]                              # leaf_type
            elif m == 4: ed = ed + b"\x01"
            elif m == 5: ed = ed[:-1]
            elif m == 6: ed = (fe.precert_chain_entry(leaf, []) if precert else fe.certificate_chain([]))  # empty chain: NO_ISSUER
            elif m == 7:                                                        # fatal error inside the certificate / TBS
                b = bytearray(li); b[(47 if precert else 15) + 5] ^= 0xFF; li = bytes(b)
            elif m == 8:                                                        # Chain[0] does not parse
                bad = bytearray(ca); bad[6] ^= 0xFF
                ed = fe.precert_chain_entry(leaf, [bytes(bad)]) if precert else fe.certificate_chain([bytes(bad)])
            elif m == 9: li = b""
            elif m == 10: ed = b""
            elif m == 11 and precert:                                           # pre_certificate does not parse: worker's "Problem decoding"
                bad = bytearray(leaf); bad[9] ^= 0xFF
                ed = fe.precert_chain_entry(bytes(bad), chain)
        entries.append((li, ed))
    bodies = [fe.get_entries_body(entries[a:a + page]) for a in range(0, n, page)]
    text = bytearray(b"HTTP/1.1 200 OK\r\n\r\n".join(bodies))
    los, lls, xos, xls = [], [], [], []
    base = 0
    for bd in bodies:
        lo, ll, xo, xl = fe.find_entry_spans(bd, base)
        los.append(lo); lls.append(ll); xos.append(xo); xls.append(xl)
        base += len(bd) + len(b"HTTP/1.1 200 OK\r\n\r\n")
    lo, ll, xo, xl = (np.concatenate(v) for v in (los, lls, xos, xls))
    if malformed:  # base64 damage, applied to the text itself
        for j, i in enumerate(range(41, n, 211)):
            o, l = (int(lo[i]), int(ll[i])) if j % 2 else (int(xo[i]), int(xl[i]))
            if l < 8:
                continue
            w = j % 4
            if w == 0: text[o + l // 2] = ord("-")
            elif w == 1: text[o + 3] = ord("=")
            elif w == 2: ll[i] -= 1                                             # not a whole number of quanta
            else: text[o + l - 5] = 0x80
    return bytes(text), lo, ll, xo, xl, cfg


def digest_of_issuer_der(ora, der):
    rc, c = ora.parse_cert(der)
    if rc != 0:
        return None
    return ora.issuer_id(der[c.spki_off:c.spki_off + c.spki_len])[0]


def check(eng, ora, db, odb, text, lo, ll, xo, xl, now_ns=NOW_NS, want_meta=False):
    r_o = ora.raw_process(odb, text, lo, ll, xo, xl, now_ns)
    r_g = db.store_raw_entries(text, lo, ll, xo, xl, now_ns, want_meta=want_meta)
    for f in ("entry_status", "entry_type"):
        a, b = getattr(r_g, f), getattr(r_o, f)
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (f, bad[:10], a[bad[:10]], b[bad[:10]])
    hdr = (r_o.entry_status != ora.FE_BAD_BASE64) & (r_o.entry_status != ora.FE_BAD_LEAF)
    assert np.array_equal(r_g.timestamp_ms[r_o.entry_status == 0], r_o.timestamp_ms[r_o.entry_status == 0])
    assert hdr.any()
    located = (r_o.entry_status == 0) | ((r_o.entry_status == ora.FE_BAD_CERT) & (r_o.entry_type == 0))
    for f in ("leaf_src", "leaf_off", "leaf_len"):
        assert np.array_equal(getattr(r_g, f)[located], getattr(r_o, f)[located]), f
    assert not r_g.leaf_len[~located].any()
    fields = PATH_FIELDS + (("issuer_name_off", "issuer_name_len", "crldp_off", "crldp_len", "first_issuer_dn", "first_crldp") if want_meta else ())
    for f in fields:
        a, b = getattr(r_g.path, f), getattr(r_o.path, f)
        if f in ("exp_hour", "serial_off", "serial_len", "issuer_name_off", "issuer_name_len", "crldp_off", "crldp_len"):
            m = r_o.path.status != 1
            a, b = a[m], b[m]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (f, bad[:10], a[bad[:10]], b[bad[:10]])
    bad = np.nonzero((r_g.path.sha256 != r_o.path.sha256).any(axis=1))[0]
    assert bad.size == 0, ("sha256", bad[:10])
    # Chain[0] -> Issuer.ID(): per entry, through the digests
    memo = {}
    for i in np.nonzero(r_o.path.status == 0)[0][:: max(1, lo.size // 3000)]:
        der = r_o.issuer_der[i]
        want = memo.get(der) or memo.setdefault(der, digest_of_issuer_der(ora, der))
        assert db.issuer_digest(int(r_g.issuer[i])) == want, i
    return r_g, r_o


def test_pages_with_every_entry_kind_and_malformation(eng, ora):
    n = 20000
    text, lo, ll, xo, xl, _ = synth_pages(ora, n, dup_mode=1)
    odb = ora.DB(README_FILTER, False)
    with eng.GpuCertDatabase(issuer_cn_filter=README_FILTER, table_capacity=1 << 18) as db:
        r_g, r_o = check(eng, ora, db, odb, text, lo, ll, xo, xl, want_meta=True)
        assert {k: v for k, v in db.issuer_counts().items() if v} == odb.issuer_counts()
        assert np.array_equal(db.status_counters(), odb.filter_counters())
        seen = set(r_o.entry_status.tolist())
        assert seen == {0, 1, 2, 3, 4, 5}, seen                      # the corpus reaches every CTMR_FE_* code
        assert set(r_o.path.status.tolist()) >= {0, 1, 2, 3, 5, 6}   # ... and the worker's own outcomes behind it
        fe_ms, path_ms, launches = db.frontend_profile_last()
        assert fe_ms > 0 and path_ms > 0 and launches >= 7


def test_chunks_calls_and_scattered_strings(eng, ora, monkeypatch):
    """Small staging budget: many chunks per call, several calls per ctx (the known-certificate set and the
    issuer mirror persist), and strings handed over in an order that forces the host packing path."""
    monkeypatch.setenv("CTMR_FE_TEXT_CAP", str(1 << 20))
    n = 6000
    text, lo, ll, xo, xl, _ = synth_pages(ora, n, seed=11, page=500, dup_mode=1)
    odb = ora.DB(b"", True)
    with eng.GpuCertDatabase(log_expired_entries=True, table_capacity=1 << 16, max_batch_entries=700) as db:
        cuts = [0, 1, 1500, 1501, 4000, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            sl = np.arange(a, b)
            if a == 1501:  # shuffled: the strings of one chunk lie megabytes apart -> packed on the host
                sl = np.random.default_rng(3).permutation(sl)
            check(eng, ora, db, odb, text, lo[sl].copy(), ll[sl].copy(), xo[sl].copy(), xl[sl].copy())
        assert {k: v for k, v in db.issuer_counts().items() if v} == odb.issuer_counts()


def test_front_end_and_leaf_api_share_one_known_set(eng, ora):
    """The same certificates through ctmr_process_batch first and as get-entries pages second are all known."""
    n = 3000
    text, lo, ll, xo, xl, cfg = synth_pages(ora, n, malformed=False)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    with eng.GpuCertDatabase(log_expired_entries=True, table_capacity=1 << 16) as db:
        r1 = db.store_batch(blob, offs, iblob, ioffs, idx, NOW_NS)
        r2 = db.store_raw_entries(text, lo, ll, xo, xl, NOW_NS)
        assert np.array_equal(r1.status, r2.path.status) and np.array_equal(r1.sha256, r2.path.sha256)
        assert not r2.path.was_unknown.any() and int(r1.was_unknown.sum()) == int((r1.status == 0).sum())
        assert np.array_equal(r2.timestamp_ms, TS0 + np.arange(n, dtype=np.uint64))


def test_pem_of_new_certificates_from_raw_pages(eng, ora):
    """The decoded DER exists only in HBM: its PEM (Store's argument to StoreCertificatePEM) comes back for NEW certificates."""
    from conftest import go_pem
    n = 4000
    text, lo, ll, xo, xl, _ = synth_pages(ora, n, seed=5, dup_mode=1)
    odb = ora.DB(b"", True)
    r_o = ora.raw_process(odb, text, lo, ll, xo, xl, NOW_NS)
    with eng.GpuCertDatabase(log_expired_entries=True, table_capacity=1 << 16, max_batch_entries=900) as db:
        r_g = db.store_raw_entries(text, lo, ll, xo, xl, NOW_NS, want_pem=True)
    assert np.array_equal(r_g.path.was_unknown, r_o.path.was_unknown) and 0 < int(r_o.path.was_unknown.sum()) < n
    for i in range(n):
        assert r_g.path.pem_of(i) == (go_pem(r_o.leaves[i]) if r_o.path.was_unknown[i] else b""), i


def test_empty_and_tiny_batches(eng, ora):
    from ct_mapreduce_b200 import frontend as fe
    with eng.GpuCertDatabase(table_capacity=1 << 12) as db:
        z64, z32 = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        r = db.store_raw_entries(b"x", z64, z32, z64, z32, NOW_NS)
        assert r.entry_status.size == 0
        # two empty strings: a well-formed request for an ill-formed entry
        r = db.store_raw_entries(b"x", np.zeros(1, np.uint64), np.zeros(1, np.uint32), np.zeros(1, np.uint64), np.zeros(1, np.uint32), NOW_NS)
        assert r.entry_status.tolist() == [2] and r.path.status.tolist() == [1]
