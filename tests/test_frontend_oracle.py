"""Pins the front-end oracle (oracle/ctmr_oracle_frontend.c) on the CPU.

ct-go v1.1.0 (ct.LogEntryFromLeaf, reference call site cmd/ct-fetch/ct-fetch.go:452) is not vendored under
the reference and there is no Go toolchain, so the pins are: Python's base64 for the string decode
(hypothesis), and hand-built RFC 6962 §3.4 / §4.6 structures around the reference's own golden certificates
(tests/golden/*.pem, lifted from storage/types_test.go) for the TLS framing and the drop decisions."""
import base64
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from ct_mapreduce_b200 import frontend as fe


@settings(max_examples=300, deadline=None)
@given(st.binary(min_size=0, max_size=200))
def test_b64_decode_matches_python(ora, data):
    assert ora.b64_decode(base64.b64encode(data)) == data


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=1, max_size=60), st.integers(0, 59), st.sampled_from(list(b"-_=. \n\r\\\"*\x00\x80\xff")))
def test_b64_decode_rejects_what_go_rejects(ora, data, pos, junk):
    good = bytearray(base64.b64encode(data))
    pos %= len(good)
    if good[pos] == junk:
        return
    good[pos] = junk
    # after the substitution the string is either still canonical padding-wise (junk '=' landing on the last
    # one or two positions of a padded string) or corrupt
    try:
        want = base64.b64decode(bytes(good), validate=True)
        canonical = base64.b64encode(want) == bytes(good) or bytes(good).rstrip(b"=") != bytes(good)
    except Exception:
        want, canonical = None, True
    got = ora.b64_decode(bytes(good))
    if want is None:
        assert got is None
    elif canonical:
        assert got == want


def test_b64_decode_edge_forms(ora):
    assert ora.b64_decode(b"") == b""
    assert ora.b64_decode(b"QQ==") == b"A"
    assert ora.b64_decode(b"QUI=") == b"AB"
    assert ora.b64_decode(b"QUJD") == b"ABC"
    assert ora.b64_decode(b"QR==") == b"A"          # non-zero trailing bits: StdEncoding is not Strict()
    for bad in (b"QQ=", b"QQ", b"Q===", b"QQ=A", b"=QQQ", b"QUJD=", b"QU-D", b"QU_D", b"QUJD\n", b"QQ==QUJD"):
        assert ora.b64_decode(bad) is None, bad


# ---- framing -------------------------------------------------------------------------------------------

TS = 1_700_000_123_456


@pytest.fixture(scope="module")
def certs(golden):
    leaf = golden["kLeadingZeroes"]["der"]
    others = [golden[k]["der"] for k in sorted(golden) if k != "kLeadingZeroes"]
    return leaf, others


def test_x509_entry(ora, certs):
    leaf, (ca, root) = certs[0], certs[1][:2]
    li = fe.merkle_tree_leaf_x509(TS, leaf)
    ed = fe.certificate_chain([ca, root])
    st_, e = ora.entry_from_leaf(li, ed)
    assert st_ == ora.FE_OK
    assert (e.entry_type, e.timestamp_ms, e.leaf_src) == (0, TS, 0)
    assert li[e.leaf_off:e.leaf_off + e.leaf_len] == leaf
    assert ed[e.chain0_off:e.chain0_off + e.chain0_len] == ca and e.chain_count == 2
    # CtExtensions are opaque and may be non-empty
    assert ora.entry_from_leaf(fe.merkle_tree_leaf_x509(TS, leaf, b"\x01\x02\x03"), ed)[0] == ora.FE_OK
    # an empty chain is a valid CertificateChain: the worker then logs "No issuer known" (ct-fetch.go:215-219)
    st_, e = ora.entry_from_leaf(li, fe.certificate_chain([]))
    assert st_ == ora.FE_OK and e.chain_count == 0 and e.chain0_len == 0


def test_precert_entry(ora, certs):
    leaf, (ca, root) = certs[0], certs[1][:2]
    tbs = fe.tbs_of(leaf)
    assert ora.parse_tbs(tbs)[0] == 0
    li = fe.merkle_tree_leaf_precert(TS, bytes(range(32)), tbs)
    ed = fe.precert_chain_entry(leaf, [ca, root])
    st_, e = ora.entry_from_leaf(li, ed)
    assert st_ == ora.FE_OK
    assert (e.entry_type, e.timestamp_ms, e.leaf_src) == (1, TS, 1)
    assert ed[e.leaf_off:e.leaf_off + e.leaf_len] == leaf            # Precert.Submitted (ct-fetch.go:202)
    assert li[e.tbs_off:e.tbs_off + e.tbs_len] == tbs
    assert ed[e.chain0_off:e.chain0_off + e.chain0_len] == ca and e.chain_count == 2
    # ParseTBSCertificate: trailing data after the TBSCertificate is an error -> entry dropped
    assert ora.parse_tbs(tbs + b"\x00")[0] != 0
    assert ora.entry_from_leaf(fe.merkle_tree_leaf_precert(TS, bytes(32), tbs + b"\x00"), ed)[0] == ora.FE_BAD_CERT
    assert ora.entry_from_leaf(fe.merkle_tree_leaf_precert(TS, bytes(32), tbs[:-1]), ed)[0] == ora.FE_BAD_CERT


def test_leaf_framing_errors(ora, certs):
    leaf, (ca, _) = certs[0], certs[1][:2]
    li, ed = fe.merkle_tree_leaf_x509(TS, leaf), fe.certificate_chain([ca])
    assert ora.entry_from_leaf(li + b"\x00", ed)[0] == ora.FE_BAD_LEAF                 # trailing data after MerkleTreeLeaf
    assert ora.entry_from_leaf(li[:-1], ed)[0] == ora.FE_BAD_LEAF                      # extensions length cut
    assert ora.entry_from_leaf(li[:-3], ed)[0] == ora.FE_BAD_LEAF
    assert ora.entry_from_leaf(li[:11], ed)[0] == ora.FE_BAD_LEAF
    assert ora.entry_from_leaf(b"", ed)[0] == ora.FE_BAD_LEAF
    assert ora.entry_from_leaf(b"\x00\x01" + li[2:], ed)[0] == ora.FE_BAD_LEAF         # leaf_type != timestamped_entry
    assert ora.entry_from_leaf(b"\x07" + li[1:], ed)[0] == ora.FE_OK                   # Version is not checked by tls.Unmarshal
    zero = b"\x00\x00" + struct.pack(">QH", TS, 0) + b"\x00\x00\x00" + b"\x00\x00"     # ASN.1Cert<1..>: empty is illegal
    assert ora.entry_from_leaf(zero, ed)[0] == ora.FE_BAD_LEAF
    for t in (2, 0x8000, 0xFFFF):                                                      # JSON type and undefined types
        assert ora.entry_from_leaf(li[:10] + struct.pack(">H", t) + li[12:], ed)[0] == ora.FE_UNKNOWN_TYPE
    bad = bytearray(li)
    bad[15 + 4] ^= 0xFF                                                                # TBS header of the leaf certificate
    assert ora.entry_from_leaf(bytes(bad), ed)[0] == ora.FE_BAD_CERT


def test_extra_data_framing_errors(ora, certs):
    leaf, (ca, root) = certs[0], certs[1][:2]
    li = fe.merkle_tree_leaf_x509(TS, leaf)
    ed = fe.certificate_chain([ca, root])
    assert ora.entry_from_leaf(li, ed + b"\x00")[0] == ora.FE_BAD_EXTRA               # trailing data after CertificateChain
    assert ora.entry_from_leaf(li, ed[:-1])[0] == ora.FE_BAD_EXTRA
    assert ora.entry_from_leaf(li, b"")[0] == ora.FE_BAD_EXTRA
    assert ora.entry_from_leaf(li, b"\x00\x00")[0] == ora.FE_BAD_EXTRA
    assert ora.entry_from_leaf(li, b"\x00\x00\x03\x00\x00\x00")[0] == ora.FE_BAD_EXTRA  # zero-length ASN.1Cert in the chain
    inner = fe.certificate_chain([ca])[3:]
    assert ora.entry_from_leaf(li, struct.pack(">I", len(inner) + 2)[1:] + inner + b"\x00\x00")[0] == ora.FE_BAD_EXTRA  # ragged tail
    pli = fe.merkle_tree_leaf_precert(TS, bytes(32), fe.tbs_of(leaf))
    ped = fe.precert_chain_entry(leaf, [ca])
    assert ora.entry_from_leaf(pli, ped)[0] == ora.FE_OK
    assert ora.entry_from_leaf(pli, ped + b"\x00")[0] == ora.FE_BAD_EXTRA
    assert ora.entry_from_leaf(pli, ped[:len(leaf) + 3])[0] == ora.FE_BAD_EXTRA        # chain vector missing
    assert ora.entry_from_leaf(pli, b"\x00\x00\x00" + fe.certificate_chain([ca]))[0] == ora.FE_BAD_EXTRA  # empty pre_certificate
    assert ora.entry_from_leaf(pli, ed)[0] == ora.FE_BAD_EXTRA                          # x509-style extra_data under a precert leaf


def test_find_entry_spans_roundtrip(certs):
    leaf, (ca, root) = certs[0], certs[1][:2]
    entries = [(fe.merkle_tree_leaf_x509(TS + i, leaf), fe.certificate_chain([ca, root][:i % 3])) for i in range(7)]
    body = fe.get_entries_body(entries)
    lo, ll, xo, xl = fe.find_entry_spans(body, base=100)
    assert lo.size == 7 and lo.dtype == np.uint64 and ll.dtype == np.uint32
    for i, (li, ed) in enumerate(entries):
        assert base64.b64decode(body[int(lo[i]) - 100:int(lo[i]) - 100 + int(ll[i])]) == li
        assert base64.b64decode(body[int(xo[i]) - 100:int(xo[i]) - 100 + int(xl[i])]) == ed
    spaced = body.replace(b'":"', b'" : "')
    assert fe.find_entry_spans(spaced)[1].tolist() == ll.tolist()
    with pytest.raises(ValueError):
        fe.find_entry_spans(body.replace(b"/", b"\\/"))


def test_raw_process_composes_the_sequential_semantics(ora, certs, golden):
    """Duplicates across x509 / precert entries of the same certificate dedup to one (same issuer, serial, expDate)."""
    from conftest import NOW_NS
    leaf, (ca, root) = certs[0], certs[1][:2]
    tbs = fe.tbs_of(leaf)
    entries = [
        (fe.merkle_tree_leaf_x509(TS, leaf), fe.certificate_chain([ca, root])),
        (fe.merkle_tree_leaf_precert(TS + 1, bytes(32), tbs), fe.precert_chain_entry(leaf, [ca, root])),
        (fe.merkle_tree_leaf_x509(TS + 2, leaf), fe.certificate_chain([])),
        (fe.merkle_tree_leaf_x509(TS + 3, leaf) + b"\x00", fe.certificate_chain([ca])),
    ]
    body = fe.get_entries_body(entries)
    lo, ll, xo, xl = fe.find_entry_spans(body)
    db = ora.DB(log_expired=True)
    r = ora.raw_process(db, body, lo, ll, xo, xl, NOW_NS)
    assert r.entry_status.tolist() == [0, 0, 0, ora.FE_BAD_LEAF]
    assert r.entry_type.tolist() == [0, 1, 0, 0]
    assert r.path.status.tolist() == [0, 0, 5, 1]           # OK, OK, NO_ISSUER, PARSE_ERR (dropped)
    assert r.path.was_unknown.tolist() == [1, 0, 0, 0]
    assert r.timestamp_ms.tolist() == [TS, TS + 1, TS + 2, TS + 3]


# ---- the framing rules once more, independently (plain Python over RFC 6962), against generated structures ----

def _py_entry_from_leaf(li: bytes, ed: bytes):
    """RFC 6962 §3.4 / §4.6 read straight off the RFC with struct -- no code shared with the C oracle.
    Returns (status, entry_type, leaf bytes or None, chain list or None); the certificate parses are left out
    (status is what LogEntryFromLeaf's tls.Unmarshal calls decide)."""
    def opaque24(buf, at, lo=1):
        if at + 3 > len(buf):
            raise ValueError
        n = int.from_bytes(buf[at:at + 3], "big")
        if n < lo or at + 3 + n > len(buf):
            raise ValueError
        return buf[at + 3:at + 3 + n], at + 3 + n

    def chain_at(buf, at):
        body, end = opaque24(buf, at, lo=0)
        if end != len(buf):
            raise ValueError
        out, q = [], 0
        while q < len(body):
            c, q = opaque24(body, q)
            out.append(c)
        return out

    if len(li) < 12 or li[1] != 0:
        return 2, 0xFF, None, None
    etype = struct.unpack(">H", li[10:12])[0]
    if etype not in (0, 1):
        return 3, 0xFF, None, None
    try:
        at = 12 + (32 if etype == 1 else 0)
        body, at = opaque24(li, at)
        if at + 2 > len(li) or at + 2 + struct.unpack(">H", li[at:at + 2])[0] != len(li):
            raise ValueError
    except ValueError:
        return 2, etype, None, None
    try:
        if etype == 0:
            return 0, 0, body, chain_at(ed, 0)
        pre, at = opaque24(ed, 0)
        return 0, 1, pre, chain_at(ed, at)
    except ValueError:
        return 4, etype, None, None


_blob = st.binary(min_size=1, max_size=40)


@settings(max_examples=400, deadline=None)
@given(st.booleans(), st.integers(0, 2**64 - 1), _blob, st.lists(_blob, max_size=4), st.binary(max_size=6),
       st.sampled_from(["ok", "trunc_li", "trunc_ed", "extra_li", "extra_ed", "flip_li", "flip_ed", "type", "leaftype"]),
       st.integers(0, 10**6))
def test_framing_agrees_with_an_independent_reading_of_rfc6962(ora, precert, ts, cert, chain, ext, how, where):
    li = fe.merkle_tree_leaf_precert(ts, bytes(32), cert, ext) if precert else fe.merkle_tree_leaf_x509(ts, cert, ext)
    ed = fe.precert_chain_entry(cert, chain) if precert else fe.certificate_chain(chain)
    li, ed = bytearray(li), bytearray(ed)
    if how == "trunc_li": del li[where % len(li):]
    elif how == "trunc_ed": del ed[where % len(ed):]
    elif how == "extra_li": li += b"\x00" * (1 + where % 3)
    elif how == "extra_ed": ed += b"\x00" * (1 + where % 3)
    elif how == "flip_li": li[where % len(li)] ^= 1 << (where % 8)
    elif how == "flip_ed": ed[where % len(ed)] ^= 1 << (where % 8)
    elif how == "type": li[10:12] = struct.pack(">H", where % 65536)
    elif how == "leaftype": li[1] = where % 256
    want_st, want_type, want_leaf, want_chain = _py_entry_from_leaf(bytes(li), bytes(ed))
    got_st, e = ora.entry_from_leaf(bytes(li), bytes(ed))
    if want_st != 0:
        assert got_st == want_st
        return
    # framing is fine: what is left is the certificate parse of random bytes, which fails (BAD_CERT) unless DER by luck
    assert got_st in (ora.FE_OK, ora.FE_BAD_CERT)
    assert e.entry_type == want_type and e.timestamp_ms == struct.unpack(">Q", bytes(li[2:10]))[0]
    src = ed if e.leaf_src else li
    assert bytes(src[e.leaf_off:e.leaf_off + e.leaf_len]) == want_leaf
    assert e.chain_count == len(want_chain)
    if want_chain:
        assert bytes(ed[e.chain0_off:e.chain0_off + e.chain0_len]) == want_chain[0]


def test_host_page_generator_matches_the_python_encoders(ora):
    """libctmr's ctmr_synth_raw_pages_host (what tools/bench_frontend.py feeds ctmr_process_raw with) against the same
    pages built from the oracle's corpus with the Python encoders of ct_mapreduce_b200/frontend.py: byte for byte."""
    import ctypes as C
    import hashlib
    from ct_mapreduce_b200 import build, capi
    build.build()
    n, page = 2300, 700
    cfg = capi.synth_cfg(n, dup_mode=1, len_mode=1, len_lo=400, len_hi=3000)
    L = capi.load()
    need = L.ctmr_synth_raw_pages_host(C.byref(cfg), 0, n, page, None, 0, None, None, None, None)
    text = np.zeros(need, np.uint8)
    lo, ll, xo, xl = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    assert L.ctmr_synth_raw_pages_host(C.byref(cfg), 0, n, page, capi.ptr(text), need, capi.ptr(lo), capi.ptr(ll), capi.ptr(xo), capi.ptr(xl)) == need
    ocfg = ora.synth_cfg(n, dup_mode=1, len_mode=1, len_lo=400, len_hi=3000)
    blob, offs, idx = ora.synth_corpus(ocfg, 0, n)
    iblob, ioffs = ora.synth_issuers(ocfg)
    issuers = [iblob[int(ioffs[k]):int(ioffs[k + 1])].tobytes() for k in range(ioffs.size - 1)]
    entries = []
    for i in range(n):
        leaf, k = blob[int(offs[i]):int(offs[i + 1])].tobytes(), int(idx[i])
        chain = [issuers[k], issuers[(k + 1) % len(issuers)]][:1 + i % 2]
        if i % 3 == 0:
            entries.append((fe.merkle_tree_leaf_precert(1_690_000_000_000 + i, hashlib.sha256(issuers[k]).digest(), fe.tbs_of(leaf)),
                            fe.precert_chain_entry(leaf, chain)))
        else:
            entries.append((fe.merkle_tree_leaf_x509(1_690_000_000_000 + i, leaf), fe.certificate_chain(chain)))
    bodies = [fe.get_entries_body(entries[a:a + page]) for a in range(0, n, page)]
    assert text.tobytes() == b"".join(bodies)
    spans, base = [], 0
    for bd in bodies:
        spans.append(fe.find_entry_spans(bd, base))
        base += len(bd)
    for got, j in ((lo, 0), (ll, 1), (xo, 2), (xl, 3)):
        assert np.array_equal(got, np.concatenate([s[j] for s in spans]))
    # and the oracle accepts every generated entry
    for i in (0, 1, 2, 3, n - 1):
        li = ora.b64_decode(text[int(lo[i]):int(lo[i]) + int(ll[i])].tobytes())
        ed = ora.b64_decode(text[int(xo[i]):int(xo[i]) + int(xl[i])].tobytes())
        assert ora.entry_from_leaf(li, ed)[0] == ora.FE_OK
