"""Second opinion for the parity-unpinned fields: the oracle's DER walker against Python
`cryptography` (an independent X.509 implementation) on the synthetic corpus, and the generator's
output against a strict parser (so the corpus is valid RFC 5280 DER that the reference would accept).
"""
import hashlib
import warnings

import numpy as np
import pytest
from cryptography import x509
from cryptography.hazmat.primitives import serialization
from cryptography.x509.oid import ExtensionOID, NameOID

from conftest import NOW_NS, NOW_SEC, README_FILTER

warnings.filterwarnings("ignore")


def _check_corpus(ora, cfg, n):
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    seen_gen, seen_utf8, seen_ca, seen_nobc, seen_lz, seen_ec, seen_ldap = 0, 0, 0, 0, 0, 0, 0
    for i in range(n):
        der = blob[offs[i]:offs[i + 1]].tobytes()
        rc, c = ora.parse_cert(der)
        assert rc == 0, (i, rc)
        cert = x509.load_der_x509_certificate(der)
        cert.public_key()  # RSA / P-256 keys must be well formed (the Go parser checks the EC point)
        serial = der[c.serial_off:c.serial_off + c.serial_len]
        assert int.from_bytes(serial, "big") == cert.serial_number
        seen_lz += serial[0] == 0
        assert int(cert.not_valid_after_utc.timestamp()) == c.not_after
        assert int(cert.not_valid_before_utc.timestamp()) == c.not_before
        cns = cert.issuer.get_attributes_for_oid(NameOID.COMMON_NAME)
        assert der[c.cn_off:c.cn_off + c.cn_len].decode() == cns[-1].value
        assert der[c.tbs_off:c.tbs_off + c.tbs_len] == cert.tbs_certificate_bytes
        seen_gen += der[c.tbs_off:].find(b"\x18\x0f20") > 0
        seen_utf8 += der[c.cn_off - 2] == 0x0C
        try:
            bc = cert.extensions.get_extension_for_oid(ExtensionOID.BASIC_CONSTRAINTS).value
            assert c.bc_valid == 1 and c.is_ca == int(bc.ca)
            seen_ca += bc.ca
        except x509.ExtensionNotFound:
            assert c.bc_valid == 0
            seen_nobc += 1
        dps = cert.extensions.get_extension_for_oid(ExtensionOID.CRL_DISTRIBUTION_POINTS).value
        uri = dps[0].full_name[0].value
        assert uri.encode() in der[c.crldp_off:c.crldp_off + c.crldp_len]
        seen_ldap += uri.startswith("ldap://")
        seen_ec += c.spki_len == 91
        assert ora.sha256(der) == hashlib.sha256(der).digest()
    return dict(gen=seen_gen, utf8=seen_utf8, ca=seen_ca, nobc=seen_nobc, lz=seen_lz, ec=seen_ec, ldap=seen_ldap)


def test_uniform_corpus_matches_cryptography(ora):
    cfg = ora.synth_cfg(100000)
    seen = _check_corpus(ora, cfg, 2500)
    # the corpus must exercise every branch the survey lists (§8(d))
    assert seen["gen"] > 0 and seen["utf8"] > 0 and seen["ca"] > 0 and seen["nobc"] > 0 and seen["lz"] > 0
    blob, offs, _ = ora.synth_corpus(cfg, 0, 2500)
    lens = np.diff(offs.astype(np.int64))
    assert lens.min() >= 1436 and lens.max() <= 1564 and abs(lens.mean() - 1500) < 5


def test_mixed_corpus_matches_cryptography(ora):
    cfg = ora.synth_cfg(100000, len_mode=1, len_lo=512, len_hi=8192, dup_mode=1)
    seen = _check_corpus(ora, cfg, 2500)
    assert seen["ec"] > 0 and seen["ldap"] >= 0
    blob, offs, _ = ora.synth_corpus(cfg, 0, 2500)
    lens = np.diff(offs.astype(np.int64))
    assert lens.min() >= 512 and lens.max() < 8192


def test_issuer_certificates(ora):
    cfg = ora.synth_cfg(1000)
    iblob, ioffs = ora.synth_issuers(cfg)
    ids = set()
    for k in range(cfg.n_issuers):
        der = iblob[ioffs[k]:ioffs[k + 1]].tobytes()
        rc, c = ora.parse_cert(der)
        assert rc == 0 and c.bc_valid and c.is_ca
        cert = x509.load_der_x509_certificate(der)
        spki = cert.public_key().public_bytes(serialization.Encoding.DER, serialization.PublicFormat.SubjectPublicKeyInfo)
        assert spki == der[c.spki_off:c.spki_off + c.spki_len]
        digest, _ = ora.issuer_id(spki)
        assert digest == hashlib.sha256(spki).digest()
        ids.add(digest)
    assert len(ids) == cfg.n_issuers


def test_leaf_issuer_name_equals_ca_subject(ora):
    cfg = ora.synth_cfg(1000)
    blob, offs, idx = ora.synth_corpus(cfg, 0, 200)
    iblob, ioffs = ora.synth_issuers(cfg)
    for i in range(200):
        leaf = x509.load_der_x509_certificate(blob[offs[i]:offs[i + 1]].tobytes())
        ca = x509.load_der_x509_certificate(iblob[ioffs[idx[i]]:ioffs[idx[i] + 1]].tobytes())
        assert leaf.issuer == ca.subject


def test_duplicates_appear_exactly_twice(ora):
    n = 4000
    cfg = ora.synth_cfg(n, dup_mode=1)
    blob, offs, _ = ora.synth_corpus(cfg, 0, n)
    digests = {}
    for i in range(n):
        digests.setdefault(hashlib.sha256(blob[offs[i]:offs[i + 1]].tobytes()).digest(), []).append(i)
    assert len(digests) == n // 2 and all(len(v) == 2 for v in digests.values())
    gaps = [v[1] - v[0] for v in digests.values()]
    assert min(gaps) < n // 8 and max(gaps) > n // 2  # duplicates straddle any chunking


def test_composed_path_statistics(ora):
    # the oracle's composed path over the corpus: expected proportions of SURVEY.md §8(d)
    n = 20000
    cfg = ora.synth_cfg(n)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    db = ora.DB(README_FILTER, False)
    r = db.process(blob, offs, iblob, ioffs, idx, NOW_NS, nthreads=4)
    frac = np.bincount(r.status, minlength=8) / n
    assert 0.01 < frac[ora.ST_FILTER_CA] < 0.03
    assert 0.05 < frac[ora.ST_FILTER_EXPIRED] < 0.09
    assert 0.40 < frac[ora.ST_FILTER_CN] < 0.50   # classes 2 and 3 of 4 fail the README filter
    assert frac[ora.ST_PARSE_ERR] == 0
    ok = r.status == ora.ST_OK
    assert r.was_unknown[ok].all() and not r.was_unknown[~ok].any()
    assert sum(db.issuer_counts().values()) == int(ok.sum())
    # single-threaded and multi-threaded map halves agree
    r1 = ora.DB(README_FILTER, False).process(blob, offs, iblob, ioffs, idx, NOW_NS, nthreads=1)
    for f in ("status", "sha256", "exp_hour", "serial_off", "serial_len", "was_unknown", "first_issuer_hour"):
        assert np.array_equal(getattr(r, f), getattr(r1, f)), f
