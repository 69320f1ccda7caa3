"""The oracle against the reference's own known-answer tests (SURVEY.md §8(c) "golden vectors").

Each test cites the Go test it transcribes.  These pin the LEAVES of the path; the composed
map path has no test in the reference (see oracle/ctmr_oracle.h header: "parity unpinned" for
CommonName / NotAfter / IsCA against Go, cross-checked against `cryptography` instead).
"""
import calendar
import hashlib

import numpy as np

from conftest import NOW_NS, pack


def test_issuer_lazy_init_kat(ora):
    # storage/types_test.go:41-57: Issuer{spki: [0xFF]}.ID()
    digest, sid = ora.issuer_id(b"\xff")
    assert sid == "qBAK5qoZQNC2Y7sxzUZhQuu9vVGHExuS2TgYmHgy64k="
    assert digest == hashlib.sha256(b"\xff").digest()


def test_sha256_nist_vectors(ora):
    # FIPS 180-4 examples; crypto/sha256.Sum256 at storage/types.go:156
    assert ora.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert ora.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    m = b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq"
    assert ora.sha256(m).hex() == "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"
    for n in (55, 56, 57, 63, 64, 65, 119, 120, 121, 1000):
        msg = bytes((i * 7 + 3) & 0xFF for i in range(n))
        assert ora.sha256(msg) == hashlib.sha256(msg).digest()


def test_serial_from_cert_with_leading_zeroes(ora, golden):
    # storage/types_test.go:81-101: serial string "00aa", ID "AKo="
    der = golden["kLeadingZeroes"]["der"]
    rc, c = ora.parse_cert(der)
    assert rc == 0
    serial = der[c.serial_off:c.serial_off + c.serial_len]
    assert serial.hex() == "00aa"
    assert ora.b64url(serial) == "AKo="


def test_b64url_matches_go_url_encoding(ora):
    # storage/types_test.go:172-201 (TestLog): CertificateLogIDFromShortURL = base64.URLEncoding
    assert ora.b64url(b"yeti2021.ct.digicert.com/log/") == "eWV0aTIwMjEuY3QuZGlnaWNlcnQuY29tL2xvZy8="
    # storage/types_test.go:254-269: serial ID "AESq_w==" uses the URL alphabet with padding
    assert ora.b64url(bytes.fromhex("0044aaff")) == "AESq_w=="


def test_exp_date_from_time(ora):
    # storage/types_test.go:239-252 + knowncertificates_test.go:85-110: 2004-01-20 04:22:19 -> "2004-01-20-04"
    t = calendar.timegm((2004, 1, 20, 4, 22, 19))
    h = ora.exp_hour(t)
    assert ora.expdate_id(h) == "2004-01-20-04"
    assert h * 3600 == calendar.timegm((2004, 1, 20, 4, 0, 0))  # ExpireAt time, knowncertificates_test.go:106
    assert ora.day_id(t) == "2004-01-20"
    # hour boundaries (types_test.go:203-237): 04:59:59 still hour 4, 05:00:00 is hour 5
    h4 = ora.exp_hour(calendar.timegm((2004, 1, 19, 4, 0, 0)))
    assert ora.exp_hour(calendar.timegm((2004, 1, 19, 4, 59, 59))) == h4
    assert ora.exp_hour(calendar.timegm((2004, 1, 19, 5, 0, 0))) == h4 + 1
    assert ora.expdate_id(ora.exp_hour(calendar.timegm((2004, 1, 19, 23, 59, 59)))) == "2004-01-19-23"
    # Truncate rounds toward -inf for pre-1970 instants
    assert ora.exp_hour(-1) == -1 and ora.expdate_id(-1) == "1969-12-31-23"


def test_unknown_sequence_and_sorted_set(ora):
    # storage/knowncertificates_test.go:11-55 (Test_Unknown)
    cache = ora.Cache()
    key = "serials::2029-01-30::test issuer"
    for s in (b"\x01", b"\x02", b"\x03", b"\x04"):
        cache.set_insert(key, s)  # backend.Data preloaded with 01..04
    for s in (b"\x01", b"\x02", b"\x03", b"\x04"):
        assert cache.set_insert(key, s) is False  # known
    assert cache.set_insert(key, b"\x05") is True
    assert cache.set_insert(key, b"\x05") is False
    # the Go test's final JSON is the sorted list of the five one-byte strings (line 52)
    assert cache.set_list(key) == [b"\x01", b"\x02", b"\x03", b"\x04", b"\x05"]
    # storage/knowncertificates_test.go:57-83 (Count)
    assert cache.set_cardinality(key) == 5


def test_serials_key_format(ora):
    # storage/knowncertificates_test.go:102: "serials::2004-01-20-04::test issuer"
    h = ora.exp_hour(calendar.timegm((2004, 1, 20, 4, 22, 19)))
    assert ora.serials_key(h, "test issuer") == "serials::2004-01-20-04::test issuer"
    cache = ora.Cache()
    assert cache.was_unknown(h, "test issuer", b"\x05") is True
    assert cache.was_unknown(h, "test issuer", b"\x05") is False


def test_binary_safe_members(ora):
    # storage/types.go:218-220: BinaryString may hold 0x00 and non-UTF-8 octets
    cache = ora.Cache()
    assert cache.set_insert("k", b"\x00") and cache.set_insert("k", b"\x00\x00") and cache.set_insert("k", b"\xff\x00a")
    assert not cache.set_insert("k", b"\x00\x00")
    assert cache.set_list("k") == [b"\x00", b"\x00\x00", b"\xff\x00a"]


def test_fixture_fields(ora, golden):
    # the three embedded PEMs (types_test.go:21-39, filesystemdatabase_test.go:17-64); expectations come from
    # tests/golden/fixtures.json = hashlib + `cryptography` (tools/extract_reference_fixtures.py)
    for name, fx in golden.items():
        der = fx["der"]
        assert len(der) == fx["der_len"]
        assert ora.sha256(der).hex() == fx["sha256_der"]
        rc, c = ora.parse_cert(der)
        assert rc == 0, name
        assert der[c.serial_off:c.serial_off + c.serial_len].hex() == fx["raw_serial_hex"]
        assert c.not_after == fx["not_after_unix"] and c.not_before == fx["not_before_unix"]
        assert der[c.cn_off:c.cn_off + c.cn_len].decode() == fx["issuer_cn"]
        assert bool(c.bc_valid) == fx["bc_valid"] and bool(c.is_ca) == fx["is_ca"]
        assert c.spki_len == fx["spki_len"]
        _, sid = ora.issuer_id(der[c.spki_off:c.spki_off + c.spki_len])
        assert sid == fx["issuer_id_of_own_spki"]
        if fx["crl_dps"]:
            assert fx["crl_dps"][0].encode() in der[c.crldp_off:c.crldp_off + c.crldp_len]


def test_filter_order_and_untrimmed_split(ora, golden):
    # cmd/ct-fetch/ct-fetch.go:44-70
    real, empty, lz = golden["kRealSPKI"]["der"], golden["kEmptySPKI"]["der"], golden["kLeadingZeroes"]["der"]
    before = (golden["kLeadingZeroes"]["not_after_unix"] - 10) * 10**9
    assert ora.filter_cert(real, b"", False, before) == ora.ST_FILTER_CA       # CA wins over everything
    assert ora.filter_cert(empty, b"", True, NOW_NS) == ora.ST_FILTER_CA
    assert ora.filter_cert(lz, b"", False, NOW_NS) == ora.ST_FILTER_EXPIRED     # expired in 2020
    assert ora.filter_cert(lz, b"", True, NOW_NS) == ora.ST_OK                  # logExpiredEntries
    assert ora.filter_cert(lz, b"", False, before) == ora.ST_OK                 # empty filter keeps all
    assert ora.filter_cert(lz, b"ca", False, before) == ora.ST_OK
    assert ora.filter_cert(lz, b"x,c", False, before) == ora.ST_OK              # HasPrefix("ca", "c")
    assert ora.filter_cert(lz, b"x, ca", False, before) == ora.ST_FILTER_CN     # " ca" keeps its space
    assert ora.filter_cert(lz, b"cab", False, before) == ora.ST_FILTER_CN
    assert ora.filter_cert(lz, b"x,", False, before) == ora.ST_OK               # empty element matches all
    # NotAfter.Before(now) is strict
    na = golden["kLeadingZeroes"]["not_after_unix"]
    assert ora.filter_cert(lz, b"", False, na * 10**9) == ora.ST_OK
    assert ora.filter_cert(lz, b"", False, na * 10**9 + 1) == ora.ST_FILTER_EXPIRED


def test_store_composition_on_fixtures(ora, golden):
    # insertCTWorker + Store (ct-fetch.go:191-245, filesystemdatabase.go:158-211) over the fixtures:
    # kLeadingZeroes issued by "ca" (kEmptySPKI carries that subject and key)
    lz, ca = golden["kLeadingZeroes"]["der"], golden["kEmptySPKI"]["der"]
    blob, offs = pack([lz, lz, ca, lz])
    iblob, ioffs = pack([ca])
    idx = np.array([0, 0, 0, ora.NO_ISSUER], np.uint32)
    before = (golden["kLeadingZeroes"]["not_after_unix"] - 10) * 10**9
    db = ora.DB(b"", False)
    r = db.process(blob, offs, iblob, ioffs, idx, before)
    assert list(r.status) == [ora.ST_OK, ora.ST_OK, ora.ST_FILTER_CA, ora.ST_NO_ISSUER]
    assert list(r.was_unknown) == [1, 0, 0, 0]
    assert list(r.first_issuer_hour) == [1, 0, 0, 0]
    rc, c = ora.parse_cert(ca)
    dig = hashlib.sha256(ca[c.spki_off:c.spki_off + c.spki_len]).digest()
    assert db.issuer_counts() == {dig: 1}
    assert db.set_cardinality(int(r.exp_hour[0]), dig) == 1
    assert r.sha256[0].tobytes() == hashlib.sha256(lz).digest()
    # a second batch remembers the first (the Redis set persists)
    r2 = db.process(blob[:len(lz)], offs[:2], iblob, ioffs, idx[:1], before)
    assert list(r2.was_unknown) == [0]
