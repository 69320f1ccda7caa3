"""ctypes binding of the CPU oracle (oracle/ctmr_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (ct_mapreduce_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libctmr_oracle.so")

ST_OK, ST_PARSE_ERR, ST_FILTER_CA, ST_FILTER_EXPIRED, ST_FILTER_CN, ST_NO_ISSUER, ST_ISSUER_PARSE_ERR, ST_SERIAL_TOO_LONG = range(8)
NO_ISSUER = 0xFFFFFFFF


class SynthCfg(C.Structure):
    """Mirror of ctmr_synth_cfg (ct_mapreduce_b200/csrc/ctmr_synth.h)."""

    _fields_ = [
        ("seed", C.c_uint64),
        ("n_total", C.c_uint64),
        ("n_issuers", C.c_uint32),
        ("len_mode", C.c_uint32),
        ("len_lo", C.c_uint32),
        ("len_hi", C.c_uint32),
        ("dup_mode", C.c_uint32),
        ("reserved", C.c_uint32),
        ("now_sec", C.c_int64),
    ]


def synth_cfg(n_total, seed=20260922, n_issuers=256, len_mode=0, len_lo=1436, len_hi=1564, dup_mode=0,
              now_sec=1767225600) -> SynthCfg:
    return SynthCfg(seed, n_total, n_issuers, len_mode, len_lo, len_hi, dup_mode, 0, now_sec)


class OraCert(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "tbs_off", "tbs_len", "serial_off", "serial_len", "issuer_off", "issuer_len", "cn_off", "cn_len",
        "spki_off", "spki_len", "crldp_off", "crldp_len")] + [
        ("not_before", C.c_int64), ("not_after", C.c_int64), ("has_cn", C.c_int32), ("bc_valid", C.c_int32),
        ("is_ca", C.c_int32)]


class OraEntry(C.Structure):
    _fields_ = [("timestamp_ms", C.c_uint64), ("entry_type", C.c_uint32), ("leaf_src", C.c_uint32),
                ("leaf_off", C.c_uint32), ("leaf_len", C.c_uint32), ("chain0_off", C.c_uint32), ("chain0_len", C.c_uint32),
                ("chain_count", C.c_uint32), ("tbs_off", C.c_uint32), ("tbs_len", C.c_uint32)]


class OraOut(C.Structure):
    _fields_ = [("status", C.c_void_p), ("sha256", C.c_void_p), ("exp_hour", C.c_void_p), ("serial_off", C.c_void_p),
                ("serial_len", C.c_void_p), ("was_unknown", C.c_void_p), ("first_issuer_hour", C.c_void_p),
                ("issuer_name_off", C.c_void_p), ("issuer_name_len", C.c_void_p), ("crldp_off", C.c_void_p),
                ("crldp_len", C.c_void_p), ("first_issuer_dn", C.c_void_p), ("first_crldp", C.c_void_p)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (Makefile in this directory)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)  # incremental: a no-op when up to date
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, i64, sz = C.c_void_p, C.c_uint64, C.c_int64, C.c_size_t
        L.ora_sha256.argtypes = [vp, sz, vp]
        L.ora_b64url.argtypes = [vp, sz, C.c_char_p]; L.ora_b64url.restype = sz
        L.ora_parse_cert.argtypes = [vp, sz, C.POINTER(OraCert)]; L.ora_parse_cert.restype = C.c_int
        L.ora_issuer_id.argtypes = [vp, sz, vp, C.c_char_p]
        L.ora_exp_hour.argtypes = [i64]; L.ora_exp_hour.restype = i64
        L.ora_expdate_id.argtypes = [i64, C.c_char_p]
        L.ora_day_id.argtypes = [i64, C.c_char_p]
        L.ora_filter.argtypes = [vp, C.POINTER(OraCert), vp, sz, C.c_int, i64]; L.ora_filter.restype = C.c_int
        L.ora_cache_new.restype = vp
        L.ora_cache_free.argtypes = [vp]
        L.ora_cache_set_insert.argtypes = [vp, C.c_char_p, sz, vp, sz]; L.ora_cache_set_insert.restype = C.c_int
        L.ora_cache_set_cardinality.argtypes = [vp, C.c_char_p, sz]; L.ora_cache_set_cardinality.restype = u64
        L.ora_cache_set_list.argtypes = [vp, C.c_char_p, sz, vp, sz, C.POINTER(sz)]; L.ora_cache_set_list.restype = u64
        L.ora_serials_key.argtypes = [i64, C.c_char_p, C.c_char_p, sz]; L.ora_serials_key.restype = sz
        L.ora_was_unknown.argtypes = [vp, i64, C.c_char_p, vp, sz]; L.ora_was_unknown.restype = C.c_int
        L.ora_db_new.argtypes = [vp, sz, C.c_int]; L.ora_db_new.restype = vp
        L.ora_db_free.argtypes = [vp]
        L.ora_db_process.argtypes = [vp, vp, vp, u64, vp, vp, C.c_uint32, vp, i64, C.c_int, C.POINTER(OraOut)]
        L.ora_db_process.restype = C.c_int
        L.ora_db_issuer_counts.argtypes = [vp, vp, vp, u64]; L.ora_db_issuer_counts.restype = u64
        L.ora_db_set_cardinality.argtypes = [vp, i64, vp]; L.ora_db_set_cardinality.restype = u64
        L.ora_db_filter_counters.argtypes = [vp, vp]
        L.ora_db_evict_expired.argtypes = [vp, i64]; L.ora_db_evict_expired.restype = u64
        L.ora_map_only.argtypes = [vp, vp, u64, vp, sz, C.c_int, i64, C.c_int, vp]; L.ora_map_only.restype = u64
        L.ora_parse_tbs.argtypes = [vp, sz, C.POINTER(OraCert)]; L.ora_parse_tbs.restype = C.c_int
        L.ora_b64_decode.argtypes = [vp, sz, vp, sz]; L.ora_b64_decode.restype = C.c_long
        L.ora_entry_from_leaf.argtypes = [vp, sz, vp, sz, C.POINTER(OraEntry)]; L.ora_entry_from_leaf.restype = C.c_int
        L.ora_synth_lengths.argtypes = [C.POINTER(SynthCfg), u64, u64, vp]; L.ora_synth_lengths.restype = u64
        L.ora_synth_write.argtypes = [C.POINTER(SynthCfg), u64, u64, vp, vp]
        L.ora_synth_issuer_idx.argtypes = [C.POINTER(SynthCfg), u64, u64, vp]
        L.ora_synth_issuers.argtypes = [C.POINTER(SynthCfg), vp, vp, sz]; L.ora_synth_issuers.restype = u64
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- leaves

def sha256(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    lib().ora_sha256(buf, len(data), out)
    return bytes(out)


def b64url(data: bytes) -> str:
    out = C.create_string_buffer(4 * ((len(data) + 2) // 3) + 1)
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
    lib().ora_b64url(buf, len(data), out)
    return out.value.decode()


def parse_cert(der: bytes):
    """Returns (rc, OraCert)."""
    c = OraCert()
    buf = (C.c_uint8 * max(1, len(der))).from_buffer_copy(der or b"\0")
    rc = lib().ora_parse_cert(buf, len(der), C.byref(c))
    return rc, c


def issuer_id(spki: bytes):
    dig = (C.c_uint8 * 32)()
    sid = C.create_string_buffer(45)
    buf = (C.c_uint8 * max(1, len(spki))).from_buffer_copy(spki or b"\0")
    lib().ora_issuer_id(buf, len(spki), dig, sid)
    return bytes(dig), sid.value.decode()


def exp_hour(sec: int) -> int:
    return lib().ora_exp_hour(sec)


def expdate_id(hour: int) -> str:
    out = C.create_string_buffer(14)
    lib().ora_expdate_id(hour, out)
    return out.value.decode()


def day_id(sec: int) -> str:
    out = C.create_string_buffer(11)
    lib().ora_day_id(sec, out)
    return out.value.decode()


def filter_cert(der: bytes, flt: bytes, log_expired: bool, now_ns: int) -> int:
    rc, c = parse_cert(der)
    if rc:
        return ST_PARSE_ERR
    buf = (C.c_uint8 * len(der)).from_buffer_copy(der)
    fb = (C.c_uint8 * max(1, len(flt))).from_buffer_copy(flt or b"\0")
    return lib().ora_filter(buf, C.byref(c), fb, len(flt), int(log_expired), now_ns)


class Cache:
    """Sets with storage.MockRemoteCache semantics (mockcache.go:38-61,120-122)."""

    def __init__(self):
        self.h = lib().ora_cache_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_cache_free(self.h)
            self.h = None

    def set_insert(self, key: str, member: bytes) -> bool:
        k = key.encode()
        m = (C.c_uint8 * max(1, len(member))).from_buffer_copy(member or b"\0")
        return bool(lib().ora_cache_set_insert(self.h, k, len(k), m, len(member)))

    def set_cardinality(self, key: str) -> int:
        k = key.encode()
        return lib().ora_cache_set_cardinality(self.h, k, len(k))

    def set_list(self, key: str):
        k = key.encode()
        buf = np.zeros(1 << 20, dtype=np.uint8)
        used = C.c_size_t(0)
        n = lib().ora_cache_set_list(self.h, k, len(k), _p(buf), buf.size, C.byref(used))
        out, o = [], 0
        raw = buf.tobytes()
        for _ in range(n):
            ln = int.from_bytes(raw[o:o + 4], "little")
            out.append(raw[o + 4:o + 4 + ln])
            o += 4 + ln
        return out

    def was_unknown(self, hour: int, issuer_id_str: str, serial: bytes) -> bool:
        s = (C.c_uint8 * max(1, len(serial))).from_buffer_copy(serial or b"\0")
        return bool(lib().ora_was_unknown(self.h, hour, issuer_id_str.encode(), s, len(serial)))


def serials_key(hour: int, issuer_id_str: str) -> str:
    out = C.create_string_buffer(200)
    lib().ora_serials_key(hour, issuer_id_str.encode(), out, 200)
    return out.value.decode()


# ---------------------------------------------------------------- composed path

class Result:
    def __init__(self, n):
        self.status = np.zeros(n, np.uint8)
        self.sha256 = np.zeros((n, 32), np.uint8)
        self.exp_hour = np.zeros(n, np.int64)
        self.serial_off = np.zeros(n, np.uint32)
        self.serial_len = np.zeros(n, np.uint32)
        self.was_unknown = np.zeros(n, np.uint8)
        self.first_issuer_hour = np.zeros(n, np.uint8)
        self.issuer_name_off = np.zeros(n, np.uint32)
        self.issuer_name_len = np.zeros(n, np.uint32)
        self.crldp_off = np.zeros(n, np.uint32)
        self.crldp_len = np.zeros(n, np.uint32)
        self.first_issuer_dn = np.zeros(n, np.uint8)
        self.first_crldp = np.zeros(n, np.uint8)


class DB:
    """insertCTWorker + FilesystemDatabase.Store over Mock cache/backend, numThreads=1."""

    def __init__(self, issuer_cn_filter: bytes = b"", log_expired: bool = False):
        f = (C.c_uint8 * max(1, len(issuer_cn_filter))).from_buffer_copy(issuer_cn_filter or b"\0")
        self.h = lib().ora_db_new(f, len(issuer_cn_filter), int(log_expired))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_db_free(self.h)
            self.h = None

    def process(self, blob, offsets, issuer_blob, issuer_offsets, issuer_idx, now_ns, nthreads=1) -> Result:
        blob = np.ascontiguousarray(blob, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        issuer_blob = np.ascontiguousarray(issuer_blob, np.uint8)
        issuer_offsets = np.ascontiguousarray(issuer_offsets, np.uint64)
        issuer_idx = np.ascontiguousarray(issuer_idx, np.uint32)
        n = offsets.size - 1
        r = Result(n)
        o = OraOut(r.status.ctypes.data, r.sha256.ctypes.data, r.exp_hour.ctypes.data, r.serial_off.ctypes.data,
                   r.serial_len.ctypes.data, r.was_unknown.ctypes.data, r.first_issuer_hour.ctypes.data,
                   r.issuer_name_off.ctypes.data, r.issuer_name_len.ctypes.data, r.crldp_off.ctypes.data,
                   r.crldp_len.ctypes.data, r.first_issuer_dn.ctypes.data, r.first_crldp.ctypes.data)
        rc = lib().ora_db_process(self.h, _p(blob), _p(offsets), n, _p(issuer_blob), _p(issuer_offsets),
                                  issuer_offsets.size - 1, _p(issuer_idx), now_ns, nthreads, C.byref(o))
        assert rc == 0
        return r

    def issuer_counts(self):
        ids = np.zeros((65536, 32), np.uint8)
        cnt = np.zeros(65536, np.uint64)
        n = lib().ora_db_issuer_counts(self.h, _p(ids), _p(cnt), 65536)
        return {bytes(ids[i]): int(cnt[i]) for i in range(n)}

    def set_cardinality(self, hour: int, digest: bytes) -> int:
        d = (C.c_uint8 * 32).from_buffer_copy(digest)
        return lib().ora_db_set_cardinality(self.h, hour, d)

    def evict_expired(self, now_sec: int) -> int:
        """Redis EXPIREAT firing on every serials:: set whose expDate <= now (knowncertificates.go:98-104)."""
        return int(lib().ora_db_evict_expired(self.h, now_sec))

    def filter_counters(self):
        out = np.zeros(8, np.uint64)
        lib().ora_db_filter_counters(self.h, _p(out))
        return out


def map_only(blob, offsets, flt: bytes, log_expired: bool, now_ns: int, nthreads: int, want_sha=True):
    blob = np.ascontiguousarray(blob, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = offsets.size - 1
    sha = np.zeros((n, 32), np.uint8) if want_sha else None
    f = (C.c_uint8 * max(1, len(flt))).from_buffer_copy(flt or b"\0")
    kept = lib().ora_map_only(_p(blob), _p(offsets), n, f, len(flt), int(log_expired), now_ns, nthreads,
                              _p(sha) if want_sha else None)
    return kept, sha


# ---------------------------------------------------------------- synthetic corpus

def synth_corpus(cfg: SynthCfg, first: int, n: int):
    """Returns (blob u8, offsets u64[n+1], issuer_idx u32[n]) for entries [first, first+n)."""
    offsets = np.zeros(n + 1, np.uint64)
    total = lib().ora_synth_lengths(C.byref(cfg), first, n, _p(offsets))
    blob = np.zeros(total + 64, np.uint8)[:total]
    lib().ora_synth_write(C.byref(cfg), first, n, _p(offsets), _p(blob))
    idx = np.zeros(n, np.uint32)
    lib().ora_synth_issuer_idx(C.byref(cfg), first, n, _p(idx))
    return blob, offsets, idx


def synth_issuers(cfg: SynthCfg):
    offsets = np.zeros(cfg.n_issuers + 1, np.uint64)
    total = lib().ora_synth_issuers(C.byref(cfg), _p(offsets), None, 0)
    blob = np.zeros(total, np.uint8)
    lib().ora_synth_issuers(C.byref(cfg), _p(offsets), _p(blob), total)
    return blob, offsets


# ---------------------------------------------------------------- CT wire-format front end

FE_OK, FE_BAD_BASE64, FE_BAD_LEAF, FE_UNKNOWN_TYPE, FE_BAD_EXTRA, FE_BAD_CERT = range(6)


def b64_decode(text: bytes):
    """base64.StdEncoding as encoding/json applies it to []byte fields; None = corrupt input."""
    src = (C.c_uint8 * max(1, len(text))).from_buffer_copy(text or b"\0")
    out = (C.c_uint8 * (len(text) // 4 * 3 + 3))()
    n = lib().ora_b64_decode(src, len(text), out, len(out))
    return None if n < 0 else bytes(out[:n])


def parse_tbs(tbs: bytes):
    c = OraCert()
    buf = (C.c_uint8 * max(1, len(tbs))).from_buffer_copy(tbs or b"\0")
    return lib().ora_parse_tbs(buf, len(tbs), C.byref(c)), c


def entry_from_leaf(leaf_input: bytes, extra_data: bytes):
    """ct.LogEntryFromLeaf on decoded bytes -> (ORA_FE_*, OraEntry)."""
    e = OraEntry()
    a = (C.c_uint8 * max(1, len(leaf_input))).from_buffer_copy(leaf_input or b"\0")
    b = (C.c_uint8 * max(1, len(extra_data))).from_buffer_copy(extra_data or b"\0")
    return lib().ora_entry_from_leaf(a, len(leaf_input), b, len(extra_data), C.byref(e)), e


class RawResult:
    pass


def raw_process(db: "DB", text: bytes, leaf_off, leaf_len, extra_off, extra_len, now_ns: int) -> RawResult:
    """get-entries strings -> what the reference's downloader + worker + Store do with them, entry by entry
    (cmd/ct-fetch/ct-fetch.go:446-484 then :191-245), through the sequential oracle DB."""
    n = len(leaf_off)
    r = RawResult()
    r.entry_status = np.zeros(n, np.uint8)
    r.entry_type = np.full(n, 0xFF, np.uint8)
    r.timestamp_ms = np.zeros(n, np.uint64)
    r.leaf_src = np.zeros(n, np.uint8)
    r.leaf_off = np.zeros(n, np.uint32)
    r.leaf_len = np.zeros(n, np.uint32)
    r.issuer_der = [None] * n          # Chain[0] bytes of entries that reached the worker with a chain
    leaves, issuers, issuer_pos, issuer_idx = [], [], {}, np.full(n, 0xFFFFFFFF, np.uint32)
    for i in range(n):
        li = b64_decode(text[int(leaf_off[i]):int(leaf_off[i]) + int(leaf_len[i])])
        ed = b64_decode(text[int(extra_off[i]):int(extra_off[i]) + int(extra_len[i])])
        leaf = b""
        if li is None or ed is None:
            r.entry_status[i] = FE_BAD_BASE64
        else:
            st, e = entry_from_leaf(li, ed)
            r.entry_status[i] = st
            r.entry_type[i] = e.entry_type
            r.timestamp_ms[i] = e.timestamp_ms
            if st == FE_OK or (st == FE_BAD_CERT and e.entry_type == 0):
                # an x509 leaf that fails to parse is dropped by the downloader; the GPU path reports it as
                # status PARSE_ERR from the same bytes, so the oracle hands the bytes over as well
                src = ed if e.leaf_src else li
                leaf = src[e.leaf_off:e.leaf_off + e.leaf_len]
                r.leaf_src[i], r.leaf_off[i], r.leaf_len[i] = e.leaf_src, e.leaf_off, e.leaf_len
                if e.chain0_len and st == FE_OK:
                    ch = ed[e.chain0_off:e.chain0_off + e.chain0_len]
                    r.issuer_der[i] = ch
                    k = issuer_pos.get(ch)
                    if k is None:
                        k = issuer_pos[ch] = len(issuers)
                        issuers.append(ch)
                    issuer_idx[i] = k
        leaves.append(leaf)
    offsets = np.zeros(n + 1, np.uint64)
    offsets[1:] = np.cumsum([len(x) for x in leaves], dtype=np.uint64)
    blob = np.frombuffer(b"".join(leaves) or b"\0", np.uint8)
    ioffs = np.zeros(len(issuers) + 1, np.uint64)
    ioffs[1:] = np.cumsum([len(x) for x in issuers], dtype=np.uint64)
    iblob = np.frombuffer(b"".join(issuers) or b"\0", np.uint8)
    r.path = db.process(blob, offsets, iblob, ioffs, issuer_idx, now_ns)
    r.leaves = leaves  # the DER each entry's certificate was taken from (b"" for dropped entries)
    return r
