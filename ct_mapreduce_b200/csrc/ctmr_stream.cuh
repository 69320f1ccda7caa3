// ctmr_stream.cuh -- resumable DER walker for the streaming map kernel.
//
// The v1 kernel walked the TLV tree with byte loads from global memory; ncu showed that with
// ~210 KB of shared memory carved out per SM the remaining L1 (~45 KB) thrashes under 384 lanes
// each touching its own lines (25 % miss, 21 % of warp time in long-scoreboard stalls).  This walker
// instead consumes the certificate from the SAME shared-memory chunks the SHA-256 loop streams
// through, so every certificate byte crosses L2->SM exactly once, by an asynchronous copy.
//
// It is a flat state machine (one switch in a loop) rather than recursive descent: a lane can
// suspend between any two TLV headers when it runs out of staged bytes and resume when the next
// chunk has landed.  Every state reads at most kNeed bytes from `pos`, and a step only runs when
// those bytes are staged (or the record is complete), so reads never leave the window.
//
// Fields extracted = what the reference worker reads from ct-go's x509.Certificate (SURVEY.md
// §8(a) a3); acceptance rules = Go encoding/asn1 as restated in oracle/ctmr_oracle.c -- the two
// implementations are deliberately structured differently (nested loops there, states here).
#pragma once
#include "ctmr_device.cuh"
#include "ctmr_kernels.cuh"

namespace ctmr {

enum WalkState : uint32_t {
    W_CERT = 0, W_TBS, W_VER, W_SERIAL, W_SIGALG, W_NAME, W_RDN, W_ATV, W_ATV_OID, W_ATV_VAL, W_VALIDITY, W_TIME1,
    W_TIME2, W_SPKI, W_SPKI_ALG, W_SPKI_BITS, W_OPT1, W_OPT2, W_OPT3, W_EXTS, W_EXT, W_EXT_OID, W_EXT_CRIT, W_EXT_VAL,
    W_BC_SEQ, W_BC_BOOL, W_BC_INT, W_SIGALG2, W_SIG, W_DONE, W_ERR
};

enum : uint32_t {
    WF_HAS_CN = 1u, WF_BC_VALID = 2u, WF_IS_CA = 4u, WF_IN_SUBJECT = 8u, WF_IS_CN_OID = 16u, WF_CN_MATCH = 32u,
    WF_BC_CA_TMP = 64u
};

struct Walker {
    uint32_t pos, st, flags;
    uint32_t end_tbs, end_b, end_c, end_d;  // tbs | name or extensions | set, validity, spki or extension | atv or bc value
    uint32_t which;                          // last arc of a 2.5.29.x extension OID, 0 otherwise
    uint32_t serial_off, serial_len;
    uint32_t name_off, name_len;    // issuer Name, full TLV
    uint32_t crldp_off, crldp_len;  // cRLDistributionPoints extnValue content
    int64_t not_after;

    __device__ __forceinline__ void init() {
        pos = 0; st = W_CERT; flags = 0; end_tbs = end_b = end_c = end_d = 0; which = 0;
        serial_off = serial_len = 0; not_after = 0;
        name_off = name_len = crldp_off = crldp_len = 0;
    }
};

constexpr uint32_t kWalkNeed = 48;  // most bytes any single state reads from `pos` (W_SERIAL: 6 + 39)

// Byte x of the record, read from the lane's shared-memory window.  `base` is the 32-bit
// shared-space address that record offset 0 maps to (it may lie below the slot: only staged
// offsets are ever dereferenced).
struct SmemWindow {
    uint32_t base;
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + x));
        return v;
    }
};

// TLV header at `pos` inside [pos, lim): Go parseTagAndLength rules (single-octet tag, definite
// minimal length, value inside the container).  Deliberately NOT inlined: the walker has ~30 call
// sites and the map kernel's instruction footprint is what the SHA loop competes with for I-cache.
static __device__ __noinline__ uint64_t w_hdr_packed(SmemWindow rd, uint32_t pos, uint32_t lim) {
    constexpr uint64_t kFail = ~0ull;
    if (pos + 2u > lim) return kFail;
    // the four bytes at `pos` in one go: two aligned 32-bit shared loads + a funnel shift
    // (tag | len0 << 8 | len1 << 16 | len2 << 24); the window has slack past the staged bytes
    const uint32_t a = rd.base + pos, wa = a & ~3u;
    uint32_t lo, hi;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(wa));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(wa + 4u));
    const uint32_t v = __funnelshift_r(lo, hi, (a & 3u) * 8u);
    const uint32_t tag = v & 0xffu, l = (v >> 8) & 0xffu;
    if ((tag & 0x1fu) == 0x1fu) return kFail;
    uint32_t hdr, len;
    if (l < 0x80u) {
        hdr = 2;
        len = l;
    } else {
        const uint32_t nb = l & 0x7fu;
        if (nb == 0u || nb > 4u || pos + 2u + nb > lim) return kFail;
        if (nb == 1u) {
            len = (v >> 16) & 0xffu;
            if (len < 0x80u) return kFail;  // non-minimal (covers the leading-zero rule too)
        } else if (nb == 2u) {
            len = ((v >> 8) & 0xff00u) | (v >> 24);
            if (len < 0x100u) return kFail;  // superfluous leading zero
        } else {  // 3 or 4 length octets: certificates of 64 KiB and more, byte by byte
            len = 0;
            for (uint32_t i = 0; i < nb; ++i) {
                if (len >= (1u << 23)) return kFail;
                len = (len << 8) | rd(pos + 2u + i);
                if (len == 0u) return kFail;
            }
            if (len < 0x80u) return kFail;
        }
        hdr = 2u + nb;
    }
    if (len > lim - pos - hdr) return kFail;
    return ((uint64_t)len << 32) | (hdr << 8) | tag;  // registers only: no out-parameter, no stack
}

__device__ __forceinline__ bool w_hdr(SmemWindow rd, uint32_t pos, uint32_t lim, Tlv& t) {
    const uint64_t r = w_hdr_packed(rd, pos, lim);
    t.tag = (uint32_t)r & 0xffu;
    t.hdr = ((uint32_t)r >> 8) & 0xffu;
    t.len = (uint32_t)(r >> 32);
    return r != ~0ull;
}

template <class R>
__device__ __forceinline__ bool w_2d(const R& rd, uint32_t p, uint32_t& v) {
    const uint32_t a = rd(p) - (uint32_t)'0', b = rd(p + 1) - (uint32_t)'0';
    v = a * 10u + b;
    return a <= 9u && b <= 9u;
}

// UTCTime / GeneralizedTime content at cert offset p (same acceptance rules as der_time)
constexpr int64_t kBadTime = INT64_MIN;

template <class R>
__device__ __noinline__ int64_t w_time_raw(const R rd, uint32_t tag, uint32_t p, uint32_t len) {
    uint32_t yy, cc, mo, dd, hh, mi, ss = 0, pos;
    int64_t year;
    if (tag == 0x18u) {
        if (len < 15u || !w_2d(rd, p, cc) || !w_2d(rd, p + 2, yy)) return kBadTime;
        year = (int64_t)cc * 100 + yy;
        pos = 4;
    } else if (tag == 0x17u) {
        if (len < 11u || !w_2d(rd, p, yy)) return kBadTime;
        year = yy >= 50u ? 1900 + (int64_t)yy : 2000 + (int64_t)yy;
        pos = 2;
    } else {
        return kBadTime;
    }
    if (len > 19u) return kBadTime;  // longest legal form is YYYYMMDDhhmmss+hhmm
    if (!w_2d(rd, p + pos, mo) || !w_2d(rd, p + pos + 2, dd) || !w_2d(rd, p + pos + 4, hh) || !w_2d(rd, p + pos + 6, mi))
        return kBadTime;
    pos += 8;
    bool has_sec = false;
    if (pos + 2u <= len) {
        const uint32_t c0 = rd(p + pos);
        if (c0 >= '0' && c0 <= '9') {
            if (!w_2d(rd, p + pos, ss)) return kBadTime;
            pos += 2;
            has_sec = true;
        }
    }
    if (tag == 0x18u && !has_sec) return kBadTime;
    if (pos >= len) return kBadTime;
    int64_t off = 0;
    const uint32_t z = rd(p + pos);
    if (z == 'Z') {
        if (pos + 1u != len) return kBadTime;
    } else if (z == '+' || z == '-') {
        uint32_t oh, om;
        if (pos + 5u != len || !w_2d(rd, p + pos + 1, oh) || !w_2d(rd, p + pos + 3, om)) return kBadTime;
        if (oh > 23u || om > 59u) return kBadTime;
        off = (int64_t)oh * 3600 + (int64_t)om * 60;
        if (off == 0) return kBadTime;
        if (z == '-') off = -off;
    } else {
        return kBadTime;
    }
    if (mo < 1u || mo > 12u || dd < 1u || hh > 23u || mi > 59u || ss > 59u) return kBadTime;
    uint32_t maxd = (mo == 2u) ? 28u : ((0xAD5u >> (mo - 1u)) & 1u ? 31u : 30u);
    if (mo == 2u && (year % 4 == 0) && (year % 100 != 0 || year % 400 == 0)) maxd = 29u;
    if (dd > maxd) return kBadTime;
    return days_from_civil(year, mo, dd) * 86400 + (int64_t)hh * 3600 + (int64_t)mi * 60 + (int64_t)ss - off;
}

template <class R>
__device__ __forceinline__ bool w_time(const R rd, uint32_t tag, uint32_t p, uint32_t len, int64_t& out) {
    const int64_t v = w_time_raw(rd, tag, p, len);
    if (v == kBadTime) return false;
    out = v;
    return true;
}

// Runs the walker as far as the staged bytes allow.
//   rd(x)      byte x of the record from the shared-memory window (valid for x < avail)
//   far(x)     byte x from global memory (only for a CommonName running past the window)
//   avail      record bytes staged so far; avail == L means the whole record has been seen
//   key_words  where the raw serial goes when it is met (10 words: len | serial[39], zero padded)
template <class R, class F>
__device__ inline void walk_advance(Walker& w, const R& rd, const F& far, uint32_t avail, uint32_t L,
                                    const FilterCfg& flt, uint32_t* __restrict__ key_words) {
    const bool final = avail >= L;
    while (w.st < W_DONE) {
        if (!final && w.pos + kWalkNeed > avail) return;  // suspend until the next chunk lands
        Tlv t;
        switch (w.st) {
        case W_CERT:
            if (!w_hdr(rd, 0, L, t) || t.tag != 0x30u || t.hdr + t.len != L) { w.st = W_ERR; break; }
            w.pos = t.hdr;
            w.st = W_TBS;
            break;
        case W_TBS:
            if (!w_hdr(rd, w.pos, L, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_tbs = w.pos + t.len;
            w.st = W_VER;
            break;
        case W_VER:  // [0] EXPLICIT version, optional
            if (!w_hdr(rd, w.pos, w.end_tbs, t)) { w.st = W_ERR; break; }
            if (t.tag == 0xa0u) w.pos += t.hdr + t.len;
            w.st = W_SERIAL;
            break;
        case W_SERIAL: {  // raw content octets, leading zeros kept (storage/types.go:171-178)
            if (!w_hdr(rd, w.pos, w.end_tbs, t) || t.tag != 0x02u || t.len == 0u) { w.st = W_ERR; break; }
            const uint32_t c0 = w.pos + t.hdr;
            if (t.len > 1u) {
                const uint32_t b0 = rd(c0), b1 = rd(c0 + 1);
                if ((b0 == 0x00u && (b1 & 0x80u) == 0u) || (b0 == 0xffu && (b1 & 0x80u) != 0u)) { w.st = W_ERR; break; }
            }
            w.serial_off = c0;
            w.serial_len = t.len;
            if (key_words != nullptr) {  // {len, serial[39]} zero padded = words 4..13 of the key record
                const uint32_t n = t.len < CTMR_MAX_SERIAL ? t.len : CTMR_MAX_SERIAL;
#pragma unroll 1
                for (uint32_t wi = 0; wi < 10u; ++wi) {
                    uint32_t v = 0;
#pragma unroll
                    for (uint32_t k = 0; k < 4u; ++k) {
                        const uint32_t i = 4u * wi + k;  // byte i of {len, serial...}
                        const uint32_t b = i == 0u ? n : (i - 1u < n ? rd(c0 + i - 1u) : 0u);
                        v |= b << (8u * k);
                    }
                    key_words[wi] = v;
                }
            }
            w.pos = c0 + t.len;
            w.st = W_SIGALG;
            break;
        }
        case W_SIGALG:
            if (!w_hdr(rd, w.pos, w.end_tbs, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr + t.len;
            w.st = W_NAME;
            break;
        case W_NAME:  // issuer, then (WF_IN_SUBJECT) subject
            if (!w_hdr(rd, w.pos, w.end_tbs, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            if (!(w.flags & WF_IN_SUBJECT)) { w.name_off = w.pos; w.name_len = t.hdr + t.len; }
            w.pos += t.hdr;
            w.end_b = w.pos + t.len;
            w.st = w.pos < w.end_b ? W_RDN : ((w.flags & WF_IN_SUBJECT) ? W_SPKI : W_VALIDITY);
            break;
        case W_RDN:
            if (!w_hdr(rd, w.pos, w.end_b, t) || t.tag != 0x31u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_c = w.pos + t.len;
            if (w.pos < w.end_c) w.st = W_ATV;
            else w.st = w.pos < w.end_b ? W_RDN : ((w.flags & WF_IN_SUBJECT) ? W_SPKI : W_VALIDITY);
            break;
        case W_ATV:
            if (!w_hdr(rd, w.pos, w.end_c, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_d = w.pos + t.len;
            w.st = W_ATV_OID;
            break;
        case W_ATV_OID: {
            if (!w_hdr(rd, w.pos, w.end_d, t) || t.tag != 0x06u) { w.st = W_ERR; break; }
            const uint32_t o = w.pos + t.hdr;
            w.flags &= ~WF_IS_CN_OID;
            if (!(w.flags & WF_IN_SUBJECT) && t.len == 3u && rd(o) == 0x55u && rd(o + 1) == 0x04u && rd(o + 2) == 0x03u)
                w.flags |= WF_IS_CN_OID;
            w.pos = o + t.len;
            w.st = W_ATV_VAL;
            break;
        }
        case W_ATV_VAL: {
            if (!w_hdr(rd, w.pos, w.end_d, t)) { w.st = W_ERR; break; }
            if ((w.flags & WF_IS_CN_OID) &&
                (t.tag == 0x0cu || t.tag == 0x13u || t.tag == 0x16u || t.tag == 0x14u || t.tag == 0x12u)) {
                // Issuer.CommonName = this value (the last one wins); evaluate the issuerCNFilter
                // prefixes right now, while the bytes are in shared memory (ct-fetch.go:57-63)
                const uint32_t c0 = w.pos + t.hdr, cnl = t.len;
                bool match = false;
                for (uint32_t q = 0; q < flt.n_prefix && !match; ++q) {
                    const uint32_t po = flt.off[q], pl = flt.off[q + 1] - po;
                    if (pl > cnl) continue;
                    bool eq = true;
                    for (uint32_t i = 0; i < pl; ++i) {
                        const uint32_t x = c0 + i;
                        const uint32_t b = x < avail ? rd(x) : far(x);
                        if (b != flt.bytes[po + i]) { eq = false; break; }
                    }
                    match = eq;
                }
                w.flags |= WF_HAS_CN;
                w.flags = match ? (w.flags | WF_CN_MATCH) : (w.flags & ~WF_CN_MATCH);
            }
            w.pos = w.end_d;
            if (w.pos < w.end_c) w.st = W_ATV;
            else if (w.pos < w.end_b) w.st = W_RDN;
            else w.st = (w.flags & WF_IN_SUBJECT) ? W_SPKI : W_VALIDITY;
            break;
        }
        case W_VALIDITY:
            if (!w_hdr(rd, w.pos, w.end_tbs, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_c = w.pos + t.len;
            w.st = W_TIME1;
            break;
        case W_TIME1: {
            int64_t nb;
            if (!w_hdr(rd, w.pos, w.end_c, t) || !w_time(rd, t.tag, w.pos + t.hdr, t.len, nb)) { w.st = W_ERR; break; }
            w.pos += t.hdr + t.len;
            w.st = W_TIME2;
            break;
        }
        case W_TIME2:
            if (!w_hdr(rd, w.pos, w.end_c, t) || !w_time(rd, t.tag, w.pos + t.hdr, t.len, w.not_after)) { w.st = W_ERR; break; }
            w.pos = w.end_c;
            w.flags |= WF_IN_SUBJECT;
            w.st = W_NAME;
            break;
        case W_SPKI:
            if (!w_hdr(rd, w.pos, w.end_tbs, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_c = w.pos + t.len;
            w.st = W_SPKI_ALG;
            break;
        case W_SPKI_ALG:
            if (!w_hdr(rd, w.pos, w.end_c, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr + t.len;
            w.st = W_SPKI_BITS;
            break;
        case W_SPKI_BITS:
            if (!w_hdr(rd, w.pos, w.end_c, t) || t.tag != 0x03u || t.len == 0u) { w.st = W_ERR; break; }
            w.pos = w.end_c;
            w.st = W_OPT1;
            break;
        case W_OPT1:  // [1] issuerUniqueID
            if (w.pos >= w.end_tbs) { w.pos = w.end_tbs; w.st = W_SIGALG2; break; }
            if (!w_hdr(rd, w.pos, w.end_tbs, t)) { w.st = W_ERR; break; }
            if (t.tag == 0x81u || t.tag == 0xa1u) w.pos += t.hdr + t.len;
            w.st = W_OPT2;
            break;
        case W_OPT2:  // [2] subjectUniqueID
            if (w.pos >= w.end_tbs) { w.pos = w.end_tbs; w.st = W_SIGALG2; break; }
            if (!w_hdr(rd, w.pos, w.end_tbs, t)) { w.st = W_ERR; break; }
            if (t.tag == 0x82u || t.tag == 0xa2u) w.pos += t.hdr + t.len;
            w.st = W_OPT3;
            break;
        case W_OPT3:  // [3] EXPLICIT extensions; anything else is tolerated trailing data
            if (w.pos >= w.end_tbs) { w.pos = w.end_tbs; w.st = W_SIGALG2; break; }
            if (!w_hdr(rd, w.pos, w.end_tbs, t)) { w.st = W_ERR; break; }
            if (t.tag == 0xa3u) {
                w.pos += t.hdr;
                w.end_b = w.pos + t.len;
                w.st = W_EXTS;
            } else {
                w.pos = w.end_tbs;
                w.st = W_SIGALG2;
            }
            break;
        case W_EXTS:
            if (!w_hdr(rd, w.pos, w.end_b, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_b = w.pos + t.len;
            if (w.pos < w.end_b) w.st = W_EXT;
            else { w.pos = w.end_tbs; w.st = W_SIGALG2; }
            break;
        case W_EXT:
            if (!w_hdr(rd, w.pos, w.end_b, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.end_c = w.pos + t.len;
            w.st = W_EXT_OID;
            break;
        case W_EXT_OID: {
            if (!w_hdr(rd, w.pos, w.end_c, t) || t.tag != 0x06u) { w.st = W_ERR; break; }
            const uint32_t o = w.pos + t.hdr;
            w.which = (t.len == 3u && rd(o) == 0x55u && rd(o + 1) == 0x1du) ? rd(o + 2) : 0u;
            w.pos = o + t.len;
            w.st = W_EXT_CRIT;
            break;
        }
        case W_EXT_CRIT:  // critical BOOLEAN DEFAULT FALSE
            if (!w_hdr(rd, w.pos, w.end_c, t)) { w.st = W_ERR; break; }
            if (t.tag == 0x01u) {
                if (t.len != 1u) { w.st = W_ERR; break; }
                const uint32_t bv = rd(w.pos + t.hdr);
                if (bv != 0x00u && bv != 0xffu) { w.st = W_ERR; break; }
                w.pos += t.hdr + t.len;
            }
            w.st = W_EXT_VAL;
            break;
        case W_EXT_VAL:
            if (!w_hdr(rd, w.pos, w.end_c, t) || t.tag != 0x04u) { w.st = W_ERR; break; }
            if (w.which == 0x13u) {  // basicConstraints: look inside the OCTET STRING
                w.pos += t.hdr;
                w.end_d = w.pos + t.len;
                w.st = W_BC_SEQ;
            } else {
                if (w.which == 0x1fu) { w.crldp_off = w.pos + t.hdr; w.crldp_len = t.len; }  // cRLDistributionPoints
                w.pos = w.end_c;
                if (w.pos < w.end_b) w.st = W_EXT;
                else { w.pos = w.end_tbs; w.st = W_SIGALG2; }
            }
            break;
        case W_BC_SEQ:  // SEQ { cA BOOLEAN DEFAULT FALSE, pathLen INTEGER OPTIONAL }, no trailing data
            if (!w_hdr(rd, w.pos, w.end_d, t) || t.tag != 0x30u || w.pos + t.hdr + t.len != w.end_d) { w.st = W_ERR; break; }
            w.pos += t.hdr;
            w.flags &= ~WF_BC_CA_TMP;
            w.st = W_BC_BOOL;
            break;
        case W_BC_BOOL:
            if (w.pos < w.end_d) {
                if (!w_hdr(rd, w.pos, w.end_d, t)) { w.st = W_ERR; break; }
                if (t.tag == 0x01u) {
                    if (t.len != 1u) { w.st = W_ERR; break; }
                    const uint32_t bv = rd(w.pos + t.hdr);
                    if (bv != 0x00u && bv != 0xffu) { w.st = W_ERR; break; }
                    if (bv) w.flags |= WF_BC_CA_TMP;
                    w.pos += t.hdr + t.len;
                }
            }
            w.st = W_BC_INT;
            break;
        case W_BC_INT:
            if (w.pos < w.end_d) {
                if (!w_hdr(rd, w.pos, w.end_d, t) || t.tag != 0x02u || t.len == 0u) { w.st = W_ERR; break; }
            }
            w.flags |= WF_BC_VALID;  // a later basicConstraints overrides an earlier one
            w.flags = (w.flags & WF_BC_CA_TMP) ? (w.flags | WF_IS_CA) : (w.flags & ~WF_IS_CA);
            w.pos = w.end_c;
            if (w.pos < w.end_b) w.st = W_EXT;
            else { w.pos = w.end_tbs; w.st = W_SIGALG2; }
            break;
        case W_SIGALG2:
            if (!w_hdr(rd, w.pos, L, t) || t.tag != 0x30u) { w.st = W_ERR; break; }
            w.pos += t.hdr + t.len;
            w.st = W_SIG;
            break;
        case W_SIG:
            if (!w_hdr(rd, w.pos, L, t) || t.tag != 0x03u || t.len == 0u) { w.st = W_ERR; break; }
            w.st = W_DONE;
            break;
        default:
            w.st = W_ERR;
            break;
        }
    }
}

}  // namespace ctmr
