/*
 * ctmr_synth.h -- deterministic synthetic CT corpus, usable from C (gcc), C++ and CUDA device code.
 *
 * This is bench/test tooling, not part of the hot path: it manufactures the inputs that
 * BASELINE.json's configs name (SURVEY.md §8(d) "Synthetic inputs"): valid RFC 5280 v3 DER
 * leaf certificates of a target size, issued by one of `n_issuers` synthetic CAs, plus the CA
 * certificates themselves (the `Chain[0]` the reference worker parses at
 * cmd/ct-fetch/ct-fetch.go:221).  Entry i of a corpus depends only on (cfg, i), so the device
 * generator (one thread per certificate) and the CPU oracle regenerate identical bytes.
 *
 * Shape of a leaf (what the reference's parser, ct-go x509, will see):
 *   Certificate ::= SEQ { TBS SEQ { [0] v3, INTEGER serial, sigalg, issuer Name, validity,
 *                                   subject Name, SPKI, [3] extensions }, sigalg, BIT STRING sig }
 *   serial   16 PRNG octets, top bit clear (1 %: 17 octets 00|1xxxxxxx.. -> exercises the raw
 *            leading-zero rule of storage/types.go:171-178)
 *   issuer   C=US, O=Synth CA <k>, CN=<class prefix> ... <k>; k%4 selects the CN class:
 *            0 "Let's Encrypt Authority X<k>", 1 " ISRG Root X<k>" (leading space, matches the
 *            README's untrimmed filter " ISRG"), 2 "ISRG Root X<k>" (near miss), 3 other.
 *            10 % of issuers use UTF8String instead of PrintableString.
 *   validity notAfter uniform in [now-30d, now+397d] at 1 s resolution; 5 % GeneralizedTime
 *   SPKI     RSA-2048 (294 B) or, for small targets, P-256 (91 B, point taken from a table of
 *            real curve points so strict parsers accept it)
 *   exts     basicConstraints (88 % CA:FALSE critical, 2 % CA:TRUE, 10 % absent), keyUsage,
 *            cRLDistributionPoints (1 % ldap://), private-arc padding extension sized so the
 *            whole certificate hits the target length
 *   sig      256 PRNG octets (RSA issuer) or a 73-octet ECDSA-Sig-Value (EC issuer)
 */
#ifndef CTMR_SYNTH_H
#define CTMR_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define CTMR_HD __host__ __device__ static inline
#else
#define CTMR_HD static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ctmr_synth_cfg {
    uint64_t seed;      /* corpus seed (SURVEY: 20260922) */
    uint64_t n_total;   /* corpus size; domain of the duplicate permutation */
    uint32_t n_issuers; /* distinct issuing CAs (256) */
    uint32_t len_mode;  /* 0: uniform in [len_lo, len_hi]; 1: octave log-uniform in [len_lo, len_hi) */
    uint32_t len_lo;
    uint32_t len_hi;
    uint32_t dup_mode;  /* 0: all certificates distinct; 1: every certificate appears exactly twice;
                         * d >= 2: one entry in d repeats another certificate (pairs at permuted positions) */
    uint32_t reserved;
    int64_t now_sec;    /* fixed "now" (2026-01-01T00:00:00Z = 1767225600) */
} ctmr_synth_cfg;

#define CTMR_SYNTH_NOW_DEFAULT 1767225600LL
#define CTMR_SYNTH_SEED_DEFAULT 20260922ULL

typedef struct ctmr_synth_plan {
    uint64_t cert_id;
    uint32_t issuer;        /* k */
    uint32_t target_len;
    int64_t not_before;
    int64_t not_after;
    uint8_t leaf_ec;        /* leaf key: 1 = P-256, 0 = RSA-2048 */
    uint8_t issuer_ec;      /* issuer key type -> signature algorithm */
    uint8_t serial_len;     /* 16 or 17 */
    uint8_t na_generalized; /* notAfter encoded as GeneralizedTime */
    uint8_t bc_mode;        /* 0 absent, 1 CA:FALSE, 2 CA:TRUE */
    uint8_t crl_ldap;
    uint8_t issuer_utf8;
    uint8_t pad0;
    uint32_t sz_issuer_name; /* full TLV sizes */
    uint32_t sz_validity;
    uint32_t sz_subject;
    uint32_t sz_spki;
    uint32_t sz_bc;
    uint32_t sz_crl;
    uint32_t crl_uri_len;
    uint32_t pad_len;        /* content octets of the padding extension's OCTET STRING; 0 = no ext */
    uint32_t sz_pad_ext;
    uint32_t exts_content;
    uint32_t tbs_content;
    uint32_t sz_sigalg;
    uint32_t sz_sig;
    uint32_t cert_content;
    uint32_t total;
} ctmr_synth_plan;

/* ---------------------------------------------------------------- PRNG */

CTMR_HD uint64_t ctmr_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* counter-based: value k of stream (domain, id) */
CTMR_HD uint64_t ctmr_synth_rand(const ctmr_synth_cfg* c, uint32_t domain, uint64_t id, uint32_t k) {
    uint64_t z = c->seed + 0x9E3779B97F4A7C15ULL * (id + 1);
    z = ctmr_mix64(z ^ ((uint64_t)domain << 56));
    return ctmr_mix64(z + 0xD1B54A32D192ED03ULL * (uint64_t)(k + 1));
}

/* Feistel permutation of [0, n) with cycle walking (duplicate placement) */
CTMR_HD uint64_t ctmr_synth_perm(const ctmr_synth_cfg* c, uint64_t i) {
    uint64_t n = c->n_total;
    uint32_t bits = 2;
    while (bits < 64 && ((uint64_t)1 << bits) < n) bits += 2;
    uint32_t half = bits / 2;
    uint64_t mask = ((uint64_t)1 << half) - 1;
    uint64_t x = i;
    do {
        uint64_t l = x >> half, r = x & mask;
        for (uint32_t round = 0; round < 4; ++round) {
            uint64_t f = ctmr_mix64(r + c->seed * 0x2545F4914F6CDD1DULL + ((uint64_t)(round + 1) << 58)) & mask;
            uint64_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}

CTMR_HD uint64_t ctmr_synth_cert_id(const ctmr_synth_cfg* c, uint64_t i) {
    if (c->dup_mode == 1 && c->n_total >= 2) return ctmr_synth_perm(c, i) >> 1;
    if (c->dup_mode >= 2 && c->n_total >= 2) {  /* positions p = d-1 (mod d) repeat the certificate of position p-1 */
        const uint64_t p = ctmr_synth_perm(c, i);
        return (p % c->dup_mode == c->dup_mode - 1) ? p - 1 : p;
    }
    return i;
}

/* ---------------------------------------------------------------- DER helpers */

CTMR_HD uint32_t ctmr_der_lsz(uint32_t n) { return n < 128u ? 1u : (n < 256u ? 2u : (n < 65536u ? 3u : 4u)); }
CTMR_HD uint32_t ctmr_der_tlv(uint32_t n) { return 1u + ctmr_der_lsz(n) + n; }

CTMR_HD uint8_t* ctmr_der_hdr(uint8_t* p, uint8_t tag, uint32_t n) {
    *p++ = tag;
    if (n < 128u) {
        *p++ = (uint8_t)n;
    } else if (n < 256u) {
        *p++ = 0x81; *p++ = (uint8_t)n;
    } else if (n < 65536u) {
        *p++ = 0x82; *p++ = (uint8_t)(n >> 8); *p++ = (uint8_t)n;
    } else {
        *p++ = 0x83; *p++ = (uint8_t)(n >> 16); *p++ = (uint8_t)(n >> 8); *p++ = (uint8_t)n;
    }
    return p;
}

CTMR_HD uint8_t* ctmr_emit(uint8_t* p, const uint8_t* s, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) p[i] = s[i];
    return p + n;
}

CTMR_HD uint32_t ctmr_fmt_u32(char* dst, uint32_t v) {
    /* digits are written straight into dst (no scratch array: nvcc 12.9 overlapped a local scratch
     * buffer with the caller's string buffer here, caught by tests/test_gpu_parity.py) */
    uint32_t nd = 1;
    for (uint32_t t = v; t >= 10u; t /= 10u) ++nd;
    for (uint32_t i = nd; i-- > 0u;) {
        dst[i] = (char)('0' + v % 10u);
        v /= 10u;
    }
    return nd;
}

CTMR_HD uint32_t ctmr_cat(char* dst, uint32_t at, const char* s) {
    while (*s) dst[at++] = *s++;
    return at;
}

/* civil date from days since 1970-01-01 (proleptic Gregorian) */
CTMR_HD void ctmr_civil_from_days(int64_t z, int32_t* y, uint32_t* m, uint32_t* d) {
    z += 719468;
    int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    uint32_t doe = (uint32_t)(z - era * 146097);
    uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int64_t yy = (int64_t)yoe + era * 400;
    uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    uint32_t mp = (5 * doy + 2) / 153;
    *d = doy - (153 * mp + 2) / 5 + 1;
    *m = mp < 10 ? mp + 3 : mp - 9;
    *y = (int32_t)(yy + (*m <= 2));
}

CTMR_HD uint8_t* ctmr_emit_2d(uint8_t* p, uint32_t v) {
    *p++ = (uint8_t)('0' + (v / 10u) % 10u);
    *p++ = (uint8_t)('0' + v % 10u);
    return p;
}

/* UTCTime (13 content octets) or GeneralizedTime (15), always the Z form */
CTMR_HD uint8_t* ctmr_emit_time(uint8_t* p, int64_t sec, int generalized) {
    int64_t days = sec >= 0 ? sec / 86400 : -((-sec + 86399) / 86400);
    uint32_t sod = (uint32_t)(sec - days * 86400);
    int32_t y; uint32_t m, d;
    ctmr_civil_from_days(days, &y, &m, &d);
    if (generalized) {
        *p++ = 0x18; *p++ = 15;
        p = ctmr_emit_2d(p, (uint32_t)y / 100u);
    } else {
        *p++ = 0x17; *p++ = 13;
    }
    p = ctmr_emit_2d(p, (uint32_t)y % 100u);
    p = ctmr_emit_2d(p, m);
    p = ctmr_emit_2d(p, d);
    p = ctmr_emit_2d(p, sod / 3600u);
    p = ctmr_emit_2d(p, (sod / 60u) % 60u);
    p = ctmr_emit_2d(p, sod % 60u);
    *p++ = 'Z';
    return p;
}

/* SET { SEQ { OID 2.5.4.<attr>, <strtag> value } } */
CTMR_HD uint32_t ctmr_name_attr_size(uint32_t len) { return ctmr_der_tlv(ctmr_der_tlv(5u + ctmr_der_tlv(len))); }

CTMR_HD uint8_t* ctmr_emit_name_attr(uint8_t* p, uint8_t attr, uint8_t strtag, const char* s, uint32_t len) {
    uint32_t seq = 5u + ctmr_der_tlv(len);
    p = ctmr_der_hdr(p, 0x31, ctmr_der_tlv(seq));
    p = ctmr_der_hdr(p, 0x30, seq);
    *p++ = 0x06; *p++ = 0x03; *p++ = 0x55; *p++ = 0x04; *p++ = attr;
    p = ctmr_der_hdr(p, strtag, len);
    for (uint32_t i = 0; i < len; ++i) p[i] = (uint8_t)s[i];
    return p + len;
}

/* ---------------------------------------------------------------- issuer identity */

CTMR_HD int ctmr_synth_issuer_is_ec(const ctmr_synth_cfg* c, uint32_t k) { (void)c; return (k & 7u) == 0u; }
CTMR_HD int ctmr_synth_issuer_is_utf8(const ctmr_synth_cfg* c, uint32_t k) {
    return (ctmr_synth_rand(c, 2, k, 0) % 10u) == 0u;
}

CTMR_HD uint32_t ctmr_synth_issuer_org(uint32_t k, char* buf) {
    uint32_t n = ctmr_cat(buf, 0, "Synth CA ");
    return n + ctmr_fmt_u32(buf + n, k);
}

CTMR_HD uint32_t ctmr_synth_issuer_cn(uint32_t k, char* buf) {
    uint32_t n;
    switch (k & 3u) {
    case 0: n = ctmr_cat(buf, 0, "Let's Encrypt Authority X"); break;
    case 1: n = ctmr_cat(buf, 0, " ISRG Root X"); break;
    case 2: n = ctmr_cat(buf, 0, "ISRG Root X"); break;
    default: n = ctmr_cat(buf, 0, "Synth Trust Services CA "); break;
    }
    return n + ctmr_fmt_u32(buf + n, k);
}

CTMR_HD uint32_t ctmr_synth_issuer_name_size(const ctmr_synth_cfg* c, uint32_t k) {
    char b[48];
    (void)c;
    uint32_t lo = ctmr_synth_issuer_org(k, b);
    uint32_t lc = ctmr_synth_issuer_cn(k, b);
    return ctmr_der_tlv(ctmr_name_attr_size(2) + ctmr_name_attr_size(lo) + ctmr_name_attr_size(lc));
}

CTMR_HD uint8_t* ctmr_emit_issuer_name(uint8_t* p, const ctmr_synth_cfg* c, uint32_t k) {
    char bo[48], bc[48];
    uint32_t lo = ctmr_synth_issuer_org(k, bo);
    uint32_t lc = ctmr_synth_issuer_cn(k, bc);
    uint8_t st = ctmr_synth_issuer_is_utf8(c, k) ? 0x0c : 0x13;
    p = ctmr_der_hdr(p, 0x30, ctmr_name_attr_size(2) + ctmr_name_attr_size(lo) + ctmr_name_attr_size(lc));
    p = ctmr_emit_name_attr(p, 0x06, 0x13, "US", 2);
    p = ctmr_emit_name_attr(p, 0x0a, st, bo, lo);
    p = ctmr_emit_name_attr(p, 0x03, st, bc, lc);
    return p;
}

/* ---------------------------------------------------------------- keys and signatures */

CTMR_HD void ctmr_synth_ec_point(uint32_t idx, uint8_t* out64) {
    const uint8_t tbl[64][64] = {
#include "ctmr_synth_ecpoints.inc"
    };
    for (uint32_t i = 0; i < 64; ++i) out64[i] = tbl[idx & 63u][i];
}

/* SPKI: 294 octets (RSA-2048, e=65537) or 91 octets (P-256) */
CTMR_HD uint8_t* ctmr_emit_spki(uint8_t* p, const ctmr_synth_cfg* c, uint32_t domain, uint64_t id, int ec) {
    if (ec) {
        const uint8_t h[27] = {0x30, 0x59, 0x30, 0x13, 0x06, 0x07, 0x2a, 0x86, 0x48, 0xce, 0x3d, 0x02, 0x01, 0x06,
                               0x08, 0x2a, 0x86, 0x48, 0xce, 0x3d, 0x03, 0x01, 0x07, 0x03, 0x42, 0x00, 0x04};
        p = ctmr_emit(p, h, 27);
        /* issuers (domain 2) take distinct table rows so that every CA has its own Issuer.ID */
        ctmr_synth_ec_point(domain == 2 ? (uint32_t)(id >> 3) : (uint32_t)ctmr_synth_rand(c, domain, id, 63), p);
        return p + 64;
    }
    const uint8_t h[33] = {0x30, 0x82, 0x01, 0x22, 0x30, 0x0d, 0x06, 0x09, 0x2a, 0x86, 0x48, 0x86, 0xf7, 0x0d, 0x01, 0x01, 0x01,
                           0x05, 0x00, 0x03, 0x82, 0x01, 0x0f, 0x00, 0x30, 0x82, 0x01, 0x0a, 0x02, 0x82, 0x01, 0x01, 0x00};
    p = ctmr_emit(p, h, 33);
    for (uint32_t w = 0; w < 32; ++w) {
        uint64_t r = ctmr_synth_rand(c, domain, id, 64 + w);
        for (uint32_t b = 0; b < 8; ++b) p[w * 8 + b] = (uint8_t)(r >> (8 * b));
    }
    p[0] |= 0x80;   /* 2048-bit modulus */
    p[255] |= 0x01; /* odd */
    p += 256;
    *p++ = 0x02; *p++ = 0x03; *p++ = 0x01; *p++ = 0x00; *p++ = 0x01;
    return p;
}

CTMR_HD uint8_t* ctmr_emit_sigalg(uint8_t* p, int ec) {
    if (ec) {
        const uint8_t a[12] = {0x30, 0x0a, 0x06, 0x08, 0x2a, 0x86, 0x48, 0xce, 0x3d, 0x04, 0x03, 0x02};
        return ctmr_emit(p, a, 12);
    }
    const uint8_t a[15] = {0x30, 0x0d, 0x06, 0x09, 0x2a, 0x86, 0x48, 0x86, 0xf7, 0x0d, 0x01, 0x01, 0x0b, 0x05, 0x00};
    return ctmr_emit(p, a, 15);
}

CTMR_HD uint8_t* ctmr_emit_sig(uint8_t* p, const ctmr_synth_cfg* c, uint32_t domain, uint64_t id, int ec) {
    if (ec) {
        const uint8_t h[5] = {0x03, 0x47, 0x00, 0x30, 0x44};
        p = ctmr_emit(p, h, 5);
        for (uint32_t half = 0; half < 2; ++half) {
            *p++ = 0x02; *p++ = 0x20;
            for (uint32_t w = 0; w < 4; ++w) {
                uint64_t r = ctmr_synth_rand(c, domain, id, 128 + half * 4 + w);
                for (uint32_t b = 0; b < 8; ++b) p[w * 8 + b] = (uint8_t)(r >> (8 * b));
            }
            p[0] = (uint8_t)((p[0] & 0x7f) | 0x01);
            p += 32;
        }
        return p;
    }
    const uint8_t h[5] = {0x03, 0x82, 0x01, 0x01, 0x00};
    p = ctmr_emit(p, h, 5);
    for (uint32_t w = 0; w < 32; ++w) {
        uint64_t r = ctmr_synth_rand(c, domain, id, 128 + w);
        for (uint32_t b = 0; b < 8; ++b) p[w * 8 + b] = (uint8_t)(r >> (8 * b));
    }
    return p + 256;
}

/* ---------------------------------------------------------------- leaf plan */

CTMR_HD uint32_t ctmr_synth_crl_uri(uint32_t k, int ldap, char* buf) {
    uint32_t n;
    if (ldap) {
        n = ctmr_cat(buf, 0, "ldap://ldap.synth");
        n += ctmr_fmt_u32(buf + n, k);
        n = ctmr_cat(buf, n, ".example/cn=");
        n += ctmr_fmt_u32(buf + n, k);
    } else {
        n = ctmr_cat(buf, 0, "http://crl.synth");
        n += ctmr_fmt_u32(buf + n, k);
        n = ctmr_cat(buf, n, ".example/");
        n += ctmr_fmt_u32(buf + n, k);
        n = ctmr_cat(buf, n, ".crl");
    }
    return n;
}

#define CTMR_SYNTH_SUBJECT_CN_LEN 30u /* "h<16 hex>.synth.example" padded form below */

CTMR_HD uint32_t ctmr_synth_subject_cn(const ctmr_synth_cfg* c, uint64_t cert_id, char* buf) {
    uint64_t r = ctmr_synth_rand(c, 1, cert_id, 7);
    uint32_t n = 0;
    buf[n++] = 'h';
    for (uint32_t i = 0; i < 16; ++i) {
        uint32_t v = (uint32_t)(r >> (4 * i)) & 15u;
        buf[n++] = (char)(v < 10 ? '0' + v : 'a' + (v - 10));
    }
    n = ctmr_cat(buf, n, ".synth.example");
    return n; /* 31 */
}

CTMR_HD uint32_t ctmr_synth_pad_ext_size(uint32_t pad_len) {
    /* SEQ { OID(12 octets TLV), OCTET STRING(pad_len) } */
    return pad_len ? ctmr_der_tlv(12u + ctmr_der_tlv(pad_len)) : 0u;
}

CTMR_HD void ctmr_synth_plan_sizes(ctmr_synth_plan* pl) {
    pl->sz_pad_ext = ctmr_synth_pad_ext_size(pl->pad_len);
    pl->exts_content = pl->sz_bc + 16u + pl->sz_crl + pl->sz_pad_ext;
    uint32_t exts = ctmr_der_tlv(ctmr_der_tlv(pl->exts_content)); /* [3] { SEQ { ... } } */
    pl->tbs_content = 5u + ctmr_der_tlv(pl->serial_len) + pl->sz_sigalg + pl->sz_issuer_name + pl->sz_validity +
                      pl->sz_subject + pl->sz_spki + exts;
    pl->cert_content = ctmr_der_tlv(pl->tbs_content) + pl->sz_sigalg + pl->sz_sig;
    pl->total = ctmr_der_tlv(pl->cert_content);
}

CTMR_HD void ctmr_synth_plan_make(const ctmr_synth_cfg* c, uint64_t i, ctmr_synth_plan* pl) {
    uint64_t id = ctmr_synth_cert_id(c, i);
    pl->cert_id = id;
    uint64_t r0 = ctmr_synth_rand(c, 1, id, 0);
    uint64_t r1 = ctmr_synth_rand(c, 1, id, 1);
    uint64_t r2 = ctmr_synth_rand(c, 1, id, 2);

    /* target length */
    uint32_t target;
    if (c->len_mode == 1) {
        uint32_t octaves = 0, lo = c->len_lo;
        while (((uint64_t)lo << (octaves + 1)) <= c->len_hi) ++octaves;
        if (octaves == 0) octaves = 1;
        uint32_t e = (uint32_t)(r0 % octaves);
        uint32_t base = lo << e;
        target = base + (uint32_t)((r0 >> 8) % base);
    } else {
        uint32_t span = c->len_hi >= c->len_lo ? c->len_hi - c->len_lo + 1u : 1u;
        target = c->len_lo + (uint32_t)((r0 >> 8) % span);
    }
    pl->target_len = target;

    /* issuer: small certificates need an EC issuer (k multiple of 8) to fit */
    uint32_t k = (uint32_t)(r1 % c->n_issuers);
    if (target < 760u && c->n_issuers >= 8u) k &= ~7u;
    pl->issuer = k;
    pl->issuer_ec = (uint8_t)ctmr_synth_issuer_is_ec(c, k);
    pl->issuer_utf8 = (uint8_t)ctmr_synth_issuer_is_utf8(c, k);
    pl->leaf_ec = (uint8_t)(target < 1100u);

    /* validity */
    int64_t span_sec = (int64_t)(30 + 397) * 86400;
    pl->not_after = c->now_sec - (int64_t)30 * 86400 + (int64_t)((r1 >> 16) % (uint64_t)span_sec);
    pl->not_before = pl->not_after - (int64_t)90 * 86400;
    pl->na_generalized = (uint8_t)(((r2 >> 0) % 100u) < 5u);

    pl->serial_len = (uint8_t)((((r2 >> 8) % 100u) == 0u) ? 17 : 16);
    uint32_t bc = (uint32_t)((r2 >> 16) % 100u);
    pl->bc_mode = (uint8_t)(bc < 2u ? 2 : (bc < 90u ? 1 : 0));
    pl->crl_ldap = (uint8_t)(((r2 >> 24) % 100u) == 0u);
    pl->pad0 = 0;

    char buf[64];
    pl->sz_issuer_name = ctmr_synth_issuer_name_size(c, k);
    pl->sz_validity = 2u + 15u + (pl->na_generalized ? 17u : 15u);
    pl->sz_subject = ctmr_der_tlv(ctmr_name_attr_size(ctmr_synth_subject_cn(c, id, buf)));
    pl->sz_spki = pl->leaf_ec ? 91u : 294u;
    pl->sz_bc = pl->bc_mode == 0 ? 0u : (pl->bc_mode == 1 ? 14u : 17u);
    pl->crl_uri_len = ctmr_synth_crl_uri(k, pl->crl_ldap, buf);
    pl->sz_crl = 19u + pl->crl_uri_len;
    pl->sz_sigalg = pl->issuer_ec ? 12u : 15u;
    pl->sz_sig = pl->issuer_ec ? 73u : 261u;

    /* padding extension sized to reach the target */
    pl->pad_len = 0;
    ctmr_synth_plan_sizes(pl);
    if (target > pl->total + 20u) {
        uint32_t pad = target - pl->total - 16u;
        for (int it = 0; it < 4; ++it) {
            pl->pad_len = pad;
            ctmr_synth_plan_sizes(pl);
            if (pl->total == target) break;
            if (pl->total > target) {
                uint32_t over = pl->total - target;
                pad = pad > over ? pad - over : 1u;
            } else {
                pad += target - pl->total;
            }
        }
        pl->pad_len = pad;
        ctmr_synth_plan_sizes(pl);
    }
}

CTMR_HD uint32_t ctmr_synth_cert_len(const ctmr_synth_cfg* c, uint64_t i) {
    ctmr_synth_plan pl;
    ctmr_synth_plan_make(c, i, &pl);
    return pl.total;
}

/* serial content octets (pl->serial_len of them) */
CTMR_HD void ctmr_synth_serial(const ctmr_synth_cfg* c, const ctmr_synth_plan* pl, uint8_t* out) {
    uint64_t a = ctmr_synth_rand(c, 1, pl->cert_id, 3);
    uint64_t b = ctmr_synth_rand(c, 1, pl->cert_id, 4);
    uint8_t raw[16];
    for (uint32_t i = 0; i < 8; ++i) { raw[i] = (uint8_t)(a >> (8 * i)); raw[8 + i] = (uint8_t)(b >> (8 * i)); }
    if (pl->serial_len == 17) {
        out[0] = 0x00;
        raw[0] |= 0x80;
        for (uint32_t i = 0; i < 16; ++i) out[1 + i] = raw[i];
    } else {
        raw[0] &= 0x7f;
        if (raw[0] == 0) raw[0] = 0x5a;
        for (uint32_t i = 0; i < 16; ++i) out[i] = raw[i];
    }
}

/* writes pl->total octets at out; returns the end pointer */
CTMR_HD uint8_t* ctmr_synth_cert_write(const ctmr_synth_cfg* c, const ctmr_synth_plan* pl, uint8_t* out) {
    uint8_t* p = out;
    char buf[64];
    p = ctmr_der_hdr(p, 0x30, pl->cert_content);
    p = ctmr_der_hdr(p, 0x30, pl->tbs_content);
    *p++ = 0xa0; *p++ = 0x03; *p++ = 0x02; *p++ = 0x01; *p++ = 0x02;
    p = ctmr_der_hdr(p, 0x02, pl->serial_len);
    ctmr_synth_serial(c, pl, p);
    p += pl->serial_len;
    p = ctmr_emit_sigalg(p, pl->issuer_ec);
    p = ctmr_emit_issuer_name(p, c, pl->issuer);
    *p++ = 0x30; *p++ = (uint8_t)(pl->sz_validity - 2u);
    p = ctmr_emit_time(p, pl->not_before, 0);
    p = ctmr_emit_time(p, pl->not_after, pl->na_generalized);
    {
        uint32_t n = ctmr_synth_subject_cn(c, pl->cert_id, buf);
        p = ctmr_der_hdr(p, 0x30, ctmr_name_attr_size(n));
        p = ctmr_emit_name_attr(p, 0x03, 0x0c, buf, n);
    }
    p = ctmr_emit_spki(p, c, 1, pl->cert_id, pl->leaf_ec);
    /* [3] extensions */
    p = ctmr_der_hdr(p, 0xa3, ctmr_der_tlv(pl->exts_content));
    p = ctmr_der_hdr(p, 0x30, pl->exts_content);
    if (pl->bc_mode == 1) {
        const uint8_t e[14] = {0x30, 0x0c, 0x06, 0x03, 0x55, 0x1d, 0x13, 0x01, 0x01, 0xff, 0x04, 0x02, 0x30, 0x00};
        p = ctmr_emit(p, e, 14);
    } else if (pl->bc_mode == 2) {
        const uint8_t e[17] = {0x30, 0x0f, 0x06, 0x03, 0x55, 0x1d, 0x13, 0x01, 0x01, 0xff, 0x04, 0x05, 0x30, 0x03, 0x01, 0x01, 0xff};
        p = ctmr_emit(p, e, 17);
    }
    {
        const uint8_t ku[16] = {0x30, 0x0e, 0x06, 0x03, 0x55, 0x1d, 0x0f, 0x01, 0x01, 0xff, 0x04, 0x04, 0x03, 0x02, 0x05, 0xa0};
        p = ctmr_emit(p, ku, 16);
    }
    {
        uint32_t u = ctmr_synth_crl_uri(pl->issuer, pl->crl_ldap, buf);
        *p++ = 0x30; *p++ = (uint8_t)(17u + u);
        *p++ = 0x06; *p++ = 0x03; *p++ = 0x55; *p++ = 0x1d; *p++ = 0x1f;
        *p++ = 0x04; *p++ = (uint8_t)(10u + u);
        *p++ = 0x30; *p++ = (uint8_t)(8u + u);
        *p++ = 0x30; *p++ = (uint8_t)(6u + u);
        *p++ = 0xa0; *p++ = (uint8_t)(4u + u);
        *p++ = 0xa0; *p++ = (uint8_t)(2u + u);
        *p++ = 0x86; *p++ = (uint8_t)u;
        for (uint32_t i = 0; i < u; ++i) p[i] = (uint8_t)buf[i];
        p += u;
    }
    if (pl->pad_len) {
        const uint8_t oid[12] = {0x06, 0x0a, 0x2b, 0x06, 0x01, 0x04, 0x01, 0x83, 0xb2, 0x03, 0x01, 0x01};
        p = ctmr_der_hdr(p, 0x30, 12u + ctmr_der_tlv(pl->pad_len));
        p = ctmr_emit(p, oid, 12);
        p = ctmr_der_hdr(p, 0x04, pl->pad_len);
        uint32_t full = pl->pad_len / 8u;
        for (uint32_t w = 0; w < full; ++w) {
            uint64_t r = ctmr_synth_rand(c, 3, pl->cert_id, w);
            for (uint32_t b = 0; b < 8; ++b) p[w * 8 + b] = (uint8_t)(r >> (8 * b));
        }
        uint64_t r = ctmr_synth_rand(c, 3, pl->cert_id, full);
        for (uint32_t b = full * 8u; b < pl->pad_len; ++b) p[b] = (uint8_t)(r >> (8 * (b & 7u)));
        p += pl->pad_len;
    }
    p = ctmr_emit_sigalg(p, pl->issuer_ec);
    p = ctmr_emit_sig(p, c, 1, pl->cert_id, pl->issuer_ec);
    return p;
}

/* ---------------------------------------------------------------- issuer (CA) certificates */

#define CTMR_SYNTH_ISSUER_TARGET 1200u

typedef struct ctmr_synth_issuer_plan {
    uint32_t k;
    uint8_t ec;
    uint8_t pad0[3];
    uint32_t sz_root_name, sz_subject, sz_spki, pad_len, sz_pad_ext, exts_content, tbs_content, cert_content, total;
} ctmr_synth_issuer_plan;

CTMR_HD void ctmr_synth_issuer_sizes(ctmr_synth_issuer_plan* ip) {
    ip->sz_pad_ext = ctmr_synth_pad_ext_size(ip->pad_len);
    ip->exts_content = 17u + 16u + ip->sz_pad_ext;
    uint32_t exts = ctmr_der_tlv(ctmr_der_tlv(ip->exts_content));
    ip->tbs_content = 5u + ctmr_der_tlv(8u) + 15u + ip->sz_root_name + 32u + ip->sz_subject + ip->sz_spki + exts;
    ip->cert_content = ctmr_der_tlv(ip->tbs_content) + 15u + 261u;
    ip->total = ctmr_der_tlv(ip->cert_content);
}

CTMR_HD void ctmr_synth_issuer_plan_make(const ctmr_synth_cfg* c, uint32_t k, ctmr_synth_issuer_plan* ip) {
    ip->k = k;
    ip->ec = (uint8_t)ctmr_synth_issuer_is_ec(c, k);
    ip->pad0[0] = ip->pad0[1] = ip->pad0[2] = 0;
    ip->sz_root_name = ctmr_der_tlv(ctmr_name_attr_size(2) + ctmr_name_attr_size(10) + ctmr_name_attr_size(13));
    ip->sz_subject = ctmr_synth_issuer_name_size(c, k);
    ip->sz_spki = ip->ec ? 91u : 294u;
    ip->pad_len = 0;
    ctmr_synth_issuer_sizes(ip);
    if (CTMR_SYNTH_ISSUER_TARGET > ip->total + 20u) {
        uint32_t pad = CTMR_SYNTH_ISSUER_TARGET - ip->total - 16u;
        for (int it = 0; it < 4; ++it) {
            ip->pad_len = pad;
            ctmr_synth_issuer_sizes(ip);
            if (ip->total == CTMR_SYNTH_ISSUER_TARGET) break;
            if (ip->total > CTMR_SYNTH_ISSUER_TARGET) pad -= ip->total - CTMR_SYNTH_ISSUER_TARGET;
            else pad += CTMR_SYNTH_ISSUER_TARGET - ip->total;
        }
        ip->pad_len = pad;
        ctmr_synth_issuer_sizes(ip);
    }
}

CTMR_HD uint32_t ctmr_synth_issuer_len(const ctmr_synth_cfg* c, uint32_t k) {
    ctmr_synth_issuer_plan ip;
    ctmr_synth_issuer_plan_make(c, k, &ip);
    return ip.total;
}

CTMR_HD uint8_t* ctmr_synth_issuer_write(const ctmr_synth_cfg* c, const ctmr_synth_issuer_plan* ip, uint8_t* out) {
    uint8_t* p = out;
    p = ctmr_der_hdr(p, 0x30, ip->cert_content);
    p = ctmr_der_hdr(p, 0x30, ip->tbs_content);
    *p++ = 0xa0; *p++ = 0x03; *p++ = 0x02; *p++ = 0x01; *p++ = 0x02;
    *p++ = 0x02; *p++ = 0x08;
    {
        uint64_t r = ctmr_synth_rand(c, 2, ip->k, 1);
        for (uint32_t b = 0; b < 8; ++b) p[b] = (uint8_t)(r >> (8 * b));
        p[0] = (uint8_t)((p[0] & 0x7f) | 0x01);
        p += 8;
    }
    p = ctmr_emit_sigalg(p, 0); /* signed by the RSA root */
    p = ctmr_der_hdr(p, 0x30, ctmr_name_attr_size(2) + ctmr_name_attr_size(10) + ctmr_name_attr_size(13));
    p = ctmr_emit_name_attr(p, 0x06, 0x13, "US", 2);
    p = ctmr_emit_name_attr(p, 0x0a, 0x13, "Synth Root", 10);
    p = ctmr_emit_name_attr(p, 0x03, 0x13, "Synth Root R1", 13);
    *p++ = 0x30; *p++ = 30;
    p = ctmr_emit_time(p, 1577836800LL, 0); /* 2020-01-01 */
    p = ctmr_emit_time(p, 2208988800LL, 0); /* 2040-01-01 */
    p = ctmr_emit_issuer_name(p, c, ip->k);
    p = ctmr_emit_spki(p, c, 2, ip->k, ip->ec);
    p = ctmr_der_hdr(p, 0xa3, ctmr_der_tlv(ip->exts_content));
    p = ctmr_der_hdr(p, 0x30, ip->exts_content);
    {
        const uint8_t e[17] = {0x30, 0x0f, 0x06, 0x03, 0x55, 0x1d, 0x13, 0x01, 0x01, 0xff, 0x04, 0x05, 0x30, 0x03, 0x01, 0x01, 0xff};
        p = ctmr_emit(p, e, 17);
        const uint8_t ku[16] = {0x30, 0x0e, 0x06, 0x03, 0x55, 0x1d, 0x0f, 0x01, 0x01, 0xff, 0x04, 0x04, 0x03, 0x02, 0x01, 0x06};
        p = ctmr_emit(p, ku, 16);
    }
    if (ip->pad_len) {
        const uint8_t oid[12] = {0x06, 0x0a, 0x2b, 0x06, 0x01, 0x04, 0x01, 0x83, 0xb2, 0x03, 0x01, 0x01};
        p = ctmr_der_hdr(p, 0x30, 12u + ctmr_der_tlv(ip->pad_len));
        p = ctmr_emit(p, oid, 12);
        p = ctmr_der_hdr(p, 0x04, ip->pad_len);
        for (uint32_t b = 0; b < ip->pad_len; ++b) {
            uint64_t r = ctmr_synth_rand(c, 4, ip->k, b / 8u);
            p[b] = (uint8_t)(r >> (8 * (b & 7u)));
        }
        p += ip->pad_len;
    }
    p = ctmr_emit_sigalg(p, 0);
    p = ctmr_emit_sig(p, c, 2, ip->k, 0);
    return p;
}

#ifdef __cplusplus
}
#endif
#endif /* CTMR_SYNTH_H */
