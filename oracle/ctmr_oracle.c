/*
 * ctmr_oracle.c -- CPU ORACLE (test infrastructure only; see ctmr_oracle.h for the rules).
 * Every function cites the reference file:line it restates.  Reference root: jcjones/ct-mapreduce
 * @ 739eda2.  "ct-go" = github.com/google/certificate-transparency-go v1.1.0 (go.mod:10), absent
 * from the tree; its behaviour is restated from RFC 5280 / X.690 and Go's encoding/asn1 rules.
 */
#include "ctmr_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../ct_mapreduce_b200/csrc/ctmr_synth.h"

/* ------------------------------------------------------------------ SHA-256 (FIPS 180-4) */
/* Go crypto/sha256.Sum256, used at storage/types.go:156; also the whole-certificate fingerprint
 * the north_star adds. */

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))

static void sha256_block(uint32_t st[8], const uint8_t* b) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], bb = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t t1 = h + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    st[0] += a; st[1] += bb; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

void ora_sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    for (size_t i = 0; i < full; ++i) sha256_block(st, msg + 64 * i);
    uint8_t tail[128];
    size_t rem = len - 64 * full;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, msg + 64 * full, rem);
    tail[rem] = 0x80;
    size_t tl = rem < 56 ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}

/* ------------------------------------------------------------------ base64url (padded) */
/* Go encoding/base64.URLEncoding, storage/types.go:157 */
size_t ora_b64url(const uint8_t* in, size_t n, char* out) {
    static const char A[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
    size_t o = 0, i = 0;
    for (; i + 3 <= n; i += 3) {
        uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8) | in[i + 2];
        out[o++] = A[v >> 18]; out[o++] = A[(v >> 12) & 63]; out[o++] = A[(v >> 6) & 63]; out[o++] = A[v & 63];
    }
    if (n - i == 1) {
        uint32_t v = (uint32_t)in[i] << 16;
        out[o++] = A[v >> 18]; out[o++] = A[(v >> 12) & 63]; out[o++] = '='; out[o++] = '=';
    } else if (n - i == 2) {
        uint32_t v = ((uint32_t)in[i] << 16) | ((uint32_t)in[i + 1] << 8);
        out[o++] = A[v >> 18]; out[o++] = A[(v >> 12) & 63]; out[o++] = A[(v >> 6) & 63]; out[o++] = '=';
    }
    out[o] = 0;
    return o;
}

/* ------------------------------------------------------------------ DER walking */
/* Restates the subset of ct-go x509.ParseCertificate the path reads (call sites
 * cmd/ct-fetch/ct-fetch.go:202,221; fields listed in SURVEY.md §8(a) a3) and
 * tbsCertWithRawSerial / NewSerial (storage/types.go:165-178). */

typedef struct { uint8_t tag; uint32_t hdr; uint32_t len; } tlv_t;

/* Go encoding/asn1 parseTagAndLength: definite, minimal lengths only; multi-byte tags rejected here */
static int read_tlv(const uint8_t* d, size_t pos, size_t end, tlv_t* t) {
    if (pos + 2 > end) return -1;
    t->tag = d[pos];
    if ((t->tag & 0x1f) == 0x1f) return -1;
    uint8_t l = d[pos + 1];
    if (l < 0x80) {
        t->hdr = 2; t->len = l;
    } else {
        uint32_t nb = l & 0x7f;
        if (nb == 0 || nb > 4) return -1; /* indefinite / too large */
        if (pos + 2 + nb > end) return -1;
        uint32_t v = 0;
        for (uint32_t i = 0; i < nb; ++i) {
            if (v >= (1u << 23)) return -1;
            v = (v << 8) | d[pos + 2 + i];
            if (v == 0) return -1; /* superfluous leading zeros */
        }
        if (v < 0x80) return -1; /* non-minimal */
        t->hdr = 2 + nb; t->len = v;
    }
    if (pos + t->hdr + (size_t)t->len > end) return -1;
    return 0;
}

static int64_t days_from_civil(int64_t y, uint32_t m, uint32_t d) {
    y -= m <= 2;
    int64_t era = (y >= 0 ? y : y - 399) / 400;
    uint32_t yoe = (uint32_t)(y - era * 400);
    uint32_t doy = (153 * (m > 2 ? m - 3 : m + 9) + 2) / 5 + d - 1;
    uint32_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}

static int two(const uint8_t* p, uint32_t* v) {
    if (p[0] < '0' || p[0] > '9' || p[1] < '0' || p[1] > '9') return -1;
    *v = (uint32_t)(p[0] - '0') * 10u + (uint32_t)(p[1] - '0');
    return 0;
}

/* Go asn1 parseUTCTime ("0601021504Z0700" then "060102150405Z0700", yy>=50 -> 19yy) and
 * parseGeneralizedTime ("20060102150405Z0700"); both re-serialise and compare, so a numeric
 * zero offset (which formats back as "Z") is rejected. */
static int parse_time(uint8_t tag, const uint8_t* p, uint32_t len, int64_t* out) {
    uint32_t yy, cc = 0, mo, dd, hh, mi, ss = 0, pos = 0;
    int64_t year;
    if (tag == 0x18) {
        if (len < 15) return -1;
        if (two(p, &cc) || two(p + 2, &yy)) return -1;
        year = (int64_t)cc * 100 + yy;
        pos = 4;
    } else if (tag == 0x17) {
        if (len < 11) return -1;
        if (two(p, &yy)) return -1;
        year = yy >= 50 ? 1900 + (int64_t)yy : 2000 + (int64_t)yy;
        pos = 2;
    } else {
        return -1;
    }
    if (two(p + pos, &mo) || two(p + pos + 2, &dd) || two(p + pos + 4, &hh) || two(p + pos + 6, &mi)) return -1;
    pos += 8;
    int has_sec = 0;
    if (pos + 2 <= len && p[pos] >= '0' && p[pos] <= '9') {
        if (two(p + pos, &ss)) return -1;
        pos += 2;
        has_sec = 1;
    }
    if (tag == 0x18 && !has_sec) return -1;
    if (pos >= len) return -1;
    int64_t off = 0;
    if (p[pos] == 'Z') {
        if (pos + 1 != len) return -1;
    } else if (p[pos] == '+' || p[pos] == '-') {
        uint32_t oh, om;
        if (pos + 5 != len) return -1;
        if (two(p + pos + 1, &oh) || two(p + pos + 3, &om)) return -1;
        if (oh > 23 || om > 59) return -1;
        off = (int64_t)oh * 3600 + (int64_t)om * 60;
        if (off == 0) return -1;
        if (p[pos] == '-') off = -off;
    } else {
        return -1;
    }
    static const uint8_t dim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if (mo < 1 || mo > 12 || dd < 1 || hh > 23 || mi > 59 || ss > 59) return -1;
    uint32_t maxd = dim[mo - 1];
    if (mo == 2 && (year % 4 == 0) && (year % 100 != 0 || year % 400 == 0)) maxd = 29;
    if (dd > maxd) return -1;
    *out = days_from_civil(year, mo, dd) * 86400 + (int64_t)hh * 3600 + (int64_t)mi * 60 + ss - off;
    return 0;
}

/* pkix.Name.FillFromRDNSequence: CommonName = value of the LAST 2.5.4.3 attribute whose value
 * decodes to a Go string (Printable/Numeric/IA5/T61/UTF8String). */
static int walk_name(const uint8_t* d, size_t pos, size_t end, ora_cert* out, int record_cn) {
    while (pos < end) {
        tlv_t set;
        if (read_tlv(d, pos, end, &set) || set.tag != 0x31) return -1;
        size_t sp = pos + set.hdr, se = sp + set.len;
        while (sp < se) {
            tlv_t atv, oid, val;
            if (read_tlv(d, sp, se, &atv) || atv.tag != 0x30) return -1;
            size_t ap = sp + atv.hdr, ae = ap + atv.len;
            if (read_tlv(d, ap, ae, &oid) || oid.tag != 0x06) return -1;
            size_t vp = ap + oid.hdr + oid.len;
            if (read_tlv(d, vp, ae, &val)) return -1;
            if (record_cn && oid.len == 3 && d[ap + oid.hdr] == 0x55 && d[ap + oid.hdr + 1] == 0x04 &&
                d[ap + oid.hdr + 2] == 0x03) {
                uint8_t t = val.tag;
                if (t == 0x0c || t == 0x13 || t == 0x16 || t == 0x14 || t == 0x12) {
                    out->cn_off = (uint32_t)(vp + val.hdr);
                    out->cn_len = val.len;
                    out->has_cn = 1;
                }
            }
            sp += atv.hdr + atv.len;
        }
        pos += set.hdr + set.len;
    }
    return 0;
}

/* TBSCertificate at d[pos..cend): everything parseCertificate reads from it.  *tbs_end = offset just past it. */
static int parse_tbs(const uint8_t* d, size_t pos, size_t cend, ora_cert* out, size_t* tbs_end) {
    tlv_t tbs, t;
    if (read_tlv(d, pos, cend, &tbs) || tbs.tag != 0x30) return -1;
    out->tbs_off = (uint32_t)pos;
    out->tbs_len = tbs.hdr + tbs.len;
    size_t tp = pos + tbs.hdr, tend = tp + tbs.len;
    /* Version: optional, explicit, tag 0 */
    if (read_tlv(d, tp, tend, &t)) return -1;
    if (t.tag == 0xa0) {
        tp += t.hdr + t.len;
        if (read_tlv(d, tp, tend, &t)) return -1;
    }
    /* SerialNumber INTEGER -- raw content octets kept (types.go:165-178; types_test.go:81-101) */
    if (t.tag != 0x02 || t.len == 0) return -2;
    if (t.len > 1) { /* Go asn1 parseBigInt: integer not minimally-encoded */
        uint8_t b0 = d[tp + t.hdr], b1 = d[tp + t.hdr + 1];
        if ((b0 == 0x00 && (b1 & 0x80) == 0) || (b0 == 0xff && (b1 & 0x80) != 0)) return -2;
    }
    out->serial_off = (uint32_t)(tp + t.hdr);
    out->serial_len = t.len;
    tp += t.hdr + t.len;
    /* SignatureAlgorithm */
    if (read_tlv(d, tp, tend, &t) || t.tag != 0x30) return -3;
    tp += t.hdr + t.len;
    /* Issuer */
    if (read_tlv(d, tp, tend, &t) || t.tag != 0x30) return -4;
    out->issuer_off = (uint32_t)tp;
    out->issuer_len = t.hdr + t.len;
    if (walk_name(d, tp + t.hdr, tp + t.hdr + t.len, out, 1)) return -4;
    tp += t.hdr + t.len;
    /* Validity */
    if (read_tlv(d, tp, tend, &t) || t.tag != 0x30) return -5;
    {
        size_t vp = tp + t.hdr, ve = vp + t.len;
        tlv_t a, b;
        if (read_tlv(d, vp, ve, &a)) return -5;
        if (parse_time(a.tag, d + vp + a.hdr, a.len, &out->not_before)) return -5;
        vp += a.hdr + a.len;
        if (read_tlv(d, vp, ve, &b)) return -5;
        if (parse_time(b.tag, d + vp + b.hdr, b.len, &out->not_after)) return -5;
    }
    tp += t.hdr + t.len;
    /* Subject */
    if (read_tlv(d, tp, tend, &t) || t.tag != 0x30) return -6;
    {
        ora_cert scratch;
        if (walk_name(d, tp + t.hdr, tp + t.hdr + t.len, &scratch, 0)) return -6;
    }
    tp += t.hdr + t.len;
    /* SubjectPublicKeyInfo: SEQ { AlgorithmIdentifier, BIT STRING } */
    if (read_tlv(d, tp, tend, &t) || t.tag != 0x30) return -7;
    out->spki_off = (uint32_t)tp;
    out->spki_len = t.hdr + t.len;
    {
        size_t kp = tp + t.hdr, ke = kp + t.len;
        tlv_t a, b;
        if (read_tlv(d, kp, ke, &a) || a.tag != 0x30) return -7;
        kp += a.hdr + a.len;
        if (read_tlv(d, kp, ke, &b) || b.tag != 0x03 || b.len == 0) return -7;
    }
    tp += t.hdr + t.len;
    /* optional [1] issuerUniqueID, [2] subjectUniqueID, [3] extensions */
    if (tp < tend) {
        if (read_tlv(d, tp, tend, &t)) return -8;
        if (t.tag == 0x81 || t.tag == 0xa1) {
            tp += t.hdr + t.len;
            if (tp < tend && read_tlv(d, tp, tend, &t)) return -8;
        }
    }
    if (tp < tend && (t.tag == 0x82 || t.tag == 0xa2)) {
        tp += t.hdr + t.len;
        if (tp < tend && read_tlv(d, tp, tend, &t)) return -8;
    }
    if (tp < tend && t.tag == 0xa3) {
        size_t xp = tp + t.hdr, xe = xp + t.len;
        tlv_t seq;
        if (read_tlv(d, xp, xe, &seq) || seq.tag != 0x30) return -9;
        size_t ep = xp + seq.hdr, ee = ep + seq.len;
        while (ep < ee) {
            tlv_t ext, oid, v;
            if (read_tlv(d, ep, ee, &ext) || ext.tag != 0x30) return -9;
            size_t ip = ep + ext.hdr, ie = ip + ext.len;
            if (read_tlv(d, ip, ie, &oid) || oid.tag != 0x06) return -9;
            size_t oidp = ip + oid.hdr;
            ip += oid.hdr + oid.len;
            if (read_tlv(d, ip, ie, &v)) return -9;
            if (v.tag == 0x01) { /* critical BOOLEAN */
                if (v.len != 1 || (d[ip + v.hdr] != 0x00 && d[ip + v.hdr] != 0xff)) return -9;
                ip += v.hdr + v.len;
                if (read_tlv(d, ip, ie, &v)) return -9;
            }
            if (v.tag != 0x04) return -9;
            size_t vp = ip + v.hdr, ve = vp + v.len;
            if (oid.len == 3 && d[oidp] == 0x55 && d[oidp + 1] == 0x1d) {
                if (d[oidp + 2] == 0x13) {
                    /* basicConstraints ::= SEQ { cA BOOLEAN DEFAULT FALSE, pathLen INTEGER OPTIONAL } */
                    tlv_t bc, f;
                    if (read_tlv(d, vp, ve, &bc) || bc.tag != 0x30) return -10;
                    if (vp + bc.hdr + bc.len != ve) return -10; /* trailing data after BasicConstraints */
                    size_t bp = vp + bc.hdr, be = bp + bc.len;
                    int ca = 0;
                    if (bp < be) {
                        if (read_tlv(d, bp, be, &f)) return -10;
                        if (f.tag == 0x01) {
                            if (f.len != 1) return -10;
                            uint8_t bv = d[bp + f.hdr];
                            if (bv != 0x00 && bv != 0xff) return -10; /* asn1: invalid boolean */
                            ca = bv != 0;
                            bp += f.hdr + f.len;
                            if (bp < be && read_tlv(d, bp, be, &f)) return -10;
                        }
                        if (bp < be) {
                            if (f.tag != 0x02 || f.len == 0) return -10;
                            bp += f.hdr + f.len;
                        }
                    }
                    out->bc_valid = 1;
                    out->is_ca = ca;
                } else if (d[oidp + 2] == 0x1f) {
                    out->crldp_off = (uint32_t)vp;
                    out->crldp_len = v.len;
                }
            }
            ep += ext.hdr + ext.len;
        }
    }
    /* encoding/asn1 tolerates extra trailing elements inside a SEQUENCE parsed into a struct */
    *tbs_end = pos + tbs.hdr + tbs.len;
    return 0;
}

int ora_parse_cert(const uint8_t* d, size_t len, ora_cert* out) {
    memset(out, 0, sizeof *out);
    tlv_t cert, t;
    /* x509.ParseCertificate: asn1.Unmarshal into certificate{}, trailing data is an error */
    if (read_tlv(d, 0, len, &cert) || cert.tag != 0x30) return -1;
    if ((size_t)cert.hdr + cert.len != len) return -1;
    size_t cend = len, pos = cert.hdr;
    int rc = parse_tbs(d, pos, cend, out, &pos);
    if (rc) return rc;
    if (read_tlv(d, pos, cend, &t) || t.tag != 0x30) return -11; /* signatureAlgorithm */
    pos += t.hdr + t.len;
    if (read_tlv(d, pos, cend, &t) || t.tag != 0x03 || t.len == 0) return -12; /* signatureValue */
    return 0;
}

/* ct-go x509.ParseTBSCertificate (used by MerkleTreeLeaf.Precertificate(), reached from
 * ct.LogEntryFromLeaf, cmd/ct-fetch/ct-fetch.go:452): asn1.Unmarshal into tbsCertificate{} with
 * "trailing data" an error, then the same parseCertificate body as a full certificate. */
int ora_parse_tbs(const uint8_t* d, size_t len, ora_cert* out) {
    memset(out, 0, sizeof *out);
    size_t end = 0;
    int rc = parse_tbs(d, 0, len, out, &end);
    if (rc) return rc;
    return end == len ? 0 : -13;
}

/* ------------------------------------------------------------------ value types */

/* storage/types.go:124-130,155-159: Issuer.ID() = base64url(SHA-256(RawSubjectPublicKeyInfo)) */
void ora_issuer_id(const uint8_t* spki, size_t len, uint8_t digest[32], char id[45]) {
    ora_sha256(spki, len, digest);
    ora_b64url(digest, 32, id);
}

/* storage/types.go:339-346: NewExpDateFromTime = t.Truncate(time.Hour) -> floor(unix/3600) */
int64_t ora_exp_hour(int64_t s) { return s >= 0 ? s / 3600 : -((-s + 3599) / 3600); }

static void civil_from_days(int64_t z, int64_t* y, uint32_t* m, uint32_t* d) {
    z += 719468;
    int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    uint32_t doe = (uint32_t)(z - era * 146097);
    uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    uint32_t mp = (5 * doy + 2) / 153;
    *d = doy - (153 * mp + 2) / 5 + 1;
    *m = mp < 10 ? mp + 3 : mp - 9;
    *y = (int64_t)yoe + era * 400 + (*m <= 2);
}

/* storage/types.go:379-384: ExpDate.ID() with kExpirationFormatWithHour "2006-01-02-15" */
void ora_expdate_id(int64_t exp_hour, char out[14]) {
    int64_t days = exp_hour >= 0 ? exp_hour / 24 : -((-exp_hour + 23) / 24);
    int64_t y; uint32_t m, d;
    civil_from_days(days, &y, &m, &d);
    snprintf(out, 14, "%04lld-%02u-%02u-%02u", (long long)y, m, d, (unsigned)(exp_hour - days * 24));
}

/* storage/filesystemdatabase.go:140-143: markDirty id, kExpirationFormat "2006-01-02" */
void ora_day_id(int64_t sec, char out[11]) {
    int64_t days = sec >= 0 ? sec / 86400 : -((-sec + 86399) / 86400);
    int64_t y; uint32_t m, d;
    civil_from_days(days, &y, &m, &d);
    snprintf(out, 11, "%04lld-%02u-%02u", (long long)y, m, d);
}

/* cmd/ct-fetch/ct-fetch.go:44-70 certIsFilteredOut, same order, same string operations:
 * strings.Split(filter, ",") without trimming; HasPrefix on the raw CommonName bytes. */
int ora_filter(const uint8_t* der, const ora_cert* c, const uint8_t* filter, size_t flen, int log_expired,
               int64_t now_ns) {
    if (c->bc_valid && c->is_ca) return ORA_ST_FILTER_CA;
    /* NotAfter.Before(now): NotAfter has 1 s resolution */
    int64_t now_s = now_ns >= 0 ? now_ns / 1000000000LL : -((-now_ns + 999999999LL) / 1000000000LL);
    int64_t now_frac = now_ns - now_s * 1000000000LL;
    int before = c->not_after < now_s || (c->not_after == now_s && now_frac > 0);
    if (before && !log_expired) return ORA_ST_FILTER_EXPIRED;
    int skip = flen != 0;
    size_t start = 0;
    const uint8_t* cn = der + c->cn_off;
    size_t cnl = c->has_cn ? c->cn_len : 0;
    for (size_t i = 0; i <= flen && skip; ++i) {
        if (i == flen || filter[i] == ',') {
            size_t pl = i - start;
            if (pl <= cnl && memcmp(cn, filter + start, pl) == 0) skip = 0;
            start = i + 1;
        }
    }
    return skip ? ORA_ST_FILTER_CN : 0;
}

/* ------------------------------------------------------------------ byte-string hash map */

typedef struct { uint64_t h; uint32_t klen; uint32_t pad; uint8_t* key; uint64_t val; } bs_ent;
typedef struct { bs_ent* e; uint64_t cap, n; } bs_map;

static uint64_t fnv1a(const uint8_t* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h ? h : 1;
}

static void bs_init(bs_map* m) { m->cap = 1024; m->n = 0; m->e = (bs_ent*)calloc(m->cap, sizeof(bs_ent)); }
static void bs_free(bs_map* m) {
    for (uint64_t i = 0; i < m->cap; ++i) free(m->e[i].key);
    free(m->e);
}
static bs_ent* bs_find(bs_map* m, const uint8_t* k, size_t kl, uint64_t h) {
    uint64_t i = h & (m->cap - 1);
    for (;;) {
        bs_ent* e = &m->e[i];
        if (!e->key) return e;
        if (e->h == h && e->klen == kl && memcmp(e->key, k, kl) == 0) return e;
        i = (i + 1) & (m->cap - 1);
    }
}
static void bs_grow(bs_map* m) {
    bs_map nm;
    nm.cap = m->cap * 2; nm.n = m->n;
    nm.e = (bs_ent*)calloc(nm.cap, sizeof(bs_ent));
    for (uint64_t i = 0; i < m->cap; ++i) {
        if (!m->e[i].key) continue;
        uint64_t j = m->e[i].h & (nm.cap - 1);
        while (nm.e[j].key) j = (j + 1) & (nm.cap - 1);
        nm.e[j] = m->e[i];
    }
    free(m->e);
    *m = nm;
}
/* returns the entry; *created = 1 when newly inserted (val initialised to 0) */
static bs_ent* bs_put(bs_map* m, const uint8_t* k, size_t kl, int* created) {
    if ((m->n + 1) * 10 > m->cap * 6) bs_grow(m);
    uint64_t h = fnv1a(k, kl);
    bs_ent* e = bs_find(m, k, kl, h);
    if (e->key) { *created = 0; return e; }
    e->key = (uint8_t*)malloc(kl ? kl : 1);
    memcpy(e->key, k, kl);
    e->klen = (uint32_t)kl; e->h = h; e->val = 0;
    m->n++;
    *created = 1;
    return e;
}

/* ------------------------------------------------------------------ RemoteCache sets */
/* storage/mockcache.go:38-61 SetInsert: sorted unique strings, true iff newly added;
 * :120-122 SetCardinality = len.  Members are binary-safe (types.go:218-220). */
struct ora_cache { bs_map members; /* key = set_key 0x00 member? no: (u32 klen, set_key, member) */ bs_map sets; };

ora_cache* ora_cache_new(void) {
    ora_cache* c = (ora_cache*)calloc(1, sizeof *c);
    bs_init(&c->members); bs_init(&c->sets);
    return c;
}
void ora_cache_free(ora_cache* c) { if (c) { bs_free(&c->members); bs_free(&c->sets); free(c); } }

static size_t compose(uint8_t* buf, const char* key, size_t kl, const uint8_t* m, size_t ml) {
    uint32_t k32 = (uint32_t)kl;
    memcpy(buf, &k32, 4); memcpy(buf + 4, key, kl); memcpy(buf + 4 + kl, m, ml);
    return 4 + kl + ml;
}

int ora_cache_set_insert(ora_cache* c, const char* key, size_t kl, const uint8_t* member, size_t ml) {
    uint8_t stack[512];
    uint8_t* buf = (4 + kl + ml <= sizeof stack) ? stack : (uint8_t*)malloc(4 + kl + ml);
    size_t n = compose(buf, key, kl, member, ml);
    int created;
    bs_put(&c->members, buf, n, &created);
    if (created) {
        int cs;
        bs_ent* s = bs_put(&c->sets, (const uint8_t*)key, kl, &cs);
        s->val++;
    }
    if (buf != stack) free(buf);
    return created;
}

uint64_t ora_cache_set_cardinality(ora_cache* c, const char* key, size_t kl) {
    bs_ent* e = bs_find(&c->sets, (const uint8_t*)key, kl, fnv1a((const uint8_t*)key, kl));
    return e->key ? e->val : 0;
}

typedef struct { const uint8_t* p; uint32_t n; } span_t;
static int span_cmp(const void* a, const void* b) { /* Go strings.Compare: bytewise, shorter first on tie */
    const span_t *x = (const span_t*)a, *y = (const span_t*)b;
    uint32_t m = x->n < y->n ? x->n : y->n;
    int r = memcmp(x->p, y->p, m);
    if (r) return r;
    return x->n < y->n ? -1 : (x->n > y->n);
}

uint64_t ora_cache_set_list(ora_cache* c, const char* key, size_t kl, uint8_t* buf, size_t cap, size_t* used) {
    uint64_t cnt = 0, cap_sp = 16;
    span_t* sp = (span_t*)malloc(cap_sp * sizeof *sp);
    for (uint64_t i = 0; i < c->members.cap; ++i) {
        bs_ent* e = &c->members.e[i];
        if (!e->key || e->klen < 4 + kl) continue;
        uint32_t k32; memcpy(&k32, e->key, 4);
        if (k32 != kl || memcmp(e->key + 4, key, kl)) continue;
        if (cnt == cap_sp) { cap_sp *= 2; sp = (span_t*)realloc(sp, cap_sp * sizeof *sp); }
        sp[cnt].p = e->key + 4 + kl; sp[cnt].n = e->klen - 4 - (uint32_t)kl; cnt++;
    }
    qsort(sp, cnt, sizeof *sp, span_cmp);
    size_t o = 0;
    for (uint64_t i = 0; i < cnt; ++i) {
        if (o + 4 + sp[i].n > cap) break;
        memcpy(buf + o, &sp[i].n, 4); memcpy(buf + o + 4, sp[i].p, sp[i].n);
        o += 4 + sp[i].n;
    }
    *used = o;
    free(sp);
    return cnt;
}

/* storage/knowncertificates.go:28-34: "serials::" + expDate.ID() + "::" + issuer.ID() */
size_t ora_serials_key(int64_t exp_hour, const char* issuer_id, char* out, size_t cap) {
    char e[14];
    ora_expdate_id(exp_hour, e);
    return (size_t)snprintf(out, cap, "serials::%s::%s", e, issuer_id);
}

/* storage/knowncertificates.go:38-55 WasUnknown */
int ora_was_unknown(ora_cache* c, int64_t exp_hour, const char* issuer_id, const uint8_t* serial, size_t sl) {
    char key[160];
    size_t kl = ora_serials_key(exp_hour, issuer_id, key, sizeof key);
    return ora_cache_set_insert(c, key, kl, serial, sl);
}

/* ------------------------------------------------------------------ the composed path */

typedef struct { int status; uint8_t digest[32]; char id[45]; } ora_issuer_rec;

struct ora_db {
    uint8_t* filter; size_t filter_len; int log_expired;
    ora_cache* cache;
    bs_map issuer_hours;  /* IssuerMetadata.knownExpDates across all issuers: key = digest||exp_hour */
    bs_map issuer_strings; /* knownIssuerDNs / knownCrlDPs by raw bytes: key = digest||kind||bytes */
    bs_map set_meta;      /* set key -> index into metas */
    struct { uint8_t digest[32]; int64_t exp_hour; char key[160]; uint32_t key_len; }* metas;
    uint64_t n_metas, cap_metas;
    uint64_t counters[8];
};

ora_db* ora_db_new(const uint8_t* filter, size_t flen, int log_expired) {
    ora_db* db = (ora_db*)calloc(1, sizeof *db);
    db->filter = (uint8_t*)malloc(flen ? flen : 1);
    if (flen) memcpy(db->filter, filter, flen);
    db->filter_len = flen; db->log_expired = log_expired;
    db->cache = ora_cache_new();
    bs_init(&db->issuer_hours); bs_init(&db->set_meta); bs_init(&db->issuer_strings);
    return db;
}
void ora_db_free(ora_db* db) {
    if (!db) return;
    ora_cache_free(db->cache); bs_free(&db->issuer_hours); bs_free(&db->set_meta); bs_free(&db->issuer_strings);
    free(db->metas); free(db->filter); free(db);
}

typedef struct {
    const uint8_t* blob; const uint64_t* offsets; uint64_t lo, hi;
    const uint8_t* filter; size_t flen; int log_expired; int64_t now_ns;
    uint8_t* status; uint8_t* sha; int64_t* exp_hour; uint32_t* soff; uint32_t* slen; uint64_t kept;
    uint32_t *noff, *nlen, *coff, *clen;
} map_job;

/* map half of insertCTWorker (ct-fetch.go:198-213) + the whole-certificate fingerprint */
static void* map_worker(void* arg) {
    map_job* j = (map_job*)arg;
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        const uint8_t* der = j->blob + j->offsets[i];
        size_t len = (size_t)(j->offsets[i + 1] - j->offsets[i]);
        if (j->sha) ora_sha256(der, len, j->sha + 32 * i);
        ora_cert c;
        int st;
        if (ora_parse_cert(der, len, &c)) {
            st = ORA_ST_PARSE_ERR;
            if (j->exp_hour) { j->exp_hour[i] = 0; j->soff[i] = 0; j->slen[i] = 0; }
            if (j->noff) { j->noff[i] = j->nlen[i] = j->coff[i] = j->clen[i] = 0; }
        } else {
            st = ora_filter(der, &c, j->filter, j->flen, j->log_expired, j->now_ns);
            if (j->exp_hour) { j->exp_hour[i] = ora_exp_hour(c.not_after); j->soff[i] = c.serial_off; j->slen[i] = c.serial_len; }
            if (j->noff) { j->noff[i] = c.issuer_off; j->nlen[i] = c.issuer_len; j->coff[i] = c.crldp_off; j->clen[i] = c.crldp_len; }
        }
        if (j->status) j->status[i] = (uint8_t)st;
        if (st == 0) j->kept++;
    }
    return NULL;
}

static uint64_t run_map(map_job* proto, uint64_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n) nthreads = n ? (int)n : 1;
    map_job* jobs = (map_job*)malloc(sizeof(map_job) * nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = n * t / nthreads; jobs[t].hi = n * (t + 1) / nthreads; jobs[t].kept = 0;
        if (nthreads > 1) pthread_create(&th[t], NULL, map_worker, &jobs[t]);
    }
    uint64_t kept = 0;
    if (nthreads == 1) map_worker(&jobs[0]);
    for (int t = 0; t < nthreads; ++t) { if (nthreads > 1) pthread_join(th[t], NULL); kept += jobs[t].kept; }
    free(jobs); free(th);
    return kept;
}

uint64_t ora_map_only(const uint8_t* blob, const uint64_t* offsets, uint64_t n, const uint8_t* filter, size_t flen,
                      int log_expired, int64_t now_ns, int nthreads, uint8_t* sha_out) {
    map_job p;
    memset(&p, 0, sizeof p);
    p.blob = blob; p.offsets = offsets; p.filter = filter; p.flen = flen; p.log_expired = log_expired; p.now_ns = now_ns;
    p.sha = sha_out;
    return run_map(&p, n, nthreads);
}

int ora_db_process(ora_db* db, const uint8_t* blob, const uint64_t* offsets, uint64_t n, const uint8_t* iblob,
                   const uint64_t* ioffsets, uint32_t n_issuers, const uint32_t* issuer_idx, int64_t now_ns,
                   int nthreads, ora_out* out) {
    /* issuers: x509.ParseCertificate(Chain[0]) (ct-fetch.go:221) + NewIssuer/Issuer.ID (types.go:109-130) */
    ora_issuer_rec* irec = (ora_issuer_rec*)calloc(n_issuers ? n_issuers : 1, sizeof *irec);
    for (uint32_t k = 0; k < n_issuers; ++k) {
        const uint8_t* der = iblob + ioffsets[k];
        size_t len = (size_t)(ioffsets[k + 1] - ioffsets[k]);
        ora_cert c;
        if (ora_parse_cert(der, len, &c)) { irec[k].status = -1; continue; }
        ora_issuer_id(der + c.spki_off, c.spki_len, irec[k].digest, irec[k].id);
    }
    map_job p;
    memset(&p, 0, sizeof p);
    p.blob = blob; p.offsets = offsets; p.filter = db->filter; p.flen = db->filter_len;
    p.log_expired = db->log_expired; p.now_ns = now_ns;
    p.status = out->status; p.sha = out->sha256; p.exp_hour = out->exp_hour; p.soff = out->serial_off; p.slen = out->serial_len;
    const int meta = out->issuer_name_off && out->issuer_name_len && out->crldp_off && out->crldp_len;
    if (meta) { p.noff = out->issuer_name_off; p.nlen = out->issuer_name_len; p.coff = out->crldp_off; p.clen = out->crldp_len; }
    run_map(&p, n, nthreads);

    /* reduce half, strictly sequential = numThreads 1 (config/config.go:187) */
    for (uint64_t i = 0; i < n; ++i) {
        out->was_unknown[i] = 0;
        out->first_issuer_hour[i] = 0;
        if (meta && out->first_issuer_dn) out->first_issuer_dn[i] = 0;
        if (meta && out->first_crldp) out->first_crldp[i] = 0;
        int st = out->status[i];
        if (st == 0) {
            uint32_t k = issuer_idx[i];
            if (k == 0xFFFFFFFFu || k >= n_issuers) st = ORA_ST_NO_ISSUER;      /* ct-fetch.go:215-219 */
            else if (irec[k].status) st = ORA_ST_ISSUER_PARSE_ERR;             /* ct-fetch.go:221-225 */
            /* no serial-length limit: NewSerial keeps whatever the INTEGER holds (storage/types.go:171-178).  The GPU key
             * record holds 39 octets and DECLINES longer serials (CTMR_ST_SERIAL_TOO_LONG): a divergence of the product,
             * pinned by tests/test_gpu_parity.py::test_long_serials_are_declined_not_misdeduplicated, not modelled here */
        }
        out->status[i] = (uint8_t)st;
        db->counters[st & 7]++;
        if (st) continue;
        /* FilesystemDatabase.Store (filesystemdatabase.go:158-211) */
        const ora_issuer_rec* is = &irec[issuer_idx[i]];
        const uint8_t* serial = blob + offsets[i] + out->serial_off[i];
        char key[160];
        size_t kl = ora_serials_key(out->exp_hour[i], is->id, key, sizeof key);
        int unknown = ora_cache_set_insert(db->cache, key, kl, serial, out->serial_len[i]);
        out->was_unknown[i] = (uint8_t)unknown;
        if (unknown) {
            int created;
            bs_ent* sm = bs_put(&db->set_meta, (const uint8_t*)key, kl, &created);
            if (created) {
                if (db->n_metas == db->cap_metas) {
                    db->cap_metas = db->cap_metas ? db->cap_metas * 2 : 256;
                    db->metas = realloc(db->metas, db->cap_metas * sizeof *db->metas);
                }
                memcpy(db->metas[db->n_metas].digest, is->digest, 32);
                db->metas[db->n_metas].exp_hour = out->exp_hour[i];
                memcpy(db->metas[db->n_metas].key, key, kl + 1);
                db->metas[db->n_metas].key_len = (uint32_t)kl;
                sm->val = db->n_metas++;
            }
            /* IssuerMetadata.Accumulate -> seenExpDateBefore (issuermetadata.go:95-108) */
            uint8_t ih[40];
            memcpy(ih, is->digest, 32); memcpy(ih + 32, &out->exp_hour[i], 8);
            int first;
            bs_put(&db->issuer_hours, ih, 40, &first);
            out->first_issuer_hour[i] = (uint8_t)first;
            if (meta) {
                /* Accumulate's memo lookups for the issuer DN (issuermetadata.go:94,130-135) and the CRL
                 * distribution points (:110-128), on the raw bytes those strings are formatted from */
                const uint8_t* der = blob + offsets[i];
                for (int kind = 1; kind <= 2; ++kind) {
                    const uint32_t so = kind == 1 ? out->issuer_name_off[i] : out->crldp_off[i];
                    const uint32_t sl = kind == 1 ? out->issuer_name_len[i] : out->crldp_len[i];
                    uint8_t* dstbit = kind == 1 ? out->first_issuer_dn : out->first_crldp;
                    if (!sl || !dstbit) continue;
                    uint8_t* kb = (uint8_t*)malloc(33 + sl);
                    memcpy(kb, is->digest, 32); kb[32] = (uint8_t)kind; memcpy(kb + 33, der + so, sl);
                    int fresh;
                    bs_put(&db->issuer_strings, kb, 33 + sl, &fresh);
                    dstbit[i] = (uint8_t)fresh;
                    free(kb);
                }
            }
        }
    }
    free(irec);
    return 0;
}

/* cmd/storage-statistics/storage-statistics.go:44-53: countIssuerSerials = sum over the issuer's
 * expDates of KnownCertificates.Count() = SetCardinality (knowncertificates.go:57-63) */
typedef struct { uint8_t d[32]; uint64_t c; } idc_t;
static int idc_cmp(const void* a, const void* b) { return memcmp(((const idc_t*)a)->d, ((const idc_t*)b)->d, 32); }

uint64_t ora_db_issuer_counts(ora_db* db, uint8_t* ids32, uint64_t* counts, uint64_t cap) {
    bs_map agg;
    bs_init(&agg);
    for (uint64_t m = 0; m < db->n_metas; ++m) {
        int created;
        bs_ent* e = bs_put(&agg, db->metas[m].digest, 32, &created);
        e->val += ora_cache_set_cardinality(db->cache, db->metas[m].key, db->metas[m].key_len);
    }
    idc_t* v = (idc_t*)malloc((agg.n ? agg.n : 1) * sizeof *v);
    uint64_t n = 0;
    for (uint64_t i = 0; i < agg.cap; ++i)
        if (agg.e[i].key) { memcpy(v[n].d, agg.e[i].key, 32); v[n].c = agg.e[i].val; n++; }
    qsort(v, n, sizeof *v, idc_cmp);
    for (uint64_t i = 0; i < n && i < cap; ++i) { memcpy(ids32 + 32 * i, v[i].d, 32); counts[i] = v[i].c; }
    free(v);
    bs_free(&agg);
    return n;
}

uint64_t ora_db_set_cardinality(ora_db* db, int64_t exp_hour, const uint8_t digest[32]) {
    char id[45], key[160];
    ora_b64url(digest, 32, id);
    size_t kl = ora_serials_key(exp_hour, id, key, sizeof key);
    return ora_cache_set_cardinality(db->cache, key, kl);
}

/* What Redis does to the reference's state as time passes: KnownCertificates.WasUnknown sets EXPIREAT(key,
 * expDate) on "serials::<expDate>::<issuer>" (storage/knowncertificates.go:44-47,98-104; expDate = the
 * hour-truncated NotAfter, storage/types.go:371-373), so at time `now` every set with expDate <= now is gone:
 * its members are unknown again and it no longer counts.  IssuerMetadata's seenExpDateBefore memo is process
 * memory, not Redis, and survives.  Returns the number of serials dropped. */
uint64_t ora_db_evict_expired(ora_db* db, int64_t now_sec) {
    ora_cache* c = db->cache;
    uint64_t dropped = 0;
    bs_map nm;
    nm.cap = c->members.cap; nm.n = 0;
    nm.e = (bs_ent*)calloc(nm.cap, sizeof(bs_ent));
    for (uint64_t i = 0; i < c->members.cap; ++i) {
        bs_ent* e = &c->members.e[i];
        if (!e->key) continue;
        uint32_t kl;
        memcpy(&kl, e->key, 4);
        int gone = 0;
        for (uint64_t m = 0; m < db->n_metas && !gone; ++m)
            if (db->metas[m].exp_hour * 3600 <= now_sec && db->metas[m].key_len == kl && memcmp(db->metas[m].key, e->key + 4, kl) == 0) gone = 1;
        if (gone) { free(e->key); dropped++; continue; }
        uint64_t j = e->h & (nm.cap - 1);
        while (nm.e[j].key) j = (j + 1) & (nm.cap - 1);
        nm.e[j] = *e;
        nm.n++;
    }
    free(c->members.e);
    c->members = nm;
    for (uint64_t m = 0; m < db->n_metas; ++m)
        if (db->metas[m].exp_hour * 3600 <= now_sec) {
            uint64_t h = fnv1a((const uint8_t*)db->metas[m].key, db->metas[m].key_len);
            bs_ent* s = bs_find(&c->sets, (const uint8_t*)db->metas[m].key, db->metas[m].key_len, h);
            if (s->key) s->val = 0; /* SCARD of a missing key is 0 */
        }
    return dropped;
}

void ora_db_filter_counters(ora_db* db, uint64_t out[8]) { memcpy(out, db->counters, sizeof db->counters); }

/* ------------------------------------------------------------------ synthetic corpus on the CPU */

uint64_t ora_synth_lengths(const struct ctmr_synth_cfg* c, uint64_t first, uint64_t n, uint64_t* offsets) {
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; ++i) { offsets[i] = o; o += ctmr_synth_cert_len(c, first + i); }
    offsets[n] = o;
    return o;
}

void ora_synth_write(const struct ctmr_synth_cfg* c, uint64_t first, uint64_t n, const uint64_t* offsets, uint8_t* blob) {
    for (uint64_t i = 0; i < n; ++i) {
        ctmr_synth_plan pl;
        ctmr_synth_plan_make(c, first + i, &pl);
        ctmr_synth_cert_write(c, &pl, blob + offsets[i]);
    }
}

void ora_synth_issuer_idx(const struct ctmr_synth_cfg* c, uint64_t first, uint64_t n, uint32_t* idx) {
    for (uint64_t i = 0; i < n; ++i) {
        ctmr_synth_plan pl;
        ctmr_synth_plan_make(c, first + i, &pl);
        idx[i] = pl.issuer;
    }
}

uint64_t ora_synth_issuers(const struct ctmr_synth_cfg* c, uint64_t* offsets, uint8_t* blob, size_t cap) {
    uint64_t o = 0;
    for (uint32_t k = 0; k < c->n_issuers; ++k) {
        ctmr_synth_issuer_plan ip;
        ctmr_synth_issuer_plan_make(c, k, &ip);
        offsets[k] = o;
        if (blob && o + ip.total <= cap) ctmr_synth_issuer_write(c, &ip, blob + o);
        o += ip.total;
    }
    offsets[c->n_issuers] = o;
    return o;
}
