/*
 * ctmr_frontend.h -- CT wire-format front end of libctmr: the step BEFORE the
 * hot path.  The host ships the base64 strings of RFC 6962 §4.6 get-entries responses as they came off
 * the wire; the GPU decodes them, undoes the TLS framing, identifies Chain[0], and feeds the leaves to
 * the same map/reduce kernels as ctmr_process_batch -- no per-entry parse on the host.
 *
 * What it replaces in the reference (jcjones/ct-mapreduce):
 *   - the base64 decode of "leaf_input" / "extra_data" that encoding/json performs inside
 *     ct-go jsonclient for LogClient.GetRawEntries (cmd/ct-fetch/ct-fetch.go:424);
 *   - ct.LogEntryFromLeaf(index, &entry) (cmd/ct-fetch/ct-fetch.go:452): tls.Unmarshal of the
 *     MerkleTreeLeaf (RFC 6962 §3.4), of the CertificateChain / PrecertChainEntry in extra_data
 *     (§4.6), MerkleTreeLeaf.X509Certificate() / .Precertificate() and the "entry dropped" decision
 *     at ct-fetch.go:453-460;
 *   - the entry-type switch of insertCTWorker (ct-fetch.go:198-204): X509 entries use the leaf
 *     certificate, precert entries use PrecertChainEntry.pre_certificate (Precert.Submitted);
 *   - x509.ParseCertificate(Chain[0]) + NewIssuer (ct-fetch.go:215-225, storage/types.go:109-130),
 *     done once per distinct Chain[0] instead of once per entry;
 *   and then everything ctmr_process_batch replaces (include/ctmr.h).
 *
 * github.com/google/certificate-transparency-go v1.1.0 (go.mod:10) is not vendored under the reference,
 * so the TLS rules are restated from RFC 6962 and that version's published struct tags; see
 * oracle/ctmr_oracle_frontend.c ("parity unpinned" for ct-go specifics, pinned against RFC 6962 vectors
 * built in tests/test_frontend_oracle.py).
 */
#ifndef CTMR_FRONTEND_H
#define CTMR_FRONTEND_H

#include "ctmr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* per-entry outcome of the front end ("Erroneous certificate", ct-fetch.go:453-460, unless noted) */
enum {
    CTMR_FE_OK = 0,
    CTMR_FE_BAD_BASE64 = 1,   /* a string is not padded standard base64 (in the reference the whole page fails in
                                 encoding/json and GetRawEntries returns the error, ct-fetch.go:425-443) */
    CTMR_FE_BAD_LEAF = 2,     /* leaf_input is not a v1 timestamped MerkleTreeLeaf, or has trailing data */
    CTMR_FE_UNKNOWN_TYPE = 3, /* entry_type is neither x509_entry(0) nor precert_entry(1) */
    CTMR_FE_BAD_EXTRA = 4,    /* extra_data is not the chain structure its entry type requires */
    CTMR_FE_BAD_CERT = 5      /* x509: the leaf certificate, precert: the TBSCertificate of the leaf, has a fatal
                                 parse error ("failed to parse (pre)certificate in MerkleTreeLeaf") */
};

#define CTMR_ENTRY_X509 0u
#define CTMR_ENTRY_PRECERT 1u
#define CTMR_ENTRY_OTHER 0xFFu

/* n entries of one or more get-entries pages.  HOST buffers.  String i's characters are
 * text[off[i] .. off[i]+len[i]): the contents of the JSON strings, quotes excluded, JSON escapes (if a
 * log emitted any -- base64 needs none) already undone by the caller.  Strings may lie anywhere in
 * `text`, in any order, e.g. in place inside the raw HTTP bodies. */
typedef struct ctmr_raw_batch {
    const uint8_t* text;
    uint64_t text_bytes;
    const uint64_t* leaf_input_off; /* [n] */
    const uint32_t* leaf_input_len; /* [n] characters */
    const uint64_t* extra_data_off; /* [n] */
    const uint32_t* extra_data_len; /* [n] */
    uint64_t n;
    int64_t now_unix_ns;            /* replaces time.Now() at ct-fetch.go:52 */
} ctmr_raw_batch;

/* Caller-allocated outputs, each [n]; any pointer may be NULL. */
typedef struct ctmr_raw_out {
    ctmr_out path;            /* exactly the outputs of ctmr_process_batch; entries with entry_status != 0 and
                                 x509 entries whose leaf does not parse carry status CTMR_ST_PARSE_ERR */
    uint8_t* entry_status;    /* CTMR_FE_* */
    uint8_t* entry_type;      /* CTMR_ENTRY_* (LogEntryType, RFC 6962 §3.1) */
    uint64_t* timestamp_ms;   /* TimestampedEntry.timestamp (feeds uint64ToTimestamp, ct-fetch.go:478) */
    uint32_t* issuer;         /* dense issuer index of Chain[0] (ctmr_issuer_digest), CTMR_ISSUER_NONE, CTMR_ISSUER_BAD */
    /* where the certificate the path processed lies, so that the host can cut the DER of the few NEW
     * certificates out of its own copy of the text without a device round trip (PEM for
     * StoreCertificatePEM, filesystemdatabase.go:197-202): */
    uint8_t* leaf_src;        /* 0: inside the decoded leaf_input, 1: inside the decoded extra_data */
    uint32_t* leaf_off;       /* byte offset inside that decoded string */
    uint32_t* leaf_len;
} ctmr_raw_out;

/* Decode + frame + identify issuers + map + reduce.  Synchronous.  Entries take the global indices
 * following those of earlier ctmr_process_batch / ctmr_process_raw calls on this ctx. */
int ctmr_process_raw(ctmr_ctx* ctx, const ctmr_raw_batch* batch, ctmr_raw_out* out);

/* The same on every GPU of a group (ctmr_group_create, ctmr.h): each round takes one chunk per GPU out of ONE contiguous
 * window of the batch, every GPU decodes and frames its chunk, Chain[0] certificates resolve against the group's issuer
 * registry, and the path runs as one round of the group's key exchange -- entry i keeps global index next_index + i, so
 * the outputs equal those of ctmr_process_raw on one GPU for the same pages. */
int ctmr_group_process_raw(ctmr_group* g, const ctmr_raw_batch* batch, ctmr_raw_out* out);

/* CUDA-event timings of the last ctmr_process_raw call: front-end kernels (decode, framing, issuer
 * identification) and the map/reduce that followed, summed over its chunks, in milliseconds */
int ctmr_frontend_profile_last(ctmr_ctx* ctx, float* frontend_ms, float* path_ms, uint64_t* frontend_launches);

#ifdef __cplusplus
}
#endif
#endif /* CTMR_FRONTEND_H */
