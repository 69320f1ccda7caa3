// ctmr_kernels.cuh -- launch-side declarations shared by ctmr_kernels.cu and ctmr_api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctmr.h"
#include "../../include/ctmr_frontend.h"

namespace ctmr {

// issuerCNFilter, pre-split on ',' by the host (strings.Split, ct-fetch.go:58): prefixes are raw bytes.
struct FilterCfg {
    uint32_t filter_nonempty;  // len(*ctconfig.IssuerCNFilter) != 0
    uint32_t n_prefix;
    uint32_t log_expired;
    uint32_t flags;            // CTMR_F_*
    uint16_t off[33];          // prefix p = bytes[off[p] .. off[p+1])
    uint8_t bytes[446];
};
static_assert(sizeof(FilterCfg) == 16 + 66 + 446, "FilterCfg layout");

// known-certificate slot: 64 bytes = two 32-byte sectors
struct __align__(16) KnownSlot {
    unsigned long long inv_first;  // ~(lowest global entry index seen for this key); 0 = none yet
    unsigned long long tag;        // 0 empty | (hash & ~3)|1 pending | (hash & ~3)|2 ready
    uint32_t body[12];             // exp_hour, issuer, {serial_len, serial[39]}
};
static_assert(sizeof(KnownSlot) == 64, "KnownSlot layout");

struct __align__(16) PairSlot {   // one per Redis set "serials::<expDate>::<issuer>" ever written
    unsigned long long key;        // ((issuer << 32) | (uint32)exp_hour) + 1; 0 = empty
    unsigned long long inv_first;  // ~(lowest index among was-unknown entries of this (issuer, hour))
    unsigned long long count;      // SetCardinality of the set: unique serials inserted and not yet expired
    unsigned long long pad;
};
static_assert(sizeof(PairSlot) == 32, "PairSlot layout");


struct __align__(16) MetaSlot {      // IssuerMetadata string sets: identity = two independent 64-bit hashes
    unsigned long long h1;           // 0 = empty
    unsigned long long h2;           // 0 = not yet published
    unsigned long long inv_first;    // ~(lowest index among new certificates carrying this string)
    unsigned long long pad;
};

// Multi-GPU: every set "serials::<expDate>::<issuer>" has ONE owner GPU
// (key_owner(exp_hour, issuer, world)), which holds its serials and its (issuer, hour) slot, and does every table
// operation on them with device-scope atomics in its own HBM.  What crosses NVLink is BULK traffic only (measured:
// fine-grained remote atomics run at ~0.4 G ops/s per GPU and made K_map 4x slower, profiles/r2_peer_atomics_2gpu.json):
// K_map appends the 64-byte key record of an entry owned elsewhere to a per-source region of the owner's INBOX
// (coalesced posted writes, cursors in local memory), the owner inserts / resolves its inbox after a barrier, and the
// source pulls the result bits back as contiguous arrays.  The peers' memory is mapped into every rank (peer access
// in one process, CUDA IPC across processes).
constexpr uint32_t kMaxWorld = 8;
constexpr uint32_t kParities = 3;   // exchange buffers in flight: the host pipeline's three stages / two rounds of the device path

// exchange areas of every rank (peer visible).  Region of (parity p, source s) = element offset (p * world + s) * X.
struct PeerExchange {
    ctmr_key* inbox[kMaxWorld];             // [kParities][world][X] key records appended by the sources
    uint8_t* out_wu[kMaxWorld];             // [kParities][world][X] was_unknown of the inbox records, same positions
    uint8_t* out_first[kMaxWorld];          // [kParities][world][X] first_issuer_hour
    unsigned long long* counts[kMaxWorld];  // [kParities][world] records per source region, published by the sources
    uint64_t X;                             // capacity of a region = the most entries one rank maps per round
    uint32_t world, rank;
};

// what K_map needs to route the keys of one round
struct RouteOut {
    ctmr_key* inbox[kMaxWorld];   // owner o's region for THIS source and this round's parity
    unsigned long long* cursor;   // local [world]: records appended per owner so far
    uint32_t* rev;                // local [world][X]: inbox position -> entry of this round (for the pull)
    uint64_t X;
    uint32_t world, rank;
};

struct PeerTables {
    KnownSlot* table[kMaxWorld];
    PairSlot* pairs[kMaxWorld];
    MetaSlot* meta[kMaxWorld];
    uint32_t world;  // 1 = single GPU: device-scope atomics, no peer traffic
    uint32_t rank;
};

// ctx-lifetime issuer registry ON THE DEVICE of rank 0: Issuer.ID digest -> dense index, find-or-insert with
// system-scope atomics, so that every rank of a group gets the same index for the same issuer without any host
// coordination (a key record carries the dense index; owner routing and the full-key compare depend on it).
struct __align__(16) IssuerRegSlot {
    unsigned long long state;  // 0 empty | 1 pending | 2 ready
    uint32_t idx;
    uint32_t pad;
    uint8_t digest[32];
    unsigned long long pad2[2];
};
static_assert(sizeof(IssuerRegSlot) == 64, "IssuerRegSlot layout");

struct IssuerRegistry {
    IssuerRegSlot* slots;          // [mask + 1]
    uint64_t mask;
    unsigned long long* counter;   // number of dense indices handed out
    uint8_t* by_index;             // [max_issuers][32]
    uint32_t max_issuers;
};

// barrier flags and histograms of every rank, for the cross-process barrier and the one-shot all-reduce
struct PeerFlags {
    unsigned long long* flags[kMaxWorld];          // [kPeerChannels][kMaxWorld] epochs, written by the peers
    unsigned long long* issuer_counts[kMaxWorld];  // [max_issuers]
    unsigned long long* status_counts[kMaxWorld];  // [CTMR_ST__COUNT]
    uint32_t world, rank;
    unsigned long long timeout_ns;                 // a barrier wait longer than this raises CTMR_E_PEER_TIMEOUT
};
constexpr uint32_t kPeerChannels = 20;

struct DeviceState {
    KnownSlot* table;          // this rank's shard (== peer.table[peer.rank])
    uint64_t table_mask;       // capacity - 1, identical on every rank
    PairSlot* pairs;
    uint64_t pair_mask;
    PeerTables peer;
    unsigned long long* issuer_counts;  // [max_issuers]
    uint32_t max_issuers;
    unsigned long long* status_counts;  // [CTMR_ST__COUNT]
    unsigned long long* slots_used;     // [1]
    int* error_flag;                    // [1]: 0 ok, CTMR_E_TABLE_FULL...
    MetaSlot* meta;                     // (issuer, kind, bytes) first-seen table
    uint64_t meta_mask;
};

struct MapParams {
    const uint8_t* blob;
    uint64_t blob_bytes;
    const uint64_t* offsets;
    const uint32_t* lens;  // optional explicit lengths (records not contiguous)
    uint64_t n;
    const uint32_t* issuer_idx;
    const uint32_t* issuer_map;
    uint32_t issuer_map_len;
    uint32_t debug_skip_walk;  // profiling aid (CTMR_DEBUG_SKIP_WALK=1): time K_map without the DER walker; results invalid
    uint64_t first_index;
    int64_t now_sec;
    uint32_t now_frac_nonzero;
    uint32_t one;  // always 1: a multiplier ptxas cannot fold (see fadd in ctmr_device.cuh)
    uint32_t rot_mul[12];  // 2^(32-n) multipliers of the IMAD.WIDE rotations (RotMul in ctmr_device.cuh)
    uint8_t* status;
    uint8_t* sha256;
    int64_t* exp_hour;
    uint32_t* serial_off;
    uint32_t* serial_len;
    ctmr_key* keys;
    uint32_t* issuer_name_off;
    uint32_t* issuer_name_len;
    uint32_t* crldp_off;
    uint32_t* crldp_len;
    unsigned long long* status_counts;
    unsigned long long* work_counter;  // [1] scratch of the dynamically scheduled map kernel
    const uint32_t* order;             // [n] length-bucketed processing order (NULL = entry order)
    // fused K_insert (slot_of != NULL): keys this rank owns go straight into its table (slot -> slot_of), keys owned
    // elsewhere are appended to the owner's inbox (slot_of = 0xFFFFFFFF)
    KnownSlot* table;
    RouteOut route;
    uint64_t table_mask;
    int* error_flag;
    uint32_t* slot_of;
    uint32_t light_prefetch;  // map_light_kernel: low 4 bits = leading 128-byte lines to prefetch per record, bit 4 = also the tail
    uint32_t pad_lp;
    FilterCfg filter;
};

cudaError_t launch_map(const MapParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_map_v1(const MapParams& p, int sm_count, cudaStream_t s);  // ctmr_map_alt.cu
cudaError_t launch_map_v3(const MapParams& p, int sm_count, cudaStream_t s);  // ctmr_map_alt.cu
cudaError_t launch_sha_ceiling(uint32_t iters, int rolled, int ctas_per_sm, int sm_count, uint32_t* sink, cudaStream_t s);
cudaError_t launch_len_order(const uint64_t* offsets, const uint32_t* lens, uint64_t n, uint64_t blob_bytes, unsigned int* hist256, uint32_t* order,
                             cudaStream_t s);
cudaError_t launch_insert(const DeviceState& st, const ctmr_key* keys, uint64_t m, uint32_t* slot_of, cudaStream_t s);
cudaError_t launch_resolve(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* slot_of,
                           uint32_t* pair_slot, uint8_t* was_unknown, cudaStream_t s);
cudaError_t launch_resolve_pairs(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* pair_slot,
                                 const uint8_t* was_unknown, uint8_t* first_issuer_hour, cudaStream_t s);
// owner side of the exchange: the same three passes over the inbox regions of one parity (record counts on the device)
cudaError_t launch_inbox_insert(const DeviceState& st, const PeerExchange& px, uint32_t parity, uint64_t max_per_region, uint32_t* in_slot,
                                cudaStream_t s);
cudaError_t launch_inbox_resolve(const DeviceState& st, const PeerExchange& px, uint32_t parity, uint64_t max_per_region,
                                 const uint32_t* in_slot, uint32_t* in_pair, cudaStream_t s);
cudaError_t launch_inbox_pairs(const DeviceState& st, const PeerExchange& px, uint32_t parity, uint64_t max_per_region,
                               const uint32_t* in_pair, cudaStream_t s);
// source side: publish the region sizes to the owners (before the barrier), pull the result bits (after the second one)
cudaError_t launch_publish_counts(const PeerExchange& px, uint32_t parity, const unsigned long long* cursor, cudaStream_t s);
cudaError_t launch_pull_bits(const PeerExchange& px, uint32_t parity, const unsigned long long* cursor, const uint32_t* rev,
                             uint64_t max_per_region, uint8_t* was_unknown, uint8_t* first_issuer_hour, cudaStream_t s);
cudaError_t launch_meta(const DeviceState& st, const uint8_t* blob, const uint64_t* offsets, const ctmr_key* keys, uint64_t m,
                        const uint8_t* was_unknown, const uint32_t* name_off, const uint32_t* name_len, const uint32_t* crl_off,
                        const uint32_t* crl_len, uint32_t* meta_slots /* [2*m] scratch */, uint8_t* first_dn, uint8_t* first_crl,
                        cudaStream_t s);
cudaError_t launch_meta_insert(const DeviceState& st, const uint8_t* blob, const uint64_t* offsets, const ctmr_key* keys, uint64_t m,
                               const uint8_t* was_unknown, const uint32_t* name_off, const uint32_t* name_len, const uint32_t* crl_off,
                               const uint32_t* crl_len, uint32_t* meta_slots, cudaStream_t s);
cudaError_t launch_meta_resolve(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* meta_slots, uint8_t* first_dn,
                                uint8_t* first_crl, cudaStream_t s);
cudaError_t launch_evict_pairs(const DeviceState& st, int64_t now_sec, cudaStream_t s);
cudaError_t launch_issuer_registry(const IssuerRegistry& reg, const uint8_t* digests, const uint8_t* ok, uint32_t n, uint32_t* idx_out,
                                   int* error_flag, cudaStream_t s);
cudaError_t launch_peer_post(const PeerFlags& pf, unsigned long long value, cudaStream_t s);
cudaError_t launch_peer_barrier(const PeerFlags& pf, uint32_t channel, unsigned long long epoch, int* error_flag, cudaStream_t s);
cudaError_t launch_hist_sum(const PeerFlags& pf, uint32_t n_slots, unsigned long long* counts_dst, unsigned long long* status_dst,
                            cudaStream_t s);
cudaError_t launch_issuer_prepare(const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint8_t* digests,
                                  uint8_t* ok, cudaStream_t s);
cudaError_t launch_table_count(const DeviceState& st, unsigned long long* out, cudaStream_t s);
cudaError_t launch_evict_count(const DeviceState& st, int64_t now_sec, unsigned long long* counters /* [2] */, cudaStream_t s);
cudaError_t launch_evict_compact(const DeviceState& st, int64_t now_sec, KnownSlot* keep, unsigned long long* cursor, cudaStream_t s);
cudaError_t launch_evict_reinsert(const DeviceState& st, const KnownSlot* keep, uint64_t n, cudaStream_t s);
cudaError_t launch_cardinality(const DeviceState& st, int32_t exp_hour, uint32_t issuer, unsigned long long* out,
                               cudaStream_t s);

// ---- CT wire-format front end (ctmr_frontend.cu, include/ctmr_frontend.h) ------------------------
#define CTMR_ISSUER_UNRESOLVED 0xFFFFFFFDu  // internal: Chain[0] present, dense index not looked up yet

struct FeParams {
    const uint8_t* text;        // device copy of the batch's characters (slack of 16 bytes on both sides)
    uint64_t text_bytes;
    const uint64_t* leaf_off;   // [n] string spans inside text
    const uint32_t* leaf_len;
    const uint64_t* extra_off;
    const uint32_t* extra_len;
    uint64_t n;
    // string s = 2*entry + (0: leaf_input, 1: extra_data)
    uint64_t* pad_size;         // [2n+1] decoded size rounded up to 16
    uint64_t* dec_off;          // [2n+1] placement in the decoded arena (exclusive scan of pad_size)
    uint32_t* dec_len;          // [2n]
    uint8_t* str_bad;           // [2n]
    uint8_t* decoded;           // arena
    // per entry
    uint8_t* entry_status;
    uint8_t* entry_type;
    uint64_t* timestamp;
    uint8_t* leaf_src;
    uint32_t* leaf_rel;
    uint32_t* leaf_len_out;
    uint64_t* leaf_abs;         // arena offset of the certificate K_map processes
    uint64_t* chain_abs;        // arena offset / length of Chain[0]
    uint32_t* chain_len;
    uint64_t* tbs_abs;          // precert entries: the leaf's TBSCertificate
    uint32_t* tbs_len;
    uint32_t* issuer_idx;
};

// Device mirror of the issuer registry keyed by the certificate BYTES: Chain[0] repeats for millions of
// entries, so it is hashed, probed and compared in full on the GPU and parsed once per distinct value.
struct IssuerCertSlot {
    uint64_t h;          // 0 = empty
    uint32_t len;
    uint32_t idx;        // dense issuer index or CTMR_ISSUER_BAD
    uint64_t arena_off;  // 16-byte aligned, zero padded to a multiple of 16
    uint64_t pad;
};
struct IssuerCertTable {
    const IssuerCertSlot* slots;
    uint64_t mask;
    const uint8_t* arena;
};

// hash of a byte string = finish(sum over its little-endian 8-byte words (zero padded) of word(w, i), length):
// a commutative sum, so that 32 lanes can each take every 32nd word; the host computes the same value.
__host__ __device__ __forceinline__ uint64_t fe_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t issuer_cert_word(uint64_t w, uint32_t i) {
    return fe_mix64(w + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1u));
}
__host__ __device__ __forceinline__ uint64_t issuer_cert_finish(uint64_t acc, uint32_t len) {
    return fe_mix64(acc ^ ((uint64_t)len << 32)) | (1ull << 63);  // never 0
}

size_t pem_scan_temp_bytes(uint64_t n_items);
cudaError_t launch_pem_encode(const uint8_t* blob, const uint64_t* offsets, const uint32_t* lens, const uint8_t* select, uint64_t n,
                              uint64_t* sizes, void* scan_temp, size_t scan_temp_bytes, uint64_t* pem_off, uint8_t* pem, uint64_t cap,
                              int* error_flag, int sm_count, cudaStream_t s);
size_t fe_scan_temp_bytes(uint64_t n_items);
cudaError_t launch_fe_decode(const FeParams& p, void* scan_temp, size_t scan_temp_bytes, int sm_count, cudaStream_t s);
cudaError_t launch_fe_frame(const FeParams& p, cudaStream_t s);
cudaError_t launch_fe_issuer(const FeParams& p, const IssuerCertTable& tab, uint64_t* pending, uint64_t pending_mask,
                             uint32_t* unknown_list, uint32_t unknown_cap, unsigned int* unknown_count, int sm_count, cudaStream_t s);
cudaError_t launch_fe_finish(const FeParams& p, const uint8_t* status, cudaStream_t s);

}  // namespace ctmr
