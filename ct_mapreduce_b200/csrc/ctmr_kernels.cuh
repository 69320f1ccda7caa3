// ctmr_kernels.cuh -- launch-side declarations shared by ctmr_kernels.cu and ctmr_api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctmr.h"

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

struct __align__(16) PairSlot {
    unsigned long long key;        // ((issuer << 32) | (uint32)exp_hour) + 1; 0 = empty
    unsigned long long inv_first;  // ~(lowest index among was-unknown entries of this (issuer, hour))
};

struct __align__(16) MetaSlot {      // IssuerMetadata string sets: identity = two independent 64-bit hashes
    unsigned long long h1;           // 0 = empty
    unsigned long long h2;           // 0 = not yet published
    unsigned long long inv_first;    // ~(lowest index among new certificates carrying this string)
    unsigned long long pad;
};

struct DeviceState {
    KnownSlot* table;
    uint64_t table_mask;       // capacity - 1
    PairSlot* pairs;
    uint64_t pair_mask;
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
    // fused K_insert (slot_of != NULL): the known-certificate table and where each entry's slot goes
    KnownSlot* table;
    uint64_t table_mask;
    int* error_flag;
    uint32_t* slot_of;
    FilterCfg filter;
};

cudaError_t launch_map(const MapParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_map_v1(const MapParams& p, int sm_count, cudaStream_t s);  // ctmr_map_alt.cu
cudaError_t launch_map_v3(const MapParams& p, int sm_count, cudaStream_t s);  // ctmr_map_alt.cu
cudaError_t launch_sha_ceiling(uint32_t iters, int rolled, int ctas_per_sm, int sm_count, uint32_t* sink, cudaStream_t s);
cudaError_t launch_len_order(const uint64_t* offsets, uint64_t n, uint64_t blob_bytes, unsigned int* hist256, uint32_t* order,
                             cudaStream_t s);
cudaError_t launch_insert(const DeviceState& st, const ctmr_key* keys, uint64_t m, uint32_t* slot_of, cudaStream_t s);
cudaError_t launch_resolve(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* slot_of,
                           uint32_t* pair_slot, uint8_t* was_unknown, cudaStream_t s);
cudaError_t launch_resolve_pairs(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* pair_slot,
                                 const uint8_t* was_unknown, uint8_t* first_issuer_hour, cudaStream_t s);
cudaError_t launch_meta(const DeviceState& st, const uint8_t* blob, const uint64_t* offsets, const ctmr_key* keys, uint64_t m,
                        const uint8_t* was_unknown, const uint32_t* name_off, const uint32_t* name_len, const uint32_t* crl_off,
                        const uint32_t* crl_len, uint32_t* meta_slots /* [2*m] scratch */, uint8_t* first_dn, uint8_t* first_crl,
                        cudaStream_t s);
cudaError_t launch_issuer_prepare(const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint8_t* digests,
                                  uint8_t* ok, cudaStream_t s);
cudaError_t launch_table_count(const DeviceState& st, unsigned long long* out, cudaStream_t s);
cudaError_t launch_cardinality(const DeviceState& st, int32_t exp_hour, uint32_t issuer, unsigned long long* out,
                               cudaStream_t s);
cudaError_t launch_partition(const ctmr_key* keys, uint64_t n, uint32_t world, ctmr_key* keys_by_owner,
                             uint32_t* src_pos, unsigned long long* owner_counts, unsigned long long* cursors,
                             cudaStream_t s);
cudaError_t launch_scatter_bits(const uint8_t* a, const uint8_t* b, const uint32_t* src_pos, uint64_t m, uint8_t* a_dst,
                                uint8_t* b_dst, cudaStream_t s);

}  // namespace ctmr
