// ctmr_reduce.cu -- the reduce half: known-certificate table insert / resolve, (issuer, hour) first-seen,
// IssuerMetadata string identities, issuer preparation, set cardinality, multi-GPU key partition.
#include "ctmr_common.cuh"

namespace ctmr {

// ------------------------------------------------------------------------------------------------
// K_insert / K_resolve / K_pairs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) insert_kernel(DeviceState st, const ctmr_key* __restrict__ keys, uint64_t m,
                                                     uint32_t* __restrict__ slot_of) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint4* kr = reinterpret_cast<const uint4*>(keys + j);
    const uint4 q0 = kr[0], q1 = kr[1], q2 = kr[2], q3 = kr[3];
    if (q3.z == 0u) {  // not a Store-reaching entry
        slot_of[j] = 0xFFFFFFFFu;
        return;
    }
    const unsigned long long inv_idx = ~(((unsigned long long)q0.y << 32) | q0.x);
    const uint32_t body[12] = {q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y};
    slot_of[j] = known_insert(st.table, st.table_mask, st.error_flag, body, inv_idx);
}

__global__ void __launch_bounds__(256) resolve_kernel(DeviceState st, const ctmr_key* __restrict__ keys, uint64_t m,
                                                      const uint32_t* __restrict__ slot_of, uint32_t* __restrict__ pair_slot,
                                                      uint8_t* __restrict__ was_unknown) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = j < m;
    bool unknown = false;
    uint32_t issuer = 0;
    int32_t hour = 0;
    unsigned long long inv_idx = 0;
    if (in) {
        const uint32_t s = slot_of[j];
        if (s != 0xFFFFFFFFu) {
            const uint4 q0 = reinterpret_cast<const uint4*>(keys + j)[0];
            inv_idx = ~(((unsigned long long)q0.y << 32) | q0.x);
            hour = (int32_t)q0.z;
            issuer = q0.w;
            unknown = st.table[s].inv_first == inv_idx;  // I am the first sighting of this key
        }
        was_unknown[j] = unknown ? 1 : 0;
    }
    // per-issuer unique count: one atomic per distinct issuer per warp
    const uint32_t umask = __ballot_sync(0xffffffffu, unknown);
    uint32_t ps = 0xFFFFFFFFu;
    if (unknown) {
        const uint32_t peers = __match_any_sync(umask, issuer);
        if ((threadIdx.x & 31u) == (uint32_t)__ffs(peers) - 1u && issuer < st.max_issuers)
            atomicAdd(st.issuer_counts + issuer, (unsigned long long)__popc(peers));
        // (issuer, exp_hour) first-seen table: 64-bit key claims and identifies in one CAS
        const unsigned long long pk = ((((unsigned long long)issuer) << 32) | (uint32_t)hour) + 1ull;
        uint64_t pos = mix64(pk) & st.pair_mask;
        for (uint32_t probes = 0;; ++probes) {
            unsigned long long cur = ld_volatile_u64(&st.pairs[pos].key);
            if (cur == 0ull) cur = atomicCAS(&st.pairs[pos].key, 0ull, pk);
            if (cur == 0ull || cur == pk) {
                atomicMax(&st.pairs[pos].inv_first, inv_idx);
                ps = (uint32_t)pos;
                break;
            }
            pos = (pos + 1) & st.pair_mask;
            if (probes > 4096u) {
                atomicExch(st.error_flag, CTMR_E_TABLE_FULL);
                break;
            }
        }
    }
    if (in) pair_slot[j] = ps;
}

__global__ void __launch_bounds__(256) pairs_kernel(DeviceState st, const ctmr_key* __restrict__ keys, uint64_t m,
                                                    const uint32_t* __restrict__ pair_slot,
                                                    const uint8_t* __restrict__ was_unknown,
                                                    uint8_t* __restrict__ first_issuer_hour) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    uint8_t first = 0;
    const uint32_t ps = pair_slot[j];
    if (was_unknown[j] && ps != 0xFFFFFFFFu) {
        const uint4 q0 = reinterpret_cast<const uint4*>(keys + j)[0];
        const unsigned long long inv_idx = ~(((unsigned long long)q0.y << 32) | q0.x);
        first = st.pairs[ps].inv_first == inv_idx ? 1 : 0;
    }
    first_issuer_hour[j] = first;
}

// ------------------------------------------------------------------------------------------------
// K_meta: IssuerMetadata's string reducers (storage/issuermetadata.go:92-138).  For every NEW
// certificate the reference looks up the issuer DN and each CRL distribution point in per-issuer
// memo sets and only on a miss talks to Redis.  Here the (issuer, bytes) identity of the raw issuer
// Name and of the raw cRLDistributionPoints value is inserted into a device table (two independent
// 64-bit hashes; lowest index wins), so the host formats / parses those strings only for the
// first-seen candidates.  Raw-bytes identity is finer than the reference's string identity, so the
// candidate set is a superset of the reference's misses: nothing new can be hidden.
// ------------------------------------------------------------------------------------------------
__device__ inline void hash_bytes(const uint8_t* __restrict__ p, uint32_t len, uint64_t seed, uint64_t& h1, uint64_t& h2) {
    uint64_t a = seed ^ 0x9E3779B97F4A7C15ull, b = seed * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull;
    for (uint32_t i = 0; i < len; i += 8) {
        uint64_t v = 0;
        const uint32_t n = len - i < 8u ? len - i : 8u;
        for (uint32_t k = 0; k < n; ++k) v |= (uint64_t)__ldg(p + i + k) << (8 * k);
        a = mix64(a ^ v) + 0x632BE59BD9B4E019ull;
        b = mix64(b + v * 0xFF51AFD7ED558CCDull) ^ (v >> 17);
    }
    h1 = mix64(a ^ len) | 1ull;
    h2 = mix64(b + len) | 1ull;
}

__device__ inline uint32_t meta_insert(const DeviceState& st, uint64_t h1, uint64_t h2, unsigned long long inv_idx) {
    uint64_t pos = (h1 >> 5) & st.meta_mask;
    for (uint32_t probes = 0; probes < 4096u; ++probes) {
        MetaSlot* sl = st.meta + pos;
        unsigned long long c1 = ld_volatile_u64(&sl->h1);
        if (c1 == 0ull) c1 = atomicCAS(&sl->h1, 0ull, (unsigned long long)h1);
        if (c1 == 0ull || c1 == h1) {
            unsigned long long c2 = ld_volatile_u64(&sl->h2);
            if (c2 == 0ull) c2 = atomicCAS(&sl->h2, 0ull, (unsigned long long)h2);  // first writer publishes; equal strings write equal values
            if (c2 == 0ull || c2 == h2) {
                atomicMax(&sl->inv_first, inv_idx);
                return (uint32_t)pos;
            }
        }
        pos = (pos + 1) & st.meta_mask;
    }
    atomicExch(st.error_flag, CTMR_E_TABLE_FULL);
    return 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(256) meta_insert_kernel(DeviceState st, const uint8_t* __restrict__ blob,
                                                          const uint64_t* __restrict__ offsets, const ctmr_key* __restrict__ keys,
                                                          uint64_t m, const uint8_t* __restrict__ was_unknown,
                                                          const uint32_t* __restrict__ name_off, const uint32_t* __restrict__ name_len,
                                                          const uint32_t* __restrict__ crl_off, const uint32_t* __restrict__ crl_len,
                                                          uint32_t* __restrict__ meta_slots) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    uint32_t s_dn = 0xFFFFFFFFu, s_crl = 0xFFFFFFFFu;
    if (was_unknown[j]) {
        const uint4 q0 = reinterpret_cast<const uint4*>(keys + j)[0];
        const unsigned long long inv_idx = ~(((unsigned long long)q0.y << 32) | q0.x);
        const uint8_t* d = blob + offsets[j];
        uint64_t h1, h2;
        if (name_len[j]) {
            hash_bytes(d + name_off[j], name_len[j], ((uint64_t)q0.w << 8) | 1u, h1, h2);
            s_dn = meta_insert(st, h1, h2, inv_idx);
        }
        if (crl_len[j]) {
            hash_bytes(d + crl_off[j], crl_len[j], ((uint64_t)q0.w << 8) | 2u, h1, h2);
            s_crl = meta_insert(st, h1, h2, inv_idx);
        }
    }
    meta_slots[2 * j] = s_dn;
    meta_slots[2 * j + 1] = s_crl;
}

__global__ void __launch_bounds__(256) meta_resolve_kernel(DeviceState st, const ctmr_key* __restrict__ keys, uint64_t m,
                                                           const uint32_t* __restrict__ meta_slots, uint8_t* __restrict__ first_dn,
                                                           uint8_t* __restrict__ first_crl) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint4 q0 = reinterpret_cast<const uint4*>(keys + j)[0];
    const unsigned long long inv_idx = ~(((unsigned long long)q0.y << 32) | q0.x);
    const uint32_t a = meta_slots[2 * j], b = meta_slots[2 * j + 1];
    if (first_dn) first_dn[j] = (a != 0xFFFFFFFFu && st.meta[a].inv_first == inv_idx) ? 1 : 0;
    if (first_crl) first_crl[j] = (b != 0xFFFFFFFFu && st.meta[b].inv_first == inv_idx) ? 1 : 0;
}

static inline unsigned blocks_for(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

cudaError_t launch_meta(const DeviceState& st, const uint8_t* blob, const uint64_t* offsets, const ctmr_key* keys, uint64_t m,
                        const uint8_t* was_unknown, const uint32_t* name_off, const uint32_t* name_len, const uint32_t* crl_off,
                        const uint32_t* crl_len, uint32_t* meta_slots, uint8_t* first_dn, uint8_t* first_crl, cudaStream_t s) {
    if (!m) return cudaSuccess;
    meta_insert_kernel<<<blocks_for(m, 256), 256, 0, s>>>(st, blob, offsets, keys, m, was_unknown, name_off, name_len, crl_off,
                                                          crl_len, meta_slots);
    meta_resolve_kernel<<<blocks_for(m, 256), 256, 0, s>>>(st, keys, m, meta_slots, first_dn, first_crl);
    return cudaGetLastError();
}

cudaError_t launch_insert(const DeviceState& st, const ctmr_key* keys, uint64_t m, uint32_t* slot_of, cudaStream_t s) {
    if (!m) return cudaSuccess;
    insert_kernel<<<blocks_for(m, 256), 256, 0, s>>>(st, keys, m, slot_of);
    return cudaGetLastError();
}
cudaError_t launch_resolve(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* slot_of,
                           uint32_t* pair_slot, uint8_t* was_unknown, cudaStream_t s) {
    if (!m) return cudaSuccess;
    resolve_kernel<<<blocks_for(m, 256), 256, 0, s>>>(st, keys, m, slot_of, pair_slot, was_unknown);
    return cudaGetLastError();
}
cudaError_t launch_resolve_pairs(const DeviceState& st, const ctmr_key* keys, uint64_t m, const uint32_t* pair_slot,
                                 const uint8_t* was_unknown, uint8_t* first_issuer_hour, cudaStream_t s) {
    if (!m) return cudaSuccess;
    pairs_kernel<<<blocks_for(m, 256), 256, 0, s>>>(st, keys, m, pair_slot, was_unknown, first_issuer_hour);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// issuers: x509.ParseCertificate(Chain[0]) + SHA-256(RawSubjectPublicKeyInfo)
// ------------------------------------------------------------------------------------------------
__global__ void issuer_prepare_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ offsets, uint32_t n,
                                      uint8_t* __restrict__ digests, uint8_t* __restrict__ ok) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint64_t off = offsets[k], end = offsets[k + 1];
    ParsedCert pc;
    bool good = end >= off && end - off <= 0x7fffffffull && parse_cert(blob + off, (uint32_t)(end - off), pc);
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (good) sha256_global(blob + off + pc.spki_off, pc.spki_len, h);
    for (int i = 0; i < 8; ++i) {
        digests[k * 32 + 4 * i + 0] = (uint8_t)(h[i] >> 24);
        digests[k * 32 + 4 * i + 1] = (uint8_t)(h[i] >> 16);
        digests[k * 32 + 4 * i + 2] = (uint8_t)(h[i] >> 8);
        digests[k * 32 + 4 * i + 3] = (uint8_t)h[i];
    }
    ok[k] = good ? 1 : 0;
}

cudaError_t launch_issuer_prepare(const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint8_t* digests, uint8_t* ok,
                                  cudaStream_t s) {
    if (!n) return cudaSuccess;
    issuer_prepare_kernel<<<blocks_for(n, 64), 64, 0, s>>>(blob, offsets, n, digests, ok);
    return cudaGetLastError();
}

// SetCardinality("serials::<expDate>::<issuer>"): count ready slots of that set
__global__ void __launch_bounds__(256) cardinality_kernel(DeviceState st, int32_t hour, uint32_t issuer,
                                                          unsigned long long* out) {
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= st.table_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        const KnownSlot* sl = st.table + i;
        if ((sl->tag & 3ull) == 2ull && sl->body[0] == (uint32_t)hour && sl->body[1] == issuer) ++local;
    }
    local = __reduce_add_sync(0xffffffffu, (unsigned)local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

__global__ void __launch_bounds__(256) table_count_kernel(DeviceState st, unsigned long long* out) {
    unsigned int local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= st.table_mask; i += (uint64_t)gridDim.x * blockDim.x)
        local += (st.table[i].tag & 3ull) == 2ull;
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

// ---- TTL eviction: sets whose EXPIREAT time has passed vanish ---------------------------
// pass 1: count expired / surviving ready slots, take the expired ones out of the per-issuer histogram
__global__ void __launch_bounds__(256) evict_count_kernel(DeviceState st, int64_t now_sec, unsigned long long* counters /* [2]: live, expired */) {
    unsigned int live = 0, dead = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= st.table_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        const KnownSlot& sl = st.table[i];
        if ((sl.tag & 3ull) != 2ull) continue;
        const int64_t hour = (int32_t)sl.body[0];
        if (hour * 3600 <= now_sec) {
            ++dead;
            if (sl.body[1] < st.max_issuers) atomicAdd(st.issuer_counts + sl.body[1], ~0ull);  // -1
        } else {
            ++live;
        }
    }
    live = __reduce_add_sync(0xffffffffu, live);
    dead = __reduce_add_sync(0xffffffffu, dead);
    if ((threadIdx.x & 31) == 0) {
        if (live) atomicAdd(counters, (unsigned long long)live);
        if (dead) atomicAdd(counters + 1, (unsigned long long)dead);
    }
}

// pass 2: survivors, compacted (order irrelevant)
__global__ void __launch_bounds__(256) evict_compact_kernel(DeviceState st, int64_t now_sec, KnownSlot* __restrict__ keep,
                                                            unsigned long long* cursor) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= st.table_mask; i += (uint64_t)gridDim.x * blockDim.x) {
        const KnownSlot sl = st.table[i];
        if ((sl.tag & 3ull) != 2ull) continue;
        if ((int64_t)(int32_t)sl.body[0] * 3600 <= now_sec) continue;
        keep[atomicAdd(cursor, 1ull)] = sl;
    }
}

// pass 3 (after the table has been cleared): linear probing leaves no holes only if every survivor is re-inserted
__global__ void __launch_bounds__(256) evict_reinsert_kernel(DeviceState st, const KnownSlot* __restrict__ keep, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t body[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) body[k] = keep[i].body[k];
    known_insert(st.table, st.table_mask, st.error_flag, body, keep[i].inv_first);
}

cudaError_t launch_evict_count(const DeviceState& st, int64_t now_sec, unsigned long long* counters, cudaStream_t s) {
    evict_count_kernel<<<148 * 8, 256, 0, s>>>(st, now_sec, counters);
    return cudaGetLastError();
}
cudaError_t launch_evict_compact(const DeviceState& st, int64_t now_sec, KnownSlot* keep, unsigned long long* cursor, cudaStream_t s) {
    evict_compact_kernel<<<148 * 8, 256, 0, s>>>(st, now_sec, keep, cursor);
    return cudaGetLastError();
}
cudaError_t launch_evict_reinsert(const DeviceState& st, const KnownSlot* keep, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    evict_reinsert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(st, keep, n);
    return cudaGetLastError();
}

cudaError_t launch_table_count(const DeviceState& st, unsigned long long* out, cudaStream_t s) {
    table_count_kernel<<<148 * 8, 256, 0, s>>>(st, out);
    return cudaGetLastError();
}

cudaError_t launch_cardinality(const DeviceState& st, int32_t hour, uint32_t issuer, unsigned long long* out,
                               cudaStream_t s) {
    cardinality_kernel<<<148 * 8, 256, 0, s>>>(st, hour, issuer, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// multi-GPU routing helpers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) owner_count_kernel(const ctmr_key* __restrict__ keys, uint64_t n, uint32_t world,
                                                          unsigned long long* __restrict__ counts) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint32_t owner = 0;
    if (j < n) {
        const uint4 q0 = reinterpret_cast<const uint4*>(keys + j)[0];
        valid = reinterpret_cast<const uint4*>(keys + j)[3].z != 0u;
        owner = key_owner((int32_t)q0.z, q0.w, world);
    }
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const uint32_t peers = __match_any_sync(vmask, owner);
        if ((threadIdx.x & 31u) == (uint32_t)__ffs(peers) - 1u) atomicAdd(counts + owner, (unsigned long long)__popc(peers));
    }
}

__global__ void owner_scan_kernel(const unsigned long long* counts, uint32_t world, unsigned long long* cursors) {
    unsigned long long acc = 0;
    for (uint32_t w = 0; w < world; ++w) {
        cursors[w] = acc;
        acc += counts[w];
    }
}

// `limit` = 0: buckets are contiguous (cursors start at the scanned counts).  `limit` > 0: bucket w owns the fixed
// range [w*limit, (w+1)*limit) (cursors start at w*limit); a record that does not fit raises *overflow and is dropped.
__global__ void __launch_bounds__(256) owner_scatter_kernel(const ctmr_key* __restrict__ keys, uint64_t n, uint32_t world,
                                                            unsigned long long* __restrict__ cursors,
                                                            ctmr_key* __restrict__ out, uint32_t* __restrict__ src_pos,
                                                            uint64_t limit, int* __restrict__ overflow) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    uint32_t owner = 0;
    uint4 q0, q1, q2, q3;
    if (j < n) {
        const uint4* kr = reinterpret_cast<const uint4*>(keys + j);
        q0 = kr[0]; q1 = kr[1]; q2 = kr[2]; q3 = kr[3];
        valid = q3.z != 0u;
        owner = key_owner((int32_t)q0.z, q0.w, world);
    }
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const uint32_t peers = __match_any_sync(vmask, owner);
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        unsigned long long base = 0;
        if ((threadIdx.x & 31u) == leader) base = atomicAdd(cursors + owner, (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint64_t dst = base + __popc(peers & ((1u << (threadIdx.x & 31u)) - 1u));
        if (limit && dst >= (uint64_t)(owner + 1u) * limit) {
            atomicExch(overflow, 1);
            return;
        }
        uint4* o = reinterpret_cast<uint4*>(out + dst);
        o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3;
        src_pos[dst] = (uint32_t)j;
    }
}

cudaError_t launch_partition(const ctmr_key* keys, uint64_t n, uint32_t world, ctmr_key* keys_by_owner, uint32_t* src_pos,
                             unsigned long long* owner_counts, unsigned long long* cursors, cudaStream_t s) {
    cudaError_t err = cudaMemsetAsync(owner_counts, 0, sizeof(unsigned long long) * world, s);
    if (err != cudaSuccess || !n) return err;
    owner_count_kernel<<<blocks_for(n, 256), 256, 0, s>>>(keys, n, world, owner_counts);
    owner_scan_kernel<<<1, 1, 0, s>>>(owner_counts, world, cursors);
    owner_scatter_kernel<<<blocks_for(n, 256), 256, 0, s>>>(keys, n, world, cursors, keys_by_owner, src_pos, 0, nullptr);
    return cudaGetLastError();
}

__global__ void owner_fixed_init_kernel(uint32_t world, uint64_t capacity, unsigned long long* cursors) {
    for (uint32_t w = threadIdx.x; w < world; w += blockDim.x) cursors[w] = (unsigned long long)w * capacity;
}

// Fixed-capacity routing: no bucket size ever has to reach the host (the all-to-all uses equal splits).
cudaError_t launch_partition_fixed(const ctmr_key* keys, uint64_t n, uint32_t world, uint64_t capacity, ctmr_key* keys_by_owner,
                                   uint32_t* src_pos, int* overflow, unsigned long long* cursors, cudaStream_t s) {
    cudaError_t err = cudaMemsetAsync(keys_by_owner, 0, (size_t)world * capacity * sizeof(ctmr_key), s);  // valid = 0 everywhere
    if (err == cudaSuccess) err = cudaMemsetAsync(src_pos, 0xFF, (size_t)world * capacity * sizeof(uint32_t), s);
    if (err != cudaSuccess || !n) return err;
    owner_fixed_init_kernel<<<1, 64, 0, s>>>(world, capacity, cursors);
    owner_scatter_kernel<<<blocks_for(n, 256), 256, 0, s>>>(keys, n, world, cursors, keys_by_owner, src_pos, capacity, overflow);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) scatter_bits_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                           const uint32_t* __restrict__ src_pos, uint64_t m,
                                                           uint8_t* __restrict__ a_dst, uint8_t* __restrict__ b_dst) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t d = src_pos[j];
    if (d == 0xFFFFFFFFu) return;  // an unused slot of the fixed-capacity layout
    if (a_dst) a_dst[d] = a[j];
    if (b_dst) b_dst[d] = b[j];
}

cudaError_t launch_scatter_bits(const uint8_t* a, const uint8_t* b, const uint32_t* src_pos, uint64_t m, uint8_t* a_dst,
                                uint8_t* b_dst, cudaStream_t s) {
    if (!m) return cudaSuccess;
    scatter_bits_kernel<<<blocks_for(m, 256), 256, 0, s>>>(a, b, src_pos, m, a_dst, b_dst);
    return cudaGetLastError();
}
}  // namespace ctmr
