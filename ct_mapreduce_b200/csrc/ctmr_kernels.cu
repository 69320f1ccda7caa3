// ctmr_kernels.cu -- hand-written sm_100a kernels of the CT-entry map/reduce hot path.
//
// K_map      lane-per-certificate map: DER walk + certIsFilteredOut + SHA-256(leaf DER).
//            Each lane streams ITS certificate through a private double-buffered shared-memory
//            slot with per-lane TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on a
//            per-warp mbarrier pair, so global loads never occupy registers or the LSU while the
//            INT pipe runs the 64-round compression.  (cmd/ct-fetch/ct-fetch.go:44-70,198-213,
//            storage/types.go:171-178,339-346; fingerprint = crypto/sha256.Sum256(cert.Raw).)
// K_insert   open-addressing find-or-insert of (exp_hour, issuer, raw serial) key records into the
//            persistent known-certificate table; lowest global entry index wins via atomicMax on
//            the complemented index (storage/knowncertificates.go:38-55 over SetInsert).
// K_resolve  was_unknown = "I am the lowest index of my key"; per-issuer unique counts
//            (Count()-sum semantics, cmd/storage-statistics/storage-statistics.go:44-53) with
//            warp-aggregated atomics; (issuer, exp_hour) first-seen table insert.
// K_pairs    first_issuer_hour bit (IssuerMetadata.Accumulate's seenExpDateBefore,
//            storage/issuermetadata.go:95-108).
// plus issuer preparation (SPKI SHA-256 = Issuer.ID digest, storage/types.go:124-130,155-159),
// set-cardinality scan, and the multi-GPU key partition / bit scatter helpers.
#include "ctmr_kernels.cuh"

#include "ctmr_device.cuh"
#include "ctmr_stream.cuh"
#include <cstdlib>

namespace ctmr {

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copy (TMA, non-tensor form)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared::cta bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }

// ------------------------------------------------------------------------------------------------
// known-certificate table
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__device__ __forceinline__ uint64_t key_hash(const uint32_t (&b)[12]) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < 12; i += 2) h = mix64(h ^ (((uint64_t)b[i + 1] << 32) | b[i])) + 0x632BE59BD9B4E019ull * (i + 1);
    return h;
}

// Find-or-insert of a 48-byte key body; lowest global index wins through atomicMax on ~index.
// Returns the slot, or 0xFFFFFFFF when the table is full (error flag set).  Shared by K_insert
// and by K_map when the insert is fused into the map kernel (single-GPU path).
__device__ __forceinline__ uint32_t known_insert(KnownSlot* __restrict__ table, uint64_t table_mask, int* error_flag,
                                                 const uint32_t (&body)[12], unsigned long long inv_idx) {
    const uint64_t h = key_hash(body);
    const unsigned long long tag_ready = (h & ~3ull) | 2ull, tag_pending = (h & ~3ull) | 1ull;
    uint64_t pos = (h >> 7) & table_mask;
    uint32_t probes = 0;
    for (;;) {
        KnownSlot* sl = table + pos;
        unsigned long long t = ld_volatile_u64(&sl->tag);
        if (t == 0ull) {
            t = atomicCAS(&sl->tag, 0ull, tag_pending);
            if (t == 0ull) {  // claimed: publish the key bytes, then flip to ready
                uint4* bp = reinterpret_cast<uint4*>(sl->body);
                bp[0] = make_uint4(body[0], body[1], body[2], body[3]);
                bp[1] = make_uint4(body[4], body[5], body[6], body[7]);
                bp[2] = make_uint4(body[8], body[9], body[10], body[11]);
                __threadfence();
                atomicExch(&sl->tag, tag_ready);
                atomicMax(&sl->inv_first, inv_idx);
                return (uint32_t)pos;
            }
        }
        if ((t & ~3ull) == (h & ~3ull)) {
            if ((t & 3ull) == 1ull) continue;  // another thread is publishing this slot: look again
            __threadfence();
            const uint4* bp = reinterpret_cast<const uint4*>(sl->body);
            const uint4 b0 = bp[0], b1 = bp[1], b2 = bp[2];
            const bool same = b0.x == body[0] && b0.y == body[1] && b0.z == body[2] && b0.w == body[3] && b1.x == body[4] &&
                              b1.y == body[5] && b1.z == body[6] && b1.w == body[7] && b2.x == body[8] && b2.y == body[9] &&
                              b2.z == body[10] && b2.w == body[11];
            if (same) {
                atomicMax(&sl->inv_first, inv_idx);
                return (uint32_t)pos;
            }
        }
        pos = (pos + 1) & table_mask;
        if (++probes > 4096u) {  // table effectively full
            atomicExch(error_flag, CTMR_E_TABLE_FULL);
            return 0xFFFFFFFFu;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K_map
// ------------------------------------------------------------------------------------------------
template <int WARPS, int CHUNK>
struct MapCfg {
    static constexpr int kSlot = CHUNK + 16;                 // +16: a record may start anywhere in a 16-byte line
    static constexpr int kWarpBytes = 2 * 32 * kSlot;        // two stages x 32 lanes
    static constexpr int kBlocksPerChunk = CHUNK / 64;
    static constexpr size_t kSmem = (size_t)WARPS * kWarpBytes + (size_t)WARPS * 2 * sizeof(uint64_t);
};

template <int WARPS, int CHUNK>
__global__ void __launch_bounds__(WARPS * 32) map_kernel(const __grid_constant__ MapParams p) {
    using Cfg = MapCfg<WARPS, CHUNK>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* my_slots = smem + (size_t)warp * Cfg::kWarpBytes + (size_t)lane * Cfg::kSlot;  // stage s at + s*32*kSlot
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * Cfg::kWarpBytes) + warp * 2;
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    const uint32_t slot0 = smem_u32(my_slots), slot1 = slot0 + 32 * Cfg::kSlot;
    if (lane == 0) {
        mbar_init(bar0, 32);
        mbar_init(bar1, 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;  // bit s: phase parity the next wait on stage s expects

    const uint64_t ngroups = (p.n + 31) >> 5;
    const bool want_sha = p.sha256 != nullptr;
    for (uint64_t g = (uint64_t)blockIdx.x * WARPS + warp; g < ngroups; g += (uint64_t)gridDim.x * WARPS) {
        const uint64_t e = g * 32 + lane;
        const bool act = e < p.n;
        uint64_t off = 0, end = 0;
        if (act) {
            off = p.offsets[e];
            end = p.offsets[e + 1];
        }
        // memory safety against malformed offset tables: an unusable record hashes as empty
        bool bad_span = !act || end < off || end > p.blob_bytes || end - off > 0x7fffffffull;
        const uint32_t L = bad_span ? 0u : (uint32_t)(end - off);
        const uint8_t* d = p.blob + off;

        // ---- streaming plan for the fingerprint
        const uint64_t addr = reinterpret_cast<uint64_t>(d);
        const uint32_t m = (uint32_t)(addr & 15u);
        const uint8_t* src_base = reinterpret_cast<const uint8_t*>(addr & ~15ull);
        const uint32_t nfull = L >> 6;
        const uint32_t nb = want_sha && act ? nfull + 1u + ((L & 63u) >= 56u ? 1u : 0u) : 0u;  // padded blocks
        const uint32_t ndata = want_sha && act ? (L + CHUNK - 1) / CHUNK : 0u;                  // chunks with data
        const uint32_t nch = (nb + Cfg::kBlocksPerChunk - 1) / Cfg::kBlocksPerChunk;
        uint32_t iters = warp_max_u32(nch);
        iters = iters < 2u ? 2u : iters;

        auto issue = [&](uint32_t c) {
            const uint32_t bar = (c & 1u) ? bar1 : bar0;
            if (c < ndata) {
                const uint32_t db = min((uint32_t)CHUNK, L - c * CHUNK);
                const uint32_t bytes = (m + db + 15u) & ~15u;
                mbar_arrive_expect_tx(bar, bytes);
                bulk_g2s((c & 1u) ? slot1 : slot0, src_base + (size_t)c * CHUNK, bytes, bar);
            } else {
                mbar_arrive(bar);
            }
        };
        if (want_sha) {  // both stages in flight while this lane walks the TLV tree
            issue(0);
            issue(1);
        }

        // ---- map: x509 field extraction + certIsFilteredOut
        ParsedCert pc;
        uint32_t status = CTMR_ST_PARSE_ERR;
        uint32_t issuer = CTMR_ISSUER_NONE;
        int64_t exp_hour = 0;
        if (act) {
            const bool ok = !bad_span && parse_cert(d, L, pc);
            if (ok) {
                status = CTMR_ST_OK;
                exp_hour = pc.not_after >= 0 ? pc.not_after / 3600 : -((-pc.not_after + 3599) / 3600);
                if ((pc.flags & (PC_BC_VALID | PC_IS_CA)) == (PC_BC_VALID | PC_IS_CA)) {
                    status = CTMR_ST_FILTER_CA;
                } else if (!p.filter.log_expired &&
                           (pc.not_after < p.now_sec || (pc.not_after == p.now_sec && p.now_frac_nonzero))) {
                    status = CTMR_ST_FILTER_EXPIRED;
                } else if (p.filter.filter_nonempty) {
                    bool skip = true;
                    const uint32_t cnl = (pc.flags & PC_HAS_CN) ? pc.cn_len : 0u;
                    for (uint32_t q = 0; q < p.filter.n_prefix && skip; ++q) {
                        const uint32_t po = p.filter.off[q], pl = p.filter.off[q + 1] - po;
                        if (pl > cnl) continue;
                        bool eq = true;
                        for (uint32_t i = 0; i < pl; ++i) {
                            if (__ldg(d + pc.cn_off + i) != p.filter.bytes[po + i]) {
                                eq = false;
                                break;
                            }
                        }
                        if (eq) skip = false;
                    }
                    if (skip) status = CTMR_ST_FILTER_CN;
                }
                if (status == CTMR_ST_OK) {
                    uint32_t k = p.issuer_idx ? p.issuer_idx[e] : CTMR_ISSUER_NONE;
                    if (k != CTMR_ISSUER_NONE && p.issuer_map) k = k < p.issuer_map_len ? p.issuer_map[k] : CTMR_ISSUER_NONE;
                    issuer = k;
                    if (k == CTMR_ISSUER_NONE) status = CTMR_ST_NO_ISSUER;
                    else if (k == CTMR_ISSUER_BAD) status = CTMR_ST_ISSUER_PARSE_ERR;
                    else if (pc.serial_len > CTMR_MAX_SERIAL) status = CTMR_ST_SERIAL_TOO_LONG;
                }
            } else {
                pc.serial_off = pc.serial_len = 0;
            }
            if (p.status) p.status[e] = (uint8_t)status;
            if (p.exp_hour) p.exp_hour[e] = exp_hour;
            if (p.serial_off) p.serial_off[e] = pc.serial_off;
            if (p.serial_len) p.serial_len[e] = pc.serial_len;
            if (p.keys) {
                // 64-byte key record, written as four 16-byte stores
                uint32_t kw[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) kw[i] = 0;
                const bool valid = status == CTMR_ST_OK;
                if (valid) {
                    kw[0] = pc.serial_len;
#pragma unroll
                    for (int i = 0; i < (int)CTMR_MAX_SERIAL; ++i) {
                        if ((uint32_t)i < pc.serial_len)
                            kw[(i + 1) >> 2] |= (uint32_t)__ldg(d + pc.serial_off + i) << (8 * ((i + 1) & 3));
                    }
                }
                const uint64_t gi = p.first_index + e;
                uint4* kr = reinterpret_cast<uint4*>(p.keys + e);
                kr[0] = make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)(int32_t)exp_hour, valid ? issuer : 0u);
                kr[1] = make_uint4(kw[0], kw[1], kw[2], kw[3]);
                kr[2] = make_uint4(kw[4], kw[5], kw[6], kw[7]);
                kr[3] = make_uint4(kw[8], kw[9], valid ? 1u : 0u, 0u);
            }
        }
        if (p.status_counts) {  // certIsFilteredOut.* / insertCTWorker.Inserted counters, one atomic per value per warp
            const uint32_t amask = __ballot_sync(0xffffffffu, act);
            if (act) {
                const uint32_t peers = __match_any_sync(amask, status);
                if ((uint32_t)lane == (uint32_t)__ffs(peers) - 1u)
                    atomicAdd(p.status_counts + status, (unsigned long long)__popc(peers));
            }
        }

        // ---- fingerprint: SHA-256 over the record, streamed chunk by chunk through shared memory
        if (want_sha) {
            Sha256State st;
            st.init();
            const uint32_t sel = 0x0123u + 0x1111u * (m & 3u);  // PRMT: big-endian word starting at byte (m & 3)
            for (uint32_t c = 0; c < iters; ++c) {
                const uint32_t s = c & 1u;
                mbar_wait(s ? bar1 : bar0, (parity >> s) & 1u);
                parity ^= 1u << s;
                if (c < nch) {
                    const uint8_t* slot = my_slots + (size_t)s * 32 * Cfg::kSlot;
#pragma unroll 1
                    for (uint32_t bb = 0; bb < (uint32_t)Cfg::kBlocksPerChunk; ++bb) {
                        const uint32_t b = c * Cfg::kBlocksPerChunk + bb;
                        if (b >= nb) break;
                        const uint32_t* sw = reinterpret_cast<const uint32_t*>(slot) + ((m + 64u * bb) >> 2);
                        uint32_t x[17], w[16];
#pragma unroll
                        for (int i = 0; i < 17; ++i) x[i] = sw[i];
#pragma unroll
                        for (int i = 0; i < 16; ++i) w[i] = __byte_perm(x[i], x[i + 1], sel);
                        if (b >= nfull) {  // trailing block(s): 0x80, zeros, 64-bit bit length
#pragma unroll
                            for (int i = 0; i < 16; ++i) w[i] = sha256_pad_word(w[i], b * 64u + 4u * i, L);
                            if (b == nb - 1u) {
                                w[14] = L >> 29;
                                w[15] = L << 3;
                            }
                        }
                        sha256_compress(st, w, p.one);
                    }
                }
                if (c + 2u < iters) issue(c + 2u);
            }
            if (act) {
                uint4* o = reinterpret_cast<uint4*>(p.sha256 + e * 32);
                o[0] = make_uint4(__byte_perm(st.h[0], 0, 0x0123), __byte_perm(st.h[1], 0, 0x0123),
                                  __byte_perm(st.h[2], 0, 0x0123), __byte_perm(st.h[3], 0, 0x0123));
                o[1] = make_uint4(__byte_perm(st.h[4], 0, 0x0123), __byte_perm(st.h[5], 0, 0x0123),
                                  __byte_perm(st.h[6], 0, 0x0123), __byte_perm(st.h[7], 0, 0x0123));
            }
        }
    }
}

template <int WARPS, int CHUNK>
static cudaError_t launch_map_t(const MapParams& p, int sm_count, int ctas_per_sm, cudaStream_t s) {
    using Cfg = MapCfg<WARPS, CHUNK>;
    auto kern = map_kernel<WARPS, CHUNK>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (err != cudaSuccess) return err;
    const uint64_t ngroups = (p.n + 31) / 32;
    uint64_t ctas = (uint64_t)sm_count * ctas_per_sm;  // persistent: a multiple of the SM count
    const uint64_t need = (ngroups + WARPS - 1) / WARPS;
    if (need < ctas) ctas = need ? need : 1;
    kern<<<(unsigned)ctas, WARPS * 32, Cfg::kSmem, s>>>(p);
    return cudaGetLastError();
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

cudaError_t launch_map_v1(const MapParams& p, int sm_count, cudaStream_t s) {
    static const int warps = env_int("CTMR_MAP_WARPS", 4);
    static const int cps = env_int("CTMR_MAP_CTAS_PER_SM", 0);
    if (warps == 8) return launch_map_t<8, 256>(p, sm_count, cps ? cps : 1, s);
    return launch_map_t<4, 256>(p, sm_count, cps ? cps : 3, s);
}

// ------------------------------------------------------------------------------------------------
// K_map_light: the map WITHOUT the whole-certificate fingerprint (CTMR_F_NO_FINGERPRINT, i.e. the
// reference's own semantics: it never hashes the leaf).  With no SHA-256 there is
// nothing to stream: the walker touches ~1/3 of a certificate's 32-byte sectors (TBS header,
// names, validity, the extension headers, the trailing signature header) and skips the key,
// the padding and the signature.  So: no shared memory, one thread per certificate, byte loads
// through L1 (the full 256 KB is available as cache), 32 resident warps per SM to hide the pointer
// chase.  The streaming kernel, which must stage every byte, tops out at ~2.0 TB/s here.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4) map_light_kernel(const __grid_constant__ MapParams p) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = e < p.n;
    uint32_t status = CTMR_ST_PARSE_ERR;
    if (act) {
        const uint64_t off = p.offsets[e], end = p.offsets[e + 1];
        const bool bad_span = end < off || end > p.blob_bytes || end - off > 0x7fffffffull;
        const uint32_t L = bad_span ? 0u : (uint32_t)(end - off);
        const uint8_t* d = p.blob + off;
        ParsedCert pc;
        uint32_t issuer = CTMR_ISSUER_NONE;
        int64_t exp_hour = 0;
        if (!bad_span && parse_cert(d, L, pc)) {
            status = CTMR_ST_OK;
            exp_hour = pc.not_after >= 0 ? pc.not_after / 3600 : -((-pc.not_after + 3599) / 3600);
            if ((pc.flags & (PC_BC_VALID | PC_IS_CA)) == (PC_BC_VALID | PC_IS_CA)) {
                status = CTMR_ST_FILTER_CA;
            } else if (!p.filter.log_expired &&
                       (pc.not_after < p.now_sec || (pc.not_after == p.now_sec && p.now_frac_nonzero))) {
                status = CTMR_ST_FILTER_EXPIRED;
            } else if (p.filter.filter_nonempty) {
                bool skip = true;
                const uint32_t cnl = (pc.flags & PC_HAS_CN) ? pc.cn_len : 0u;
                for (uint32_t q = 0; q < p.filter.n_prefix && skip; ++q) {
                    const uint32_t po = p.filter.off[q], pl = p.filter.off[q + 1] - po;
                    if (pl > cnl) continue;
                    bool eq = true;
                    for (uint32_t i = 0; i < pl; ++i)
                        if (__ldg(d + pc.cn_off + i) != p.filter.bytes[po + i]) { eq = false; break; }
                    if (eq) skip = false;
                }
                if (skip) status = CTMR_ST_FILTER_CN;
            }
            if (status == CTMR_ST_OK) {
                uint32_t k = p.issuer_idx ? p.issuer_idx[e] : CTMR_ISSUER_NONE;
                if (k != CTMR_ISSUER_NONE && p.issuer_map) k = k < p.issuer_map_len ? p.issuer_map[k] : CTMR_ISSUER_NONE;
                issuer = k;
                if (k == CTMR_ISSUER_NONE) status = CTMR_ST_NO_ISSUER;
                else if (k == CTMR_ISSUER_BAD) status = CTMR_ST_ISSUER_PARSE_ERR;
                else if (pc.serial_len > CTMR_MAX_SERIAL) status = CTMR_ST_SERIAL_TOO_LONG;
            }
        } else {
            pc.serial_off = pc.serial_len = 0;
            pc.issuer_off = pc.issuer_len = pc.crldp_off = pc.crldp_len = 0;
        }
        if (p.status) p.status[e] = (uint8_t)status;
        if (p.exp_hour) p.exp_hour[e] = exp_hour;
        if (p.serial_off) p.serial_off[e] = pc.serial_off;
        if (p.serial_len) p.serial_len[e] = pc.serial_len;
        if (p.issuer_name_off) {
            const bool okp = status != CTMR_ST_PARSE_ERR;
            p.issuer_name_off[e] = okp ? pc.issuer_off : 0u;
            p.issuer_name_len[e] = okp ? pc.issuer_len : 0u;
            p.crldp_off[e] = okp ? pc.crldp_off : 0u;
            p.crldp_len[e] = okp ? pc.crldp_len : 0u;
        }
        if (p.keys) {
            const bool valid = status == CTMR_ST_OK;
            uint32_t body[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) body[i] = 0;
            if (valid) {
                body[0] = (uint32_t)(int32_t)exp_hour;
                body[1] = issuer;
                body[2] = pc.serial_len;
#pragma unroll 1
                for (uint32_t i = 0; i < pc.serial_len; ++i) {  // {len, serial[39]} packed little-endian from byte 8
                    const uint32_t at = i + 1u;
                    const uint32_t b = __ldg(d + pc.serial_off + i) << (8u * (at & 3u));
                    switch (at >> 2) {  // static register indices
                    case 0: body[2] |= b; break; case 1: body[3] |= b; break; case 2: body[4] |= b; break;
                    case 3: body[5] |= b; break; case 4: body[6] |= b; break; case 5: body[7] |= b; break;
                    case 6: body[8] |= b; break; case 7: body[9] |= b; break; case 8: body[10] |= b; break;
                    default: body[11] |= b; break;
                    }
                }
            }
            const uint64_t gi = p.first_index + e;
            uint4* kr = reinterpret_cast<uint4*>(p.keys + e);
            kr[0] = make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), body[0], body[1]);
            kr[1] = make_uint4(body[2], body[3], body[4], body[5]);
            kr[2] = make_uint4(body[6], body[7], body[8], body[9]);
            kr[3] = make_uint4(body[10], body[11], valid ? 1u : 0u, 0u);
            if (p.slot_of) p.slot_of[e] = valid ? known_insert(p.table, p.table_mask, p.error_flag, body, ~gi) : 0xFFFFFFFFu;
        }
    }
    if (p.status_counts) {
        const uint32_t amask = __ballot_sync(0xffffffffu, act);
        if (act) {
            const uint32_t peers = __match_any_sync(amask, status);
            if ((threadIdx.x & 31u) == (uint32_t)__ffs(peers) - 1u)
                atomicAdd(p.status_counts + status, (unsigned long long)__popc(peers));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K_map v2: single-pass streaming map.  The DER walker (ctmr_stream.cuh) and the SHA-256 loop both
// eat from the same per-lane shared-memory window; global memory is touched once per byte, by
// asynchronous copies only.  Chunks overlap by kOverlap bytes so that a TLV header (and the small
// values the walker captures) never straddles a refill.
//   LOADER 0: per-lane cp.async 16-byte copies (LDGSTS) + commit/wait groups -- no barrier at all
//   LOADER 1: per-lane TMA bulk copy (UBLKCP) + per-warp mbarrier pair
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int WARPS, int CHUNK, int LOADER, int ROLLED = 0>
struct StreamCfg {
    static constexpr int kOverlap = 48;  // >= kWalkNeed, multiple of 16
    static constexpr int kSlot = kOverlap + CHUNK + 16;
    static constexpr int kWarpBytes = 2 * 32 * kSlot;
    static constexpr int kBlocksPerChunk = CHUNK / 64;
    static constexpr int kPieces = kSlot / 16;
    // +16: the walker's word-wise header read may touch the word after the last staged byte
    static constexpr size_t kSmem = (size_t)WARPS * kWarpBytes + 16 + (LOADER == 1 ? (size_t)WARPS * 2 * sizeof(uint64_t) : 0);
};

struct GlobalBytes {
    const uint8_t* d;
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const { return __ldg(d + x); }
};

template <int WARPS, int CHUNK, int LOADER, int ROLLED>
__global__ void __launch_bounds__(WARPS * 32) map_stream_kernel(const __grid_constant__ MapParams p) {
    using Cfg = StreamCfg<WARPS, CHUNK, LOADER>;
    constexpr uint32_t OV = Cfg::kOverlap;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* my_slots = smem + (size_t)warp * Cfg::kWarpBytes + (size_t)lane * Cfg::kSlot;  // stage s at + s*32*kSlot
    const uint32_t slot0 = smem_u32(my_slots), slot1 = slot0 + 32 * Cfg::kSlot;
    uint32_t bar0 = 0, bar1 = 0, parity = 0;
    if (LOADER == 1) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * Cfg::kWarpBytes + 16) + warp * 2;
        bar0 = smem_u32(&bars[0]);
        bar1 = smem_u32(&bars[1]);
        if (lane == 0) {
            mbar_init(bar0, 32);
            mbar_init(bar1, 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }

    const uint64_t ngroups = (p.n + 31) >> 5;
    const bool want_sha = p.sha256 != nullptr;
    for (uint64_t g = (uint64_t)blockIdx.x * WARPS + warp; g < ngroups; g += (uint64_t)gridDim.x * WARPS) {
        const uint64_t pos = g * 32 + lane;
        const bool act = pos < p.n;
        // length-bucketed assignment: the 32 records of a warp have (nearly) the same number of chunks
        const uint64_t e = act ? (p.order ? (uint64_t)p.order[pos] : pos) : 0;
        uint64_t off = 0, end = 0;
        if (act) {
            off = p.offsets[e];
            end = p.offsets[e + 1];
        }
        const bool bad_span = !act || end < off || end > p.blob_bytes || end - off > 0x7fffffffull;
        const uint32_t L = bad_span ? 0u : (uint32_t)(end - off);
        const uint8_t* d = p.blob + off;
        const uint64_t addr = reinterpret_cast<uint64_t>(d);
        const uint32_t m = (uint32_t)(addr & 15u);
        const uint8_t* src_base = reinterpret_cast<const uint8_t*>(addr & ~15ull);
        const uint32_t nfull = L >> 6;
        const uint32_t nb = want_sha && act ? nfull + 1u + ((L & 63u) >= 56u ? 1u : 0u) : 0u;
        const uint32_t ndata = act ? (L + CHUNK - 1) / CHUNK : 0u;
        uint32_t nch = (nb + Cfg::kBlocksPerChunk - 1) / Cfg::kBlocksPerChunk;
        nch = nch > ndata ? nch : ndata;                    // without SHA the walker alone drives the stream
        if (act && nch == 0u) nch = 1u;                     // empty record: one step so that the walker reports the error
        uint32_t iters = warp_max_u32(nch);
        iters = iters < 2u ? 2u : iters;

        // chunk c stages record bytes [c*CHUNK - OV, c*CHUNK + CHUNK) (chunk 0: [0, CHUNK)); byte x sits
        // at slot offset m + OV + x - c*CHUNK in every chunk
        auto issue = [&](uint32_t c) {
            const uint32_t slot = (c & 1u) ? slot1 : slot0;
            uint32_t bytes = 0, dst = slot;
            const uint8_t* src = src_base;
            if (c < ndata) {
                const uint32_t db = min((uint32_t)CHUNK, L - c * CHUNK);
                if (c == 0u) {
                    bytes = (m + db + 15u) & ~15u;
                    dst = slot + OV;
                } else {
                    bytes = (m + OV + db + 15u) & ~15u;
                    src = src_base + (size_t)c * CHUNK - OV;
                }
            }
            if (LOADER == 1) {
                const uint32_t bar = (c & 1u) ? bar1 : bar0;
                if (bytes) {
                    mbar_arrive_expect_tx(bar, bytes);
                    bulk_g2s(dst, src, bytes, bar);
                } else {
                    mbar_arrive(bar);
                }
            } else {
#pragma unroll 1
                for (uint32_t j = 0; j < bytes; j += 16u) cp_async16(dst + j, src + j);
                cp_async_commit();
            }
        };
        issue(0);
        issue(1);

        Walker w;
        w.init();
        uint32_t* key_words = (p.keys && act) ? reinterpret_cast<uint32_t*>(p.keys + e) + 4 : nullptr;
        const GlobalBytes far{d};
        Sha256State st;
        st.init();
        const uint32_t sel = 0x0123u + 0x1111u * (m & 3u);
        const uint32_t one = p.one;
        const RotMul rm = *reinterpret_cast<const RotMul*>(p.rot_mul);

        for (uint32_t c = 0; c < iters; ++c) {
            const uint32_t s = c & 1u;
            if (LOADER == 1) {
                mbar_wait(s ? bar1 : bar0, (parity >> s) & 1u);
                parity ^= 1u << s;
            } else {
                cp_async_wait<1>();  // everything but the newest group (chunk c+1) has landed
            }
            if (c < nch) {
                const uint8_t* slot = my_slots + (size_t)s * 32 * Cfg::kSlot;
                // ---- map: resume the TLV walk over the newly staged bytes
                if (act && w.st < W_DONE) {
                    const uint32_t avail = min((c + 1u) * CHUNK, L);
                    const SmemWindow rd{(s ? slot1 : slot0) + m + OV - c * CHUNK};
                    walk_advance(w, rd, far, bad_span ? 0u : avail, L, p.filter, key_words);
                }
                // ---- fingerprint: the chunk's 64-byte blocks
#pragma unroll 1
                for (uint32_t bb = 0; bb < (uint32_t)Cfg::kBlocksPerChunk; ++bb) {
                    const uint32_t b = c * Cfg::kBlocksPerChunk + bb;
                    if (b >= nb) break;
                    const uint32_t* sw = reinterpret_cast<const uint32_t*>(slot) + ((m + OV + 64u * bb) >> 2);
                    uint32_t x[17], wd[16];
#pragma unroll
                    for (int i = 0; i < 17; ++i) x[i] = sw[i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) wd[i] = __byte_perm(x[i], x[i + 1], sel);
                    if (b >= nfull) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) wd[i] = sha256_pad_word(wd[i], b * 64u + 4u * i, L);
                        if (b == nb - 1u) {
                            wd[14] = L >> 29;
                            wd[15] = L << 3;
                        }
                    }
                    if (ROLLED >= 2) sha256_compress_wide<ROLLED - 1>(st, wd, one, rm);
                    else if (ROLLED == 1) sha256_compress_rolled(st, wd, one);
                    else sha256_compress(st, wd, one);
                }
            }
            if (c + 2u < iters) issue(c + 2u);
            else if (LOADER == 0) cp_async_commit();  // keep "newest group = the one after chunk c+1" true at the tail
        }
        if (LOADER == 0) cp_async_wait<0>();

        // ---- certIsFilteredOut + Store preconditions, in the reference's order
        if (act) {
            uint32_t status = CTMR_ST_PARSE_ERR, issuer = CTMR_ISSUER_NONE;
            int64_t exp_hour = 0;
            uint32_t serial_off = 0, serial_len = 0;
            if (w.st == W_DONE) {
                status = CTMR_ST_OK;
                serial_off = w.serial_off;
                serial_len = w.serial_len;
                exp_hour = w.not_after >= 0 ? w.not_after / 3600 : -((-w.not_after + 3599) / 3600);
                if ((w.flags & (WF_BC_VALID | WF_IS_CA)) == (WF_BC_VALID | WF_IS_CA)) {
                    status = CTMR_ST_FILTER_CA;
                } else if (!p.filter.log_expired &&
                           (w.not_after < p.now_sec || (w.not_after == p.now_sec && p.now_frac_nonzero))) {
                    status = CTMR_ST_FILTER_EXPIRED;
                } else if (p.filter.filter_nonempty) {
                    // CommonName "" (no CN attribute) matches only an empty prefix
                    bool keep = (w.flags & WF_HAS_CN) ? (w.flags & WF_CN_MATCH) != 0 : false;
                    if (!(w.flags & WF_HAS_CN))
                        for (uint32_t q = 0; q < p.filter.n_prefix; ++q) keep |= p.filter.off[q + 1] == p.filter.off[q];
                    if (!keep) status = CTMR_ST_FILTER_CN;
                }
                if (status == CTMR_ST_OK) {
                    uint32_t k = p.issuer_idx ? p.issuer_idx[e] : CTMR_ISSUER_NONE;
                    if (k != CTMR_ISSUER_NONE && p.issuer_map) k = k < p.issuer_map_len ? p.issuer_map[k] : CTMR_ISSUER_NONE;
                    issuer = k;
                    if (k == CTMR_ISSUER_NONE) status = CTMR_ST_NO_ISSUER;
                    else if (k == CTMR_ISSUER_BAD) status = CTMR_ST_ISSUER_PARSE_ERR;
                    else if (serial_len > CTMR_MAX_SERIAL) status = CTMR_ST_SERIAL_TOO_LONG;
                }
            }
            if (p.status) p.status[e] = (uint8_t)status;
            if (p.exp_hour) p.exp_hour[e] = exp_hour;
            if (p.serial_off) p.serial_off[e] = serial_off;
            if (p.serial_len) p.serial_len[e] = serial_len;
            if (p.issuer_name_off) {  // spans of the strings IssuerMetadata.Accumulate looks at
                const bool okp = w.st == W_DONE;
                p.issuer_name_off[e] = okp ? w.name_off : 0u;
                p.issuer_name_len[e] = okp ? w.name_len : 0u;
                p.crldp_off[e] = okp ? w.crldp_off : 0u;
                p.crldp_len[e] = okp ? w.crldp_len : 0u;
            }
            if (p.keys) {
                const bool valid = status == CTMR_ST_OK;
                const uint64_t gi = p.first_index + e;
                uint4* kr = reinterpret_cast<uint4*>(p.keys + e);
                kr[0] = make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)(int32_t)exp_hour, valid ? issuer : 0u);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(kr) + 14) = make_uint2(valid ? 1u : 0u, 0u);
                if (p.slot_of) {
                    // fused K_insert (single-GPU path): the probe's random HBM accesses hide under the
                    // INT-bound SHA work of the other warps instead of costing a latency-bound pass
                    uint32_t slot = 0xFFFFFFFFu;
                    if (valid) {
                        const uint4 k1 = kr[1], k2 = kr[2];  // the serial words the walker stored (L2-resident)
                        const uint2 k3 = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(kr) + 12);
                        const uint32_t body[12] = {(uint32_t)(int32_t)exp_hour, issuer, k1.x, k1.y, k1.z, k1.w,
                                                   k2.x, k2.y, k2.z, k2.w, k3.x, k3.y};
                        slot = known_insert(p.table, p.table_mask, p.error_flag, body, ~gi);
                    }
                    p.slot_of[e] = slot;
                }
            }
            if (want_sha) {
                uint4* o = reinterpret_cast<uint4*>(p.sha256 + e * 32);
                o[0] = make_uint4(__byte_perm(st.h[0], 0, 0x0123), __byte_perm(st.h[1], 0, 0x0123),
                                  __byte_perm(st.h[2], 0, 0x0123), __byte_perm(st.h[3], 0, 0x0123));
                o[1] = make_uint4(__byte_perm(st.h[4], 0, 0x0123), __byte_perm(st.h[5], 0, 0x0123),
                                  __byte_perm(st.h[6], 0, 0x0123), __byte_perm(st.h[7], 0, 0x0123));
            }
            if (p.status_counts) {
                const uint32_t peers = __match_any_sync(__activemask(), status);
                if ((uint32_t)lane == (uint32_t)__ffs(peers) - 1u)
                    atomicAdd(p.status_counts + status, (unsigned long long)__popc(peers));
            }
        }
    }
}

template <int WARPS, int CHUNK, int LOADER, int ROLLED = 0>
static cudaError_t launch_stream_t(const MapParams& p, int sm_count, int ctas_per_sm, cudaStream_t s) {
    using Cfg = StreamCfg<WARPS, CHUNK, LOADER>;
    auto kern = map_stream_kernel<WARPS, CHUNK, LOADER, ROLLED>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (err != cudaSuccess) return err;
    const uint64_t ngroups = (p.n + 31) / 32;
    uint64_t ctas = (uint64_t)sm_count * ctas_per_sm;
    const uint64_t need = (ngroups + WARPS - 1) / WARPS;
    if (need < ctas) ctas = need ? need : 1;
    kern<<<(unsigned)ctas, WARPS * 32, Cfg::kSmem, s>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K_map v3: the streaming map with PER-LANE dynamic scheduling.
//
// In v2 a warp takes 32 consecutive records and runs until its longest record is done, so on
// mixed-size input (BASELINE configs[4]: 512 B..8 KB) lanes idle for most of the time (measured
// 294 GB/s vs 626 GB/s on uniform sizes).  Here every lane owns an endless stream of "slots"
// (one slot = one CHUNK of one record); a loader cursor runs exactly two slots ahead of the consumer
// cursor and crosses record boundaries on its own, so the first chunks of a lane's NEXT record are
// already in flight while it finishes the current one.  Records are handed out from a warp-local
// queue (ballot + prefix popcount) that is refilled in spans from one global atomic counter.
// cp.async only: completion is tracked per thread (wait_group), no barrier is shared by lanes.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kDynSpan = 128;  // records a warp takes from the global counter at a time

struct DynCert {  // what a lane needs to know about a record
    uint32_t e;     // entry index inside the batch
    uint32_t L;     // length (0 for an unusable span)
    uint64_t addr;  // address of byte 0
};

template <int WARPS, int CHUNK, int ROLLED>
__global__ void __launch_bounds__(WARPS * 32) map_dyn_kernel(const __grid_constant__ MapParams p,
                                                             unsigned long long* __restrict__ work_counter) {
    using Cfg = StreamCfg<WARPS, CHUNK, 0>;
    constexpr uint32_t OV = Cfg::kOverlap;
    constexpr uint32_t BPC = Cfg::kBlocksPerChunk;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slot_base = smem_u32(smem + (size_t)warp * Cfg::kWarpBytes + (size_t)lane * Cfg::kSlot);
    const uint32_t stage_stride = 32 * Cfg::kSlot;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool want_sha = p.sha256 != nullptr;
    const uint32_t one = p.one;

    // warp-uniform work queue
    uint64_t q_next = 0, q_end = 0;
    bool global_done = false;

    // per-lane cursors
    DynCert cur{0, 0, 0}, nxt{0, 0, 0};
    bool have_cur = false, have_nxt = false, drained = false;
    uint32_t c = 0, cur_nch = 0;                    // consumer: slot index inside cur, slots of cur
    uint32_t ld_c = 0, ld_nch = 0, ld_L = 0;        // loader: slot index inside its record, slots, length
    uint64_t ld_addr = 0;
    uint32_t lag = 0, stage_c = 0, stage_l = 0;     // slots issued - consumed; stage of next consume / next issue
    uint32_t tags = 0;                              // bit s: the slot staged in stage s belongs to a record (else: a bubble)
    Walker w;
    w.init();
    Sha256State st;
    st.init();

    auto slots_of = [&](uint32_t L) -> uint32_t {
        const uint32_t nb = want_sha ? (L >> 6) + 1u + ((L & 63u) >= 56u ? 1u : 0u) : 0u;
        uint32_t n = (nb + BPC - 1) / BPC;
        const uint32_t nd = (L + CHUNK - 1) / CHUNK;
        n = n > nd ? n : nd;
        return n < 2u ? 2u : n;  // >= 2 slots per record: the loader is never more than one record ahead
    };

    for (;;) {
        // ---- (a) hand out records to lanes whose loader has run off the end of its record
        const bool wants = !drained && !have_nxt && ld_c == ld_nch;
        const uint32_t want_mask = __ballot_sync(0xffffffffu, wants);
        if (want_mask) {
            if (q_next >= q_end && !global_done) {
                unsigned long long s0 = 0;
                if (lane == 0) s0 = atomicAdd(work_counter, (unsigned long long)kDynSpan);
                s0 = __shfl_sync(0xffffffffu, s0, 0);
                if (s0 >= p.n) {
                    global_done = true;
                } else {
                    q_next = s0;
                    q_end = s0 + kDynSpan < p.n ? s0 + kDynSpan : p.n;
                }
            }
            const uint32_t avail = (uint32_t)(q_end - q_next);
            const uint32_t rank = __popc(want_mask & lt_mask);
            if (wants) {
                if (rank < avail) {
                    const uint64_t e = q_next + rank;
                    const uint64_t off = p.offsets[e], end = p.offsets[e + 1];
                    const bool bad = end < off || end > p.blob_bytes || end - off > 0x7fffffffull;
                    nxt.e = (uint32_t)e;
                    nxt.L = bad ? 0u : (uint32_t)(end - off);
                    nxt.addr = reinterpret_cast<uint64_t>(p.blob + (bad ? 0 : off));
                    have_nxt = true;
                    ld_c = 0;
                    ld_L = nxt.L;
                    ld_nch = slots_of(nxt.L);
                    ld_addr = nxt.addr;
                } else if (global_done) {
                    drained = true;  // nothing left anywhere: this lane's loader idles from now on
                }
            }
            const uint32_t taken = (uint32_t)__popc(want_mask);
            q_next += taken < avail ? taken : avail;
        }
        if (!__ballot_sync(0xffffffffu, have_cur || have_nxt || !drained)) break;

        // ---- (b) consumer: slot `c` of the current record, once two slots are in flight behind it
        if (lag == 2u) {
            cp_async_wait<1>();
            const bool real = (tags >> stage_c) & 1u;
            if (real && !have_cur) {  // first slot of the record the loader started two slots ago: it becomes current
                cur = nxt;
                have_cur = true;
                have_nxt = false;
                cur_nch = slots_of(cur.L);
                w.init();
                st.init();
            }
            if (real) {
                const uint32_t L = cur.L;
                const uint32_t m = (uint32_t)(cur.addr & 15u);
                const uint32_t sbase = slot_base + stage_c * stage_stride;
                if (w.st < W_DONE) {
                    const uint32_t avail = min((c + 1u) * CHUNK, L);
                    const SmemWindow rd{sbase + m + OV - c * CHUNK};
                    const GlobalBytes far{reinterpret_cast<const uint8_t*>(cur.addr)};
                    uint32_t* key_words = p.keys ? reinterpret_cast<uint32_t*>(p.keys + cur.e) + 4 : nullptr;
                    walk_advance(w, rd, far, avail, L, p.filter, key_words);
                }
                if (want_sha) {
                    const uint32_t nfull = L >> 6;
                    const uint32_t nb = nfull + 1u + ((L & 63u) >= 56u ? 1u : 0u);
                    const uint32_t sel = 0x0123u + 0x1111u * (m & 3u);
#pragma unroll 1
                    for (uint32_t bb = 0; bb < BPC; ++bb) {
                        const uint32_t b = c * BPC + bb;
                        if (b >= nb) break;
                        const uint32_t wa = sbase + ((m + OV + 64u * bb) & ~3u);
                        uint32_t x[17], wd[16];
#pragma unroll
                        for (int i = 0; i < 17; ++i) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"(wa + 4u * i));
#pragma unroll
                        for (int i = 0; i < 16; ++i) wd[i] = __byte_perm(x[i], x[i + 1], sel);
                        if (b >= nfull) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) wd[i] = sha256_pad_word(wd[i], b * 64u + 4u * i, L);
                            if (b == nb - 1u) {
                                wd[14] = L >> 29;
                                wd[15] = L << 3;
                            }
                        }
                        if (ROLLED) sha256_compress_rolled(st, wd, one);
                        else sha256_compress(st, wd, one);
                    }
                }
                ++c;
                if (c == cur_nch) {
                    // ---- record finished: certIsFilteredOut + Store preconditions, outputs
                    const uint64_t e = cur.e;
                    uint32_t status = CTMR_ST_PARSE_ERR, issuer = CTMR_ISSUER_NONE;
                    int64_t exp_hour = 0;
                    uint32_t serial_off = 0, serial_len = 0;
                    if (w.st == W_DONE) {
                        status = CTMR_ST_OK;
                        serial_off = w.serial_off;
                        serial_len = w.serial_len;
                        exp_hour = w.not_after >= 0 ? w.not_after / 3600 : -((-w.not_after + 3599) / 3600);
                        if ((w.flags & (WF_BC_VALID | WF_IS_CA)) == (WF_BC_VALID | WF_IS_CA)) {
                            status = CTMR_ST_FILTER_CA;
                        } else if (!p.filter.log_expired &&
                                   (w.not_after < p.now_sec || (w.not_after == p.now_sec && p.now_frac_nonzero))) {
                            status = CTMR_ST_FILTER_EXPIRED;
                        } else if (p.filter.filter_nonempty) {
                            bool keep = (w.flags & WF_HAS_CN) ? (w.flags & WF_CN_MATCH) != 0 : false;
                            if (!(w.flags & WF_HAS_CN))
                                for (uint32_t q = 0; q < p.filter.n_prefix; ++q) keep |= p.filter.off[q + 1] == p.filter.off[q];
                            if (!keep) status = CTMR_ST_FILTER_CN;
                        }
                        if (status == CTMR_ST_OK) {
                            uint32_t k = p.issuer_idx ? p.issuer_idx[e] : CTMR_ISSUER_NONE;
                            if (k != CTMR_ISSUER_NONE && p.issuer_map) k = k < p.issuer_map_len ? p.issuer_map[k] : CTMR_ISSUER_NONE;
                            issuer = k;
                            if (k == CTMR_ISSUER_NONE) status = CTMR_ST_NO_ISSUER;
                            else if (k == CTMR_ISSUER_BAD) status = CTMR_ST_ISSUER_PARSE_ERR;
                            else if (serial_len > CTMR_MAX_SERIAL) status = CTMR_ST_SERIAL_TOO_LONG;
                        }
                    }
                    if (p.status) p.status[e] = (uint8_t)status;
                    if (p.exp_hour) p.exp_hour[e] = exp_hour;
                    if (p.serial_off) p.serial_off[e] = serial_off;
                    if (p.serial_len) p.serial_len[e] = serial_len;
                    if (p.keys) {
                        const bool valid = status == CTMR_ST_OK;
                        const uint64_t gi = p.first_index + e;
                        uint4* kr = reinterpret_cast<uint4*>(p.keys + e);
                        kr[0] = make_uint4((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)(int32_t)exp_hour, valid ? issuer : 0u);
                        *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(kr) + 14) = make_uint2(valid ? 1u : 0u, 0u);
                    }
                    if (want_sha) {
                        uint4* o = reinterpret_cast<uint4*>(p.sha256 + e * 32);
                        o[0] = make_uint4(__byte_perm(st.h[0], 0, 0x0123), __byte_perm(st.h[1], 0, 0x0123),
                                          __byte_perm(st.h[2], 0, 0x0123), __byte_perm(st.h[3], 0, 0x0123));
                        o[1] = make_uint4(__byte_perm(st.h[4], 0, 0x0123), __byte_perm(st.h[5], 0, 0x0123),
                                          __byte_perm(st.h[6], 0, 0x0123), __byte_perm(st.h[7], 0, 0x0123));
                    }
                    if (p.status_counts) {
                        const uint32_t peers = __match_any_sync(__activemask(), status);
                        if ((uint32_t)lane == (uint32_t)__ffs(peers) - 1u)
                            atomicAdd(p.status_counts + status, (unsigned long long)__popc(peers));
                    }
                    have_cur = false;
                    c = 0;
                }
            }
            stage_c ^= 1u;
            --lag;
        }

        // ---- (c) loader: one slot per iteration into the stage the consumer has just left (or the pipeline fill)
        {
            uint32_t bytes = 0, dst = slot_base + stage_l * stage_stride;
            const uint8_t* src = nullptr;
            tags &= ~(1u << stage_l);
            if (ld_c < ld_nch) {
                tags |= 1u << stage_l;
                const uint32_t nd = (ld_L + CHUNK - 1) / CHUNK;
                if (ld_c < nd) {
                    const uint32_t m = (uint32_t)(ld_addr & 15u);
                    const uint32_t db = min((uint32_t)CHUNK, ld_L - ld_c * CHUNK);
                    const uint8_t* base = reinterpret_cast<const uint8_t*>(ld_addr & ~15ull);
                    if (ld_c == 0u) {
                        bytes = (m + db + 15u) & ~15u;
                        dst += OV;
                        src = base;
                    } else {
                        bytes = (m + OV + db + 15u) & ~15u;
                        src = base + (size_t)ld_c * CHUNK - OV;
                    }
                }
                ++ld_c;
            }
#pragma unroll 1
            for (uint32_t j = 0; j < bytes; j += 16u) cp_async16(dst + j, src + j);
            cp_async_commit();
            stage_l ^= 1u;
            ++lag;
        }
    }
    cp_async_wait<0>();
}

template <int WARPS, int CHUNK, int ROLLED>
static cudaError_t launch_dyn_t(const MapParams& p, int sm_count, int ctas_per_sm, unsigned long long* counter, cudaStream_t s) {
    using Cfg = StreamCfg<WARPS, CHUNK, 0>;
    auto kern = map_dyn_kernel<WARPS, CHUNK, ROLLED>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s);
    if (err != cudaSuccess) return err;
    const uint64_t nspans = (p.n + kDynSpan - 1) / kDynSpan;
    uint64_t ctas = (uint64_t)sm_count * ctas_per_sm;
    const uint64_t need = (nspans + WARPS - 1) / WARPS;
    if (need < ctas) ctas = need ? need : 1;
    kern<<<(unsigned)ctas, WARPS * 32, Cfg::kSmem, s>>>(p, counter);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Length bucketing: a counting sort of the batch's entries by their number of streaming chunks,
// longest first.  K_map then gives each warp 32 records of (nearly) equal length, so a warp no
// longer waits for its longest record (mixed 512 B..8 KB input: 292 -> see DESIGN.md GB/s), and the
// longest records are started first (LPT order) so the grid drains evenly.  Three tiny kernels:
// histogram over 256 buckets, 256-entry scan, scatter.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kLenBuckets = 256;

__device__ __forceinline__ uint32_t len_bucket(const uint64_t* __restrict__ offsets, uint64_t e, uint64_t blob_bytes) {
    const uint64_t off = offsets[e], end = offsets[e + 1];
    const uint64_t L = (end < off || end > blob_bytes) ? 0 : end - off;
    const uint64_t chunks = (L + 127) >> 7;
    const uint32_t b = chunks >= kLenBuckets ? kLenBuckets - 1 : (uint32_t)chunks;
    return (kLenBuckets - 1) - b;  // bucket 0 = longest
}

__global__ void __launch_bounds__(256) len_hist_kernel(const uint64_t* __restrict__ offsets, uint64_t n, uint64_t blob_bytes,
                                                       unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[kLenBuckets];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[len_bucket(offsets, e, blob_bytes)], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void len_scan_kernel(unsigned int* hist_then_cursor) {
    unsigned int acc = 0;
    for (uint32_t b = 0; b < kLenBuckets; ++b) {
        const unsigned int c = hist_then_cursor[b];
        hist_then_cursor[b] = acc;
        acc += c;
    }
}

__global__ void __launch_bounds__(256) len_scatter_kernel(const uint64_t* __restrict__ offsets, uint64_t n, uint64_t blob_bytes,
                                                          unsigned int* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = e < n;
    const uint32_t b = in ? len_bucket(offsets, e, blob_bytes) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, b);
    if (in) {
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        unsigned int base = 0;
        if ((threadIdx.x & 31u) == leader) base = atomicAdd(&cursor[b], (unsigned int)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        order[base + __popc(peers & ((1u << (threadIdx.x & 31u)) - 1u))] = (uint32_t)e;
    }
}

cudaError_t launch_len_order(const uint64_t* offsets, uint64_t n, uint64_t blob_bytes, unsigned int* hist256, uint32_t* order,
                             cudaStream_t s) {
    if (!n) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(hist256, 0, kLenBuckets * sizeof(unsigned int), s);
    if (err != cudaSuccess) return err;
    const unsigned hb = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
    len_hist_kernel<<<hb < 148u * 8u ? (hb ? hb : 1u) : 148u * 8u, 256, 0, s>>>(offsets, n, blob_bytes, hist256);
    len_scan_kernel<<<1, 1, 0, s>>>(hist256);
    len_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(offsets, n, blob_bytes, hist256, order);
    return cudaGetLastError();
}

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

__global__ void __launch_bounds__(256) owner_scatter_kernel(const ctmr_key* __restrict__ keys, uint64_t n, uint32_t world,
                                                            unsigned long long* __restrict__ cursors,
                                                            ctmr_key* __restrict__ out, uint32_t* __restrict__ src_pos) {
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
    owner_scatter_kernel<<<blocks_for(n, 256), 256, 0, s>>>(keys, n, world, cursors, keys_by_owner, src_pos);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) scatter_bits_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                           const uint32_t* __restrict__ src_pos, uint64_t m,
                                                           uint8_t* __restrict__ a_dst, uint8_t* __restrict__ b_dst) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t d = src_pos[j];
    if (a_dst) a_dst[d] = a[j];
    if (b_dst) b_dst[d] = b[j];
}

cudaError_t launch_scatter_bits(const uint8_t* a, const uint8_t* b, const uint32_t* src_pos, uint64_t m, uint8_t* a_dst,
                                uint8_t* b_dst, cudaStream_t s) {
    if (!m) return cudaSuccess;
    scatter_bits_kernel<<<blocks_for(m, 256), 256, 0, s>>>(a, b, src_pos, m, a_dst, b_dst);
    return cudaGetLastError();
}

// Shape of the persistent grid.  Defaults = the measured best (DESIGN.md "K_map tuning"); the
// environment overrides exist for the A/B runs recorded under profiles/.
cudaError_t launch_map(const MapParams& p, int sm_count, cudaStream_t s) {
    if (p.n == 0) return cudaSuccess;
    static const int light = env_int("CTMR_MAP_LIGHT", 1);
    if (p.sha256 == nullptr && light) {  // no fingerprint requested: nothing to stream
        map_light_kernel<<<(unsigned)((p.n + 255) / 256), 256, 0, s>>>(p);
        return cudaGetLastError();
    }
    static const int variant = env_int("CTMR_MAP_VARIANT", 2);   // 1: v1 (global-memory walk), 2: streaming walk
    static const int loader = env_int("CTMR_MAP_LOADER", 0);     // 0: cp.async (LDGSTS), 1: TMA bulk copy
    static const int warps = env_int("CTMR_MAP_WARPS", 8);
    static const int chunk = env_int("CTMR_MAP_CHUNK", 128);
    static const int cps = env_int("CTMR_MAP_CTAS_PER_SM", 0);
    if (variant == 1) return launch_map_v1(p, sm_count, s);
    if (variant == 3) {
        if (!p.work_counter) return cudaErrorInvalidValue;
        if (chunk == 256) return launch_dyn_t<8, 256, 1>(p, sm_count, cps ? cps : 1, p.work_counter, s);
        return launch_dyn_t<8, 128, 1>(p, sm_count, cps ? cps : 2, p.work_counter, s);
    }
    if (loader == 1) {
        if (chunk == 128) return launch_stream_t<4, 128, 1>(p, sm_count, cps ? cps : 4, s);
        return launch_stream_t<4, 256, 1>(p, sm_count, cps ? cps : 2, s);
    }
    static const int rolled = env_int("CTMR_MAP_ROLLED", 1);
    if (rolled >= 2 && chunk == 128) {
        if (rolled == 2) return launch_stream_t<8, 128, 0, 2>(p, sm_count, cps ? cps : 2, s);
        if (rolled == 3) return launch_stream_t<8, 128, 0, 3>(p, sm_count, cps ? cps : 2, s);
        return launch_stream_t<8, 128, 0, 4>(p, sm_count, cps ? cps : 2, s);
    }
    if (rolled && chunk == 64) return launch_stream_t<8, 64, 0, 1>(p, sm_count, cps ? cps : 3, s);
    if (rolled && chunk == 128 && warps == 6) return launch_stream_t<6, 128, 0, 1>(p, sm_count, cps ? cps : 3, s);
    if (rolled) {
        if (chunk == 128) return launch_stream_t<8, 128, 0, 1>(p, sm_count, cps ? cps : 2, s);
        if (warps == 4) return launch_stream_t<4, 256, 0, 1>(p, sm_count, cps ? cps : 2, s);
        return launch_stream_t<8, 256, 0, 1>(p, sm_count, cps ? cps : 1, s);
    }
    if (chunk == 128) {
        if (warps == 8) return launch_stream_t<8, 128, 0>(p, sm_count, cps ? cps : 2, s);
        return launch_stream_t<4, 128, 0>(p, sm_count, cps ? cps : 4, s);
    }
    if (warps == 8) return launch_stream_t<8, 256, 0>(p, sm_count, cps ? cps : 1, s);
    return launch_stream_t<4, 256, 0>(p, sm_count, cps ? cps : 2, s);
}

}  // namespace ctmr
