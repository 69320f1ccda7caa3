// ctmr_synth_kernels.cu -- device side of the synthetic corpus generator (bench/test tooling).
// One thread per certificate; byte-identical to the CPU generator because both run ctmr_synth.h.
// Not part of the hot path and excluded from every timed region.
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "../../include/ctmr.h"
#include "ctmr_synth.h"

namespace {

__global__ void synth_len_kernel(ctmr_synth_cfg cfg, uint64_t first, uint64_t n, uint64_t* lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lens[i] = ctmr_synth_cert_len(&cfg, first + i);
    if (i == n) lens[i] = 0;
}

__global__ void synth_write_kernel(ctmr_synth_cfg cfg, uint64_t first, uint64_t n, const uint64_t* offsets, uint8_t* blob,
                                   uint32_t* issuer_idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ctmr_synth_plan pl;
    ctmr_synth_plan_make(&cfg, first + i, &pl);
    if (blob) ctmr_synth_cert_write(&cfg, &pl, blob + offsets[i]);
    if (issuer_idx) issuer_idx[i] = pl.issuer;
}

__global__ void synth_truth_kernel(ctmr_synth_cfg cfg, uint64_t first, uint64_t n, uint64_t* cert_id, int64_t* not_after,
                                   uint8_t* bc_mode) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ctmr_synth_plan pl;
    ctmr_synth_plan_make(&cfg, first + i, &pl);
    if (cert_id) cert_id[i] = pl.cert_id;
    if (not_after) not_after[i] = pl.not_after;
    if (bc_mode) bc_mode[i] = pl.bc_mode;
}

}  // namespace

extern "C" {

int ctmr_synth_offsets_device(const struct ctmr_synth_cfg* cfg, uint64_t first, uint64_t n, uint64_t* offsets,
                              uint64_t* total_bytes, void* stream) {
    if (!cfg || !offsets) return CTMR_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned threads = 128;
    const unsigned blocks = (unsigned)((n + 1 + threads - 1) / threads);
    synth_len_kernel<<<blocks, threads, 0, s>>>(*cfg, first, n, offsets);
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, offsets, offsets, (int64_t)(n + 1), s) != cudaSuccess) return CTMR_E_CUDA;
    if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) return CTMR_E_NOMEM;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, offsets, offsets, (int64_t)(n + 1), s);
    if (e == cudaSuccess && total_bytes)
        e = cudaMemcpyAsync(total_bytes, offsets + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp);
    return e == cudaSuccess ? CTMR_OK : CTMR_E_CUDA;
}

int ctmr_synth_write_device(const struct ctmr_synth_cfg* cfg, uint64_t first, uint64_t n, const uint64_t* offsets,
                            uint8_t* blob, uint32_t* issuer_idx, void* stream) {
    if (!cfg || (blob && !offsets)) return CTMR_E_INVALID;
    if (!n) return CTMR_OK;
    const unsigned threads = 128;
    synth_write_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*cfg, first, n, offsets,
                                                                                                  blob, issuer_idx);
    return cudaGetLastError() == cudaSuccess ? CTMR_OK : CTMR_E_CUDA;
}

int ctmr_synth_truth_device(const struct ctmr_synth_cfg* cfg, uint64_t first, uint64_t n, uint64_t* cert_id,
                            int64_t* not_after, uint8_t* bc_mode, void* stream) {
    if (!cfg) return CTMR_E_INVALID;
    if (!n) return CTMR_OK;
    const unsigned threads = 128;
    synth_truth_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*cfg, first, n, cert_id,
                                                                                                  not_after, bc_mode);
    return cudaGetLastError() == cudaSuccess ? CTMR_OK : CTMR_E_CUDA;
}

/* host-side: the n_issuers CA certificates (Chain[0] of every synthetic entry); returns total bytes,
 * writes offsets[0..n_issuers] and, when blob != NULL and large enough, the DER bytes */
uint64_t ctmr_synth_issuers_host(const struct ctmr_synth_cfg* cfg, uint64_t* offsets, uint8_t* blob, uint64_t cap) {
    uint64_t o = 0;
    for (uint32_t k = 0; k < cfg->n_issuers; ++k) {
        ctmr_synth_issuer_plan ip;
        ctmr_synth_issuer_plan_make(cfg, k, &ip);
        offsets[k] = o;
        if (blob && o + ip.total <= cap) ctmr_synth_issuer_write(cfg, &ip, blob + o);
        o += ip.total;
    }
    offsets[cfg->n_issuers] = o;
    return o;
}

}  // extern "C"
