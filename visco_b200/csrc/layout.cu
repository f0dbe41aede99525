// Layout kernels on either side of the hot path:
//   gather : MS column  data[row][chan][corr]  ->  per-(baseline, correlation) matrices A[b][t][chan]
//   scatter: reconstructed matrices out[b][t][chan]  ->  data[row][chan][corr]
// They replace the boolean-mask isel + [:, :, ci] slicing the reference does per matrix (visco/compress_ms.py:591-592,
// 604-608, 664) and the strided host scatter of visco/decompress_ms.py:216-232, including --correlation-optimized
// stacking (two correlations vstacked into one 2m x n matrix, compress_ms.py:608,638) and its inverse
// (unstack_vis, decompress_ms.py:95-104).
//
// One CTA handles one (baseline, time) row of the MS: the whole row (nchan x ncorr complex64, contiguous) is read once,
// fully coalesced, and each selected correlation plane is written to its own matrix — so extracting all four
// correlations costs one pass over the column instead of four 25%-efficient strided passes.
#include "common.cuh"

namespace {

// matrix index and row of the output for (baseline bl, selected correlation j, time t)
__device__ __forceinline__ void dest(int bl, int j, int t, int ncs, int stack, int m, int& b, int& tt) {
    if (stack == 1) {
        b = bl * ncs + j;
        tt = t;
    } else {  // pairs (j = 0,1), (2,3), ... are stacked vertically into one matrix of 2m rows
        b = bl * (ncs / 2) + (j >> 1);
        tt = (j & 1) * m + t;
    }
}

template <bool GATHER>
__global__ void __launch_bounds__(256)
layout_kernel(float2* __restrict__ data, const int32_t* __restrict__ row_idx, const int32_t* __restrict__ corr_sel,
              int nchan, int ncorr, int m, int ncs, int stack, float2* __restrict__ cube) {
    const int bl = blockIdx.y, t = blockIdx.x;
    const int row = row_idx[(size_t)bl * m + t];
    if (row < 0) return;  // padded entry of a ragged batch
    float2* src = data + (size_t)row * nchan * ncorr;
    const int mm = stack * m;
    for (int v = threadIdx.x; v < nchan; v += blockDim.x) {
        for (int j = 0; j < ncs; ++j) {
            const int c = corr_sel[(size_t)bl * ncs + j];  // per-baseline selection
            if ((unsigned)c >= (unsigned)ncorr) continue;  // never touch memory outside the row (vk_check_layout_indices reports it)
            int b, tt;
            dest(bl, j, t, ncs, stack, m, b, tt);
            float2* q = cube + ((size_t)b * mm + tt) * nchan + v;
            if (GATHER)
                *q = src[(size_t)v * ncorr + c];
            else
                src[(size_t)v * ncorr + c] = *q;
        }
    }
}

// ---- flags ------------------------------------------------------------------------------
// np.packbits(flags, axis=None): 8 booleans per byte, first element in the MOST significant bit (bitorder 'big'),
// tail padded with zeros (reference compress_ms.py:478-483); np.unpackbits(..., count=n) is the inverse
// (reference decompress_ms.py:240-246). One thread per output byte, 8 coalesced-enough byte reads.
__global__ void __launch_bounds__(256) packbits_kernel(const uint8_t* __restrict__ flags, size_t n, uint8_t* __restrict__ out) {
    const size_t nb = (n + 7) / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (size_t)gridDim.x * blockDim.x) {
        unsigned v = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const size_t e = i * 8 + j;
            v = (v << 1) | ((e < n && flags[e]) ? 1u : 0u);
        }
        out[i] = (uint8_t)v;
    }
}
__global__ void __launch_bounds__(256) unpackbits_kernel(const uint8_t* __restrict__ packed, size_t n, uint8_t* __restrict__ out) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        out[e] = (packed[e >> 3] >> (7 - (e & 7))) & 1u;
}
// da.where(FLAG, replacement, data) with a constant or a model column (reference compress_ms.py:530-562), in place
__global__ void __launch_bounds__(256) flag_replace_kernel(float2* __restrict__ data, const uint8_t* __restrict__ flags,
                                                           const float2* __restrict__ model, float2 value, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        if (flags[e]) data[e] = model ? model[e] : value;
}

// number of entries of row_idx outside [-1, nrow) plus entries of corr_sel outside [0, ncorr)
__global__ void __launch_bounds__(256) check_indices_kernel(const int32_t* __restrict__ row_idx, size_t nrow_idx, int nrow,
                                                            const int32_t* __restrict__ corr_sel, size_t ncorr_sel, int ncorr,
                                                            int32_t* __restrict__ bad) {
    int c = 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nrow_idx; e += (size_t)gridDim.x * blockDim.x) {
        const int r = row_idx[e];
        c += (r < -1 || r >= nrow);
    }
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < ncorr_sel; e += (size_t)gridDim.x * blockDim.x) {
        const int q = corr_sel[e];
        c += (q < 0 || q >= ncorr);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(bad, c);
}

}  // namespace

static int layout_launch(vk_context* h, bool gather, float2* data, const int32_t* row_idx, const int32_t* corr_sel,
                         int nbl, int m, int nchan, int ncorr, int ncs, int stack, float2* cube) {
    if (nbl <= 0 || m <= 0) return VK_OK;
    if (stack != 1 && stack != 2) return vk_fail(h, VK_EINVAL, "stack must be 1 or 2");
    if (stack == 2 && (ncs % 2)) return vk_fail(h, VK_EINVAL, "stacking needs an even number of selected correlations");
    if (ncs < 1 || ncs > ncorr || nchan < 1) return vk_fail(h, VK_EINVAL, "bad correlation selection");
    for (int b0 = 0; b0 < nbl; b0 += 65535) {
        const int nb = (nbl - b0) < 65535 ? (nbl - b0) : 65535;
        const dim3 grid(m, nb);
        const size_t coff = (size_t)b0 * (ncs / stack) * stack * m * nchan;
        if (gather)
            layout_kernel<true><<<grid, 256, 0, h->stream>>>(data, row_idx + (size_t)b0 * m, corr_sel + (size_t)b0 * ncs, nchan, ncorr, m, ncs,
                                                             stack, cube + coff);
        else
            layout_kernel<false><<<grid, 256, 0, h->stream>>>(data, row_idx + (size_t)b0 * m, corr_sel + (size_t)b0 * ncs, nchan, ncorr, m,
                                                              ncs, stack, cube + coff);
        VK_LAUNCH_CHECK(h);
    }
    return VK_OK;
}

static unsigned flag_grid(size_t n) {
    size_t g = (n + 255) / 256;
    return (unsigned)(g > 148 * 16 ? 148 * 16 : (g ? g : 1));
}

extern "C" {

int vk_packbits(vk_handle h, const uint8_t* flags_dev, size_t n, uint8_t* packed_dev) {
    if (!h) return VK_EINVAL;
    if (n == 0) return VK_OK;
    if (!flags_dev || !packed_dev) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    packbits_kernel<<<flag_grid((n + 7) / 8), 256, 0, h->stream>>>(flags_dev, n, packed_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_unpackbits(vk_handle h, const uint8_t* packed_dev, size_t n, uint8_t* flags_dev) {
    if (!h) return VK_EINVAL;
    if (n == 0) return VK_OK;
    if (!flags_dev || !packed_dev) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    unpackbits_kernel<<<flag_grid(n), 256, 0, h->stream>>>(packed_dev, n, flags_dev);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_flag_replace(vk_handle h, void* data_dev, const uint8_t* flags_dev, const void* model_dev, float value_re,
                    float value_im, size_t n) {
    if (!h) return VK_EINVAL;
    if (n == 0) return VK_OK;
    if (!data_dev || !flags_dev) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    flag_replace_kernel<<<flag_grid(n), 256, 0, h->stream>>>(static_cast<float2*>(data_dev), flags_dev,
                                                             static_cast<const float2*>(model_dev),
                                                             make_float2(value_re, value_im), n);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int vk_check_layout_indices(vk_handle h, const int32_t* row_idx_dev, size_t nrow_idx, int nrow,
                            const int32_t* corr_sel_dev, size_t ncorr_sel, int ncorr, int32_t* bad_host) {
    if (!h) return VK_EINVAL;
    if (!row_idx_dev || !corr_sel_dev || !bad_host) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    int32_t* cnt = reinterpret_cast<int32_t*>(h->d_scratch);
    VK_CUDA(h, cudaMemsetAsync(cnt, 0, 4, h->stream));
    check_indices_kernel<<<flag_grid(nrow_idx + ncorr_sel), 256, 0, h->stream>>>(row_idx_dev, nrow_idx, nrow, corr_sel_dev,
                                                                               ncorr_sel, ncorr, cnt);
    VK_LAUNCH_CHECK(h);
    VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 8, cnt, 4, cudaMemcpyDeviceToHost, h->stream));
    VK_CUDA(h, cudaStreamSynchronize(h->stream));
    *bad_host = h->h_poll[8];
    if (*bad_host) return vk_fail(h, VK_EINVAL, std::to_string(*bad_host) + " row / correlation indices are out of range");
    return VK_OK;
}

int vk_gather_baselines(vk_handle h, const void* data_dev, int nchan, int ncorr, const int32_t* row_idx_dev, int nbl,
                        int m, const int32_t* corr_sel_dev, int ncs, int stack, void* A_dev) {
    if (!h) return VK_EINVAL;
    if (!data_dev || !row_idx_dev || !corr_sel_dev || !A_dev) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    return layout_launch(h, true, const_cast<float2*>(static_cast<const float2*>(data_dev)), row_idx_dev, corr_sel_dev,
                         nbl, m, nchan, ncorr, ncs, stack, static_cast<float2*>(A_dev));
}

int vk_scatter_baselines(vk_handle h, const void* cube_dev, int nchan, int ncorr, const int32_t* row_idx_dev, int nbl,
                         int m, const int32_t* corr_sel_dev, int ncs, int stack, void* data_dev) {
    if (!h) return VK_EINVAL;
    if (!data_dev || !row_idx_dev || !corr_sel_dev || !cube_dev) return vk_fail(h, VK_EINVAL, "null buffer");
    VK_CUDA(h, cudaSetDevice(h->device));
    return layout_launch(h, false, static_cast<float2*>(data_dev), row_idx_dev, corr_sel_dev, nbl, m, nchan, ncorr, ncs,
                         stack, const_cast<float2*>(static_cast<const float2*>(cube_dev)));
}

}  // extern "C"
