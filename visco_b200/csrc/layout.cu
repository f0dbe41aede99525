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

extern "C" {

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
