// Batched one-sided cyclic Jacobi on complex64 vectors (sm_100a).
//
// Used for both halves of north_star item (b):
//   * Hermitian eigensolver: the vectors are the columns of the (normalised) Gram matrix G. Right rotations make
//     them mutually orthogonal, G J = V diag(lambda): vector i ends as lambda_i v_i   (ldot == ltot == r).
//   * small-matrix SVD (min(m,n) <= 64): the vectors are the rows (or columns) of A itself followed by a row of the
//     identity that accumulates the rotations (ltot = ldot + r).
// Replaces the LAPACK cgesdd call that np.linalg.svd makes under da.linalg.svd (reference visco/compress_ms.py:350).
//
// Work decomposition: the r vectors are cut into nb blocks of bsz vectors. One CTA owns a PAIR of blocks held in
// shared memory, one warp per vector pair, rotations applied with warp-wide loops and shuffle reductions. A sweep is
// nb-1 launches (round-robin tournament over block pairs; every launch runs nb/2 disjoint block pairs of every
// matrix); round 0 also rotates the pairs inside each block. When the whole matrix fits one CTA (nb == 2) the kernel
// iterates sweeps to convergence by itself.
#include "common.cuh"

namespace {

// circle-method tournament: n (even) players, round q in [0, n-1), slot p in [0, n/2)
__device__ __forceinline__ void rr_pair(int n, int q, int p, int& a, int& b) {
    if (p == 0) {
        a = n - 1;
        b = q;
    } else {
        a = (q + p) % (n - 1);
        b = (q - p + (n - 1)) % (n - 1);
    }
}

// Rotate the pair (X, Y) of shared-memory vectors so that X^H Y = 0. Returns |X^H Y|^2 / (|X|^2 |Y|^2).
__device__ __forceinline__ float rotate_pair(float2* __restrict__ X, float2* __restrict__ Y, float* nx, float* ny,
                                             int ldot, int ltot, int lane, float tol2_rot) {
    float zr = 0.f, zi = 0.f;
    for (int t = lane; t < ldot; t += 32) {
        const float2 x = X[t], y = Y[t];
        zr = fmaf(x.x, y.x, zr);
        zr = fmaf(x.y, y.y, zr);
        zi = fmaf(x.x, y.y, zi);
        zi = fmaf(-x.y, y.x, zi);
    }
    zr = warp_sum(zr);
    zi = warp_sum(zi);
    const float a = *nx, b = *ny;
    __syncwarp();  // every lane has read the cached norms before lane 0 rewrites them below
    const float zz = zr * zr + zi * zi;
    float rel2 = 0.f;
    if (a > 0.f && b > 0.f) rel2 = (zz / a) / b;
    if (rel2 > tol2_rot && zz > 0.f) {
        const float az = sqrtf(zz);
        const float tau = (b - a) / (2.f * az);
        const float t = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(fmaf(tau, tau, 1.f)));
        const float c = rsqrtf(fmaf(t, t, 1.f));
        const float s = c * t;
        const float wr = s * (zr / az), wi = s * (zi / az);  // w = s e^{i phi}
        for (int tt = lane; tt < ltot; tt += 32) {
            const float2 x = X[tt], y = Y[tt];
            float2 xn, yn;
            // x' = c x - conj(w) y ; y' = w x + c y
            xn.x = fmaf(c, x.x, -(wr * y.x + wi * y.y));
            xn.y = fmaf(c, x.y, -(wr * y.y - wi * y.x));
            yn.x = fmaf(c, y.x, wr * x.x - wi * x.y);
            yn.y = fmaf(c, y.y, wr * x.y + wi * x.x);
            X[tt] = xn;
            Y[tt] = yn;
        }
        if (lane == 0) {
            *nx = fmaxf(a - t * az, 0.f);
            *ny = fmaxf(b + t * az, 0.f);
        }
    }
    return rel2;
}

__global__ void __launch_bounds__(1024)
jacobi_pairs_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int ldot, int ltot, int r, int bsz, int nb,
                    int round, int inner_max, float tol2_rot, float tol2_stop, unsigned* __restrict__ offmax,
                    int32_t* __restrict__ done, int32_t* __restrict__ sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int npairs = nb >> 1;
    const int b = blockIdx.x / npairs;
    const int pslot = blockIdx.x - b * npairs;
    if (done[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nslots = 2 * bsz;
    float2* V = reinterpret_cast<float2*>(smem_raw);
    float* nrm = reinterpret_cast<float*>(V + (size_t)nslots * ltot);
    int* gidx = reinterpret_cast<int*>(nrm + nslots);
    __shared__ unsigned cta_max;

    int bi, bj;
    rr_pair(nb, round, pslot, bi, bj);
    float2* Wb = W + (size_t)b * mat_stride;

    for (int slot = warp; slot < nslots; slot += bsz) {
        const int g = slot < bsz ? bi * bsz + slot : bj * bsz + (slot - bsz);
        const bool valid = g < r;
        float s = 0.f;
        if (valid) {
            const float2* src = Wb + (size_t)g * ld;
            float2* dst = V + (size_t)slot * ltot;
            for (int t = lane; t < ltot; t += 32) {
                const float2 v = src[t];
                dst[t] = v;
                if (t < ldot) s = fmaf(v.x, v.x, fmaf(v.y, v.y, s));
            }
        }
        s = warp_sum(s);
        if (lane == 0) {
            nrm[slot] = s;
            gidx[slot] = valid ? g : -1;
        }
    }
    if (threadIdx.x == 0) cta_max = 0u;
    __syncthreads();

    const bool full = (round == 0);
    const int nrounds = full ? nslots - 1 : bsz;
    float launch_max = 0.f;
    int it = 0;
    bool converged = false;
    for (; it < inner_max; ++it) {
        float mymax = 0.f;
        for (int q = 0; q < nrounds; ++q) {
            int s1, s2;
            if (full) {
                rr_pair(nslots, q, warp, s1, s2);
            } else {
                s1 = warp;
                s2 = bsz + (warp + q) % bsz;
            }
            if (gidx[s1] >= 0 && gidx[s2] >= 0) {
                const float rel2 = rotate_pair(V + (size_t)s1 * ltot, V + (size_t)s2 * ltot, nrm + s1, nrm + s2, ldot,
                                               ltot, lane, tol2_rot);
                mymax = fmaxf(mymax, rel2);
            }
            __syncthreads();
        }
        launch_max = fmaxf(launch_max, mymax);
        if (inner_max > 1) {
            // in-kernel convergence test (whole matrix lives in this CTA)
            if (lane == 0) atomicMax(&cta_max, __float_as_uint(mymax));
            __syncthreads();
            const float sweep_max = __uint_as_float(cta_max);
            __syncthreads();
            if (threadIdx.x == 0) cta_max = 0u;
            // refresh the cached norms from the data once per sweep (they are updated by formula in between)
            for (int slot = warp; slot < nslots; slot += bsz) {
                if (gidx[slot] >= 0) {
                    const float2* v = V + (size_t)slot * ltot;
                    float s = 0.f;
                    for (int t = lane; t < ldot; t += 32) s = fmaf(v[t].x, v[t].x, fmaf(v[t].y, v[t].y, s));
                    s = warp_sum(s);
                    if (lane == 0) nrm[slot] = s;
                }
            }
            __syncthreads();
            if (sweep_max <= tol2_stop) {
                converged = true;
                ++it;
                break;
            }
        }
    }

    for (int slot = warp; slot < nslots; slot += bsz) {
        const int g = gidx[slot];
        if (g >= 0) {
            float2* dst = Wb + (size_t)g * ld;
            const float2* src = V + (size_t)slot * ltot;
            for (int t = lane; t < ltot; t += 32) dst[t] = src[t];
        }
    }
    if (inner_max > 1) {
        if (threadIdx.x == 0) {
            sweeps[b] = it;
            done[b] = converged ? 1 : 0;
        }
    } else {
        launch_max = warp_max(launch_max);
        if (lane == 0) atomicMax(&offmax[b], __float_as_uint(launch_max));
    }
}

__global__ void sweep_check_kernel(int B, float tol2_stop, unsigned* __restrict__ offmax, int32_t* __restrict__ done,
                                   int32_t* __restrict__ sweeps, int32_t* __restrict__ active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && !done[b]) {
        sweeps[b] += 1;
        const float v = __uint_as_float(offmax[b]);
        if (v <= tol2_stop)
            done[b] = 1;
        else
            atomicAdd(active, 1);
        offmax[b] = 0u;
    }
}

}  // namespace

JacobiPlan vk_jacobi_plan(const vk_context* h, int r, int ldot, int ltot) {
    JacobiPlan p;
    p.r = r;
    p.ldot = ldot;
    p.ltot = ltot;
    p.ld = ltot;
    const size_t per_vec = (size_t)ltot * sizeof(float2);
    if (r <= 64 && (size_t)(r + (r & 1)) * per_vec + 1024 <= VK_SMEM_BUDGET) {
        p.bsz = (r + 1) / 2;
        if (p.bsz < 1) p.bsz = 1;
        p.nb = 2;
    } else {
        int bsz = (h && h->jacobi_bsz > 0) ? h->jacobi_bsz : 16;
        while (bsz > 1 && 2 * (size_t)bsz * per_vec + 1024 > VK_SMEM_BUDGET) bsz >>= 1;
        p.bsz = bsz;
        p.nb = (r + bsz - 1) / bsz;
        if (p.nb & 1) p.nb++;
        if (p.nb < 2) p.nb = 2;
    }
    p.smem = 2 * (size_t)p.bsz * per_vec + 2 * (size_t)p.bsz * (sizeof(float) + sizeof(int));
    return p;
}

int vk_launch_jacobi(vk_context* h, float2* W, int B, const JacobiPlan& p, int32_t* sweeps_dev, int32_t* done_dev,
                     unsigned* offmax_dev, int32_t* active_dev) {
    if (B <= 0) return VK_OK;
    if (p.smem > VK_SMEM_BUDGET + 4096)
        return vk_fail(h, VK_EINVAL, "jacobi: a pair of vectors does not fit shared memory (matrix too large)");
    cudaStream_t st = h->stream;
    VK_CUDA(h, cudaFuncSetAttribute(jacobi_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    VK_CUDA(h, cudaMemsetAsync(sweeps_dev, 0, sizeof(int32_t) * B, st));
    VK_CUDA(h, cudaMemsetAsync(done_dev, 0, sizeof(int32_t) * B, st));
    VK_CUDA(h, cudaMemsetAsync(offmax_dev, 0, sizeof(unsigned) * B, st));
    // fp32 inner products of length ldot carry relative noise ~ sqrt(ldot) * 2^-24 (LAPACK xGESVJ uses the same
    // scale for its threshold): never rotate below it, and call a sweep converged when every off-diagonal it met
    // was below max(user tol, 4 * noise) — the rotations of that sweep then leave the matrix at the noise floor.
    const float noise = sqrtf((float)p.ldot) * 5.9604645e-8f;
    const float tol_stop = fmaxf(h->jacobi_tol, 4.f * noise);
    const float tol_rot = fmaxf(0.1f * h->jacobi_tol, noise);
    const float tol2_stop = tol_stop * tol_stop;
    const float tol2_rot = tol_rot * tol_rot;
    const size_t mat_stride = (size_t)p.r * p.ld;
    const int threads = 32 * p.bsz;
    const long long nblocks = (long long)B * (p.nb / 2);
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "jacobi: batch too large");
    if (p.nb == 2) {
        jacobi_pairs_kernel<<<(unsigned)nblocks, threads, p.smem, st>>>(W, mat_stride, p.ld, p.ldot, p.ltot, p.r, p.bsz,
                                                                         p.nb, 0, h->max_sweeps, tol2_rot, tol2_stop,
                                                                         offmax_dev, done_dev, sweeps_dev);
        VK_LAUNCH_CHECK(h);
        return VK_OK;
    }
    for (int sweep = 0; sweep < h->max_sweeps; ++sweep) {
        for (int round = 0; round < p.nb - 1; ++round) {
            jacobi_pairs_kernel<<<(unsigned)nblocks, threads, p.smem, st>>>(W, mat_stride, p.ld, p.ldot, p.ltot, p.r,
                                                                             p.bsz, p.nb, round, 1, tol2_rot, tol2_stop,
                                                                             offmax_dev, done_dev, sweeps_dev);
            VK_LAUNCH_CHECK(h);
        }
        VK_CUDA(h, cudaMemsetAsync(active_dev, 0, sizeof(int32_t), st));
        sweep_check_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, tol2_stop, offmax_dev, done_dev, sweeps_dev, active_dev);
        VK_LAUNCH_CHECK(h);
        const bool poll = ((sweep + 1) % (h->check_every > 0 ? h->check_every : 1) == 0) || sweep + 1 == h->max_sweeps;
        if (poll) {
            VK_CUDA(h, cudaMemcpyAsync(h->h_poll, active_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            VK_CUDA(h, cudaStreamSynchronize(st));
            if (h->h_poll[0] == 0) break;
        }
    }
    return VK_OK;
}
