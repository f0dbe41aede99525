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

// Correctly rounded 1/sqrt(x). The rotation's cosine and the phase normalisation must be UNBIASED: a systematic
// 1-ulp error in c makes every rotation shrink (or grow) its vectors and, over ~600 rotations per vector, shows up as
// a 1e-5 relative error in U S Vt (measured); the SFU approximation rsqrtf() has such a bias.
__device__ __forceinline__ float rsqrt_nr(float x) { return __frsqrt_rn(x); }

// c, w (complex sine) and t for the pair with squared norms a, b and inner product z = x^H y, |z|^2 = zz > 0
__device__ __forceinline__ void rotation_params(float a, float b, float zr, float zi, float zz, float& c, float& wr,
                                                float& wi, float& taz) {
    // fast SFU ops where only convergence speed is at stake (tau, t); exact ones where unitarity is (c, 1/|z|)
    const float rz = rsqrt_nr(zz);  // 1 / |z|
    const float az = zz * rz;       // |z|
    const float tau = 0.5f * (b - a) * rz;
    const float atau = fabsf(tau);
    const float x = fmaf(tau, tau, 1.f);
    float t = atau > 1e15f ? __fdividef(0.5f, atau) : __fdividef(1.f, atau + x * rsqrtf(x));
    t = tau >= 0.f ? t : -t;
    c = rsqrt_nr(fmaf(t, t, 1.f));
    const float s = c * t;
    wr = s * (zr * rz);
    wi = s * (zi * rz);  // w = s e^{i phi}
    taz = t * az;        // norm transfer: a' = a - t|z|, b' = b + t|z|
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
    // vectors are normalised to O(1) norms by the callers, so a * b neither overflows nor underflows harmfully
    if (a > 0.f && b > 0.f) rel2 = __fdividef(zz, a * b);
    if (rel2 > tol2_rot && zz > 0.f) {
        float c, wr, wi, taz;
        rotation_params(a, b, zr, zi, zz, c, wr, wi, taz);
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
            *nx = fmaxf(a - taz, 0.f);
            *ny = fmaxf(b + taz, 0.f);
        }
    }
    return rel2;
}

__global__ void __launch_bounds__(1024)
jacobi_pairs_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int ldot, int ltot, int r, int bsz, int nb,
                    int round, int mode, int inner_max, float tol2_rot, float tol2_stop, unsigned* __restrict__ offmax,
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

    // mode 0: cross pairs (I x J) only; 1: every pair among the 2*bsz vectors; 2: pairs inside I and inside J only
    const int nrounds = mode == 1 ? nslots - 1 : (mode == 2 ? bsz - 1 : bsz);
    const int half = bsz >> 1;
    float launch_max = 0.f;
    int it = 0;
    bool converged = false;
    for (; it < inner_max; ++it) {
        float mymax = 0.f;
        for (int q = 0; q < nrounds; ++q) {
            int s1, s2;
            if (mode == 1) {
                rr_pair(nslots, q, warp, s1, s2);
            } else if (mode == 2) {  // bsz is even here: warps [0, half) play block I, warps [half, bsz) block J
                const int base = warp < half ? 0 : bsz;
                rr_pair(bsz, q, warp < half ? warp : warp - half, s1, s2);
                s1 += base;
                s2 += base;
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

// Cross-block rotations with the I block held in REGISTERS: warp w owns vector x = I[w] for the whole launch and meets
// the bsz vectors of block J (shared memory) one per round, (w + q) mod bsz, so no two warps touch the same y. Per pair
// shared memory is read once and written once (the generic kernel above moves each vector three times). Gram path
// only (ldot == ltot == r <= 32 * EPL).
template <int EPL>
__global__ void __launch_bounds__(512, (EPL <= 8) ? 2 : 1)
jacobi_cross_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int r, int bsz, int nb, int round,
                    float tol2_rot, unsigned* __restrict__ offmax, const int32_t* __restrict__ done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = 32 * EPL;
    const int npairs = nb >> 1;
    const int b = blockIdx.x / npairs;
    const int pslot = blockIdx.x - b * npairs;
    if (done[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* Y = reinterpret_cast<float2*>(smem_raw);          // [bsz][L]
    float* nrm = reinterpret_cast<float*>(Y + (size_t)bsz * L);  // [bsz]
    int* gidx = reinterpret_cast<int*>(nrm + bsz);             // [bsz]
    int bi, bj;
    rr_pair(nb, round, pslot, bi, bj);
    float2* Wb = W + (size_t)b * mat_stride;

    const int gx = bi * bsz + warp;
    const bool validx = gx < r;
    float2 x[EPL];
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int t = lane + 32 * e;
        x[e] = (validx && t < r) ? Wb[(size_t)gx * ld + t] : make_float2(0.f, 0.f);
        a = fmaf(x[e].x, x[e].x, fmaf(x[e].y, x[e].y, a));
    }
    a = warp_sum(a);
    {
        const int gy = bj * bsz + warp;
        const bool validy = gy < r;
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int t = lane + 32 * e;
            const float2 v = (validy && t < r) ? Wb[(size_t)gy * ld + t] : make_float2(0.f, 0.f);
            Y[(size_t)warp * L + t] = v;
            s = fmaf(v.x, v.x, fmaf(v.y, v.y, s));
        }
        s = warp_sum(s);
        if (lane == 0) {
            nrm[warp] = s;
            gidx[warp] = validy ? gy : -1;
        }
    }
    __syncthreads();

    float mymax = 0.f;
    for (int q = 0; q < bsz; ++q) {
        int j = warp + q;
        if (j >= bsz) j -= bsz;
        if (validx && gidx[j] >= 0) {
            float2* Yj = Y + (size_t)j * L;
            float2 y[EPL];
            float zr = 0.f, zi = 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                y[e] = Yj[lane + 32 * e];
                zr = fmaf(x[e].x, y[e].x, zr);
                zr = fmaf(x[e].y, y[e].y, zr);
                zi = fmaf(x[e].x, y[e].y, zi);
                zi = fmaf(-x[e].y, y[e].x, zi);
            }
            zr = warp_sum(zr);
            zi = warp_sum(zi);
            const float bn = nrm[j];
            const float zz = zr * zr + zi * zi;
            float rel2 = 0.f;
            if (a > 0.f && bn > 0.f) rel2 = __fdividef(zz, a * bn);
            mymax = fmaxf(mymax, rel2);
            if (rel2 > tol2_rot && zz > 0.f) {
                float c, wr, wi, taz;
                rotation_params(a, bn, zr, zi, zz, c, wr, wi, taz);
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const float2 xo = x[e], yo = y[e];
                    float2 yn;
                    x[e].x = fmaf(c, xo.x, -(wr * yo.x + wi * yo.y));
                    x[e].y = fmaf(c, xo.y, -(wr * yo.y - wi * yo.x));
                    yn.x = fmaf(c, yo.x, wr * xo.x - wi * xo.y);
                    yn.y = fmaf(c, yo.y, wr * xo.y + wi * xo.x);
                    Yj[lane + 32 * e] = yn;
                }
                a = fmaxf(a - taz, 0.f);
                if (lane == 0) nrm[j] = fmaxf(bn + taz, 0.f);
            }
        }
        __syncthreads();
    }

    if (validx) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int t = lane + 32 * e;
            if (t < r) Wb[(size_t)gx * ld + t] = x[e];
        }
    }
    {
        const int gy = gidx[warp];
        if (gy >= 0) {
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const int t = lane + 32 * e;
                if (t < r) Wb[(size_t)gy * ld + t] = Y[(size_t)warp * L + t];
            }
        }
    }
    mymax = warp_max(mymax);
    if (lane == 0) atomicMax(&offmax[b], __float_as_uint(mymax));
}

template <int EPL>
int launch_cross(vk_context* h, float2* W, size_t mat_stride, const JacobiPlan& p, int round, unsigned nblocks,
                 float tol2_rot, unsigned* offmax, const int32_t* done) {
    const size_t smem = (size_t)p.bsz * 32 * EPL * sizeof(float2) + (size_t)p.bsz * 8;
    if (smem > 48 * 1024)  // per-device attribute; cheap enough to set on every launch
        VK_CUDA(h, cudaFuncSetAttribute(jacobi_cross_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_cross_kernel<EPL><<<nblocks, 32 * p.bsz, smem, h->stream>>>(W, mat_stride, p.ld, p.r, p.bsz, p.nb, round,
                                                                        tol2_rot, offmax, done);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int cross_epl(const JacobiPlan& p) {
    if (p.ldot != p.ltot || p.ltot != p.r || p.nb <= 2 || p.bsz != 16) return 0;
    const int need = (p.r + 31) / 32;
    const int opts[] = {3, 4, 6, 8, 12, 16};
    for (int e : opts)
        if (need <= e) return e;
    return 0;
}

int launch_cross_dispatch(vk_context* h, int epl, float2* W, size_t mat_stride, const JacobiPlan& p, int round,
                          unsigned nblocks, float tol2_rot, unsigned* offmax, const int32_t* done) {
    switch (epl) {
        case 3: return launch_cross<3>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 4: return launch_cross<4>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 6: return launch_cross<6>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 8: return launch_cross<8>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 12: return launch_cross<12>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 16: return launch_cross<16>(h, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
    }
    return vk_fail(h, VK_EINVAL, "jacobi: no cross kernel for this size");
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
                                                                         p.nb, 0, 1, h->max_sweeps, tol2_rot, tol2_stop,
                                                                         offmax_dev, done_dev, sweeps_dev);
        VK_LAUNCH_CHECK(h);
        return VK_OK;
    }
    const int epl = (h->jacobi_generic ? 0 : cross_epl(p));
    for (int sweep = 0; sweep < h->max_sweeps; ++sweep) {
        if (epl) {
            // pairs inside each block (generic kernel, intra-only), then every block pair with the register kernel
            jacobi_pairs_kernel<<<(unsigned)nblocks, threads, p.smem, st>>>(W, mat_stride, p.ld, p.ldot, p.ltot, p.r,
                                                                             p.bsz, p.nb, 0, 2, 1, tol2_rot, tol2_stop,
                                                                             offmax_dev, done_dev, sweeps_dev);
            VK_LAUNCH_CHECK(h);
            for (int round = 0; round < p.nb - 1; ++round) {
                int rc = launch_cross_dispatch(h, epl, W, mat_stride, p, round, (unsigned)nblocks, tol2_rot, offmax_dev,
                                               done_dev);
                if (rc) return rc;
            }
        } else {
            for (int round = 0; round < p.nb - 1; ++round) {
                jacobi_pairs_kernel<<<(unsigned)nblocks, threads, p.smem, st>>>(
                    W, mat_stride, p.ld, p.ldot, p.ltot, p.r, p.bsz, p.nb, round, round == 0 ? 1 : 0, 1, tol2_rot,
                    tol2_stop, offmax_dev, done_dev, sweeps_dev);
                VK_LAUNCH_CHECK(h);
            }
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
