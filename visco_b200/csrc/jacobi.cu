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

// Rotation for the pair with squared norms a, b and inner product z = x^H y (zz = |z|^2 > 0):
//   x' = c x - conj(w) y ,  y' = w x + c y ,  w = c t e^{i arg z},  t = sign(d) |z| / (|d| + sqrt(d^2 + |z|^2)),  d = (b-a)/2.
// Written so that 1/|z| is never needed:  g = 1 / (|d| + h), h = sqrt(d^2 + zz)  gives  t^2 = zz g^2,  w = z * (c g sign d)
// and the norm transfer t |z| = sign(d) zz g. The rotation is unitary iff c^2 (1 + zz g^2) = 1 for the g actually used,
// so h and g may come from the fast SFU approximations (they only steer convergence) and only c must be accurate AND
// unbiased. Late sweeps apply thousands of tiny rotations whose u = 1 + t^2 = 1 + j 2^-23 has the exact answer
// c = 1 - j 2^-24: any intermediate that is rounded first (1.0f / sqrtf(u), or a Newton residual built from a rounded
// product) turns odd j into ties, ties-to-even favours the frequent small j, and the resulting ~2e-8 systematic error per
// rotation showed up as a 1e-5 relative error in U S Vt after ~600 rotations per vector (measured; correctly rounded c
// gives 1e-6). u lies in [1, 2], so no special cases: rsqrtf + one Newton step whose residual 1 - u y^2 is evaluated
// exactly with an error-free product (two fmas), then a single final rounding.
__device__ __forceinline__ float rsqrt_unit(float u) {
    const float y = rsqrtf(u);
    const float p = u * y;
    const float pe = fmaf(u, y, -p);   // u y = p + pe exactly
    float e = fmaf(-p, y, 1.f);        // 1 - p y   (exact up to ~1e-15)
    e = fmaf(-pe, y, e);               // 1 - u y^2
    return fmaf(0.5f * y, e, y);
}
__device__ __forceinline__ void rotation_params(float a, float b, float zr, float zi, float zz, float& c, float& wr,
                                                float& wi, float& taz) {
    const float d = 0.5f * (b - a);
    const float q = fmaf(d, d, zz);
    const float h = q * rsqrtf(q);
    const float g = __fdividef(1.f, fabsf(d) + h);
    const float zg = zz * g;
    c = rsqrt_unit(fmaf(zg, g, 1.f));
    const float sg = copysignf(c * g, d);
    wr = zr * sg;
    wi = zi * sg;
    taz = copysignf(zg, d);  // norm transfer: a' = a - t|z|, b' = b + t|z|
}

// Same rotation in "fast Givens" form: returns c, 1/c and tau = w / c = z * sign(d) * g, so that the caller can apply
//   x' = c (x - conj(tau) y) ,  y' = c (y + tau x)
// and fold the common factor c into per-vector scale factors instead of multiplying every element by it.
__device__ __forceinline__ void rotation_params_fast(float a, float b, float zr, float zi, float zz, float& c, float& invc,
                                                     float& tr, float& ti, float& taz) {
    const float d = 0.5f * (b - a);
    const float q = fmaf(d, d, zz);
    const float h = q * rsqrtf(q);
    const float g = __fdividef(1.f, fabsf(d) + h);
    const float zg = zz * g;
    const float u = fmaf(zg, g, 1.f);
    c = rsqrt_unit(u);
    invc = u * c;  // sqrt(u)
    const float tg = copysignf(g, d);
    tr = zr * tg;
    ti = zi * tg;
    taz = copysignf(zg, d);
}

// sum zr and zi over the warp with 7 shuffles instead of 10: the first butterfly step leaves the zr partial sums in the
// lower half-warp and the zi partial sums in the upper one
__device__ __forceinline__ void warp_sum2(float& zr, float& zi, int lane) {
    const bool hi = lane & 16;
    const float send = hi ? zr : zi;
    float v = (hi ? zi : zr) + __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    zr = __shfl_sync(0xffffffffu, v, 0);
    zi = __shfl_sync(0xffffffffu, v, 16);
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
    warp_sum2(zr, zi, lane);
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
//
// Data layout inside the kernel is PLANAR and pair-packed: a lane owns element pairs (t, t+1), t = 2*lane + 64*p, and
// holds (re_t, re_t+1) and (im_t, im_t+1) as float2 registers; block J sits in shared memory as separate re / im
// planes. Every dot product and rotation is then a packed fma.rn.f32x2 (FFMA2) with no operand shuffling:
// 4 FFMA2 per element pair for the inner product, 12 for the rotation (the interleaved form needs 32 scalar FMAs).
template <int EPL>
__global__ void __launch_bounds__(512, (EPL <= 8) ? 2 : 1)
jacobi_cross_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int r, int bsz, int nb, int round,
                    float tol2_rot, unsigned* __restrict__ offmax, const int32_t* __restrict__ done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(EPL % 2 == 0, "element pairs");
    constexpr int L = 32 * EPL;
    constexpr int NP = EPL / 2;
    const int npairs = nb >> 1;
    const int b = blockIdx.x / npairs;
    const int pslot = blockIdx.x - b * npairs;
    if (done[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* YR = reinterpret_cast<float*>(smem_raw);   // [bsz][L]
    float* YI = YR + (size_t)bsz * L;                  // [bsz][L]
    float* nrm = YI + (size_t)bsz * L;                 // [bsz]
    int* gidx = reinterpret_cast<int*>(nrm + bsz);     // [bsz]
    float* ysc = reinterpret_cast<float*>(gidx + bsz); // [2 * bsz] scale factor of each y and its reciprocal
    int bi, bj;
    rr_pair(nb, round, pslot, bi, bj);
    float2* Wb = W + (size_t)b * mat_stride;

    const int gx = bi * bsz + warp;
    const bool validx = gx < r;
    float2 xr[NP], xi[NP];
    float a = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int t = 2 * lane + 64 * p;
        const float2 g0 = (validx && t < r) ? Wb[(size_t)gx * ld + t] : make_float2(0.f, 0.f);
        const float2 g1 = (validx && t + 1 < r) ? Wb[(size_t)gx * ld + t + 1] : make_float2(0.f, 0.f);
        xr[p] = make_float2(g0.x, g1.x);
        xi[p] = make_float2(g0.y, g1.y);
        a += g0.x * g0.x + g0.y * g0.y + g1.x * g1.x + g1.y * g1.y;
    }
    a = warp_sum(a);
    {
        const int gy = bj * bsz + warp;
        const bool validy = gy < r;
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int t = 2 * lane + 64 * p;
            const float2 g0 = (validy && t < r) ? Wb[(size_t)gy * ld + t] : make_float2(0.f, 0.f);
            const float2 g1 = (validy && t + 1 < r) ? Wb[(size_t)gy * ld + t + 1] : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(YR + (size_t)warp * L + t) = make_float2(g0.x, g1.x);
            *reinterpret_cast<float2*>(YI + (size_t)warp * L + t) = make_float2(g0.y, g1.y);
            s += g0.x * g0.x + g0.y * g0.y + g1.x * g1.x + g1.y * g1.y;
        }
        s = warp_sum(s);
        if (lane == 0) {
            nrm[warp] = s;
            gidx[warp] = validy ? gy : -1;
            ysc[2 * warp] = 1.f;
            ysc[2 * warp + 1] = 1.f;
        }
    }
    __syncthreads();

    // Fast Givens: the registers / shared memory hold x~ and y~ with x = ax x~, y = ay y~. A rotation becomes
    // x~' = x~ - conj(tau) (ay/ax) y~ ,  y~' = y~ + tau (ax/ay) x~ ,  ax' = c ax ,  ay' = c ay   (8 instead of 12 packed FMAs
    // per element pair); the scales (>= 0.707^16 within a launch) are multiplied back in when the vectors are stored.
    float ax = 1.f, rax = 1.f;
    float mymax = 0.f;
    for (int q = 0; q < bsz; ++q) {
        int j = warp + q;
        if (j >= bsz) j -= bsz;
        if (validx && gidx[j] >= 0) {
            float* yrp = YR + (size_t)j * L + 2 * lane;
            float* yip = YI + (size_t)j * L + 2 * lane;
            // y stays in registers between the inner product and the rotation. (Re-reading it from shared memory to fit
            // 2 CTAs/SM at EPL = 16 was measured: 64 registers with spills, Jacobi time 219 -> 383 ms. Rejected.)
            float2 yr[NP], yi[NP];
            float2 P = make_float2(0.f, 0.f), Q = make_float2(0.f, 0.f), R = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                yr[p] = *reinterpret_cast<const float2*>(yrp + 64 * p);
                yi[p] = *reinterpret_cast<const float2*>(yip + 64 * p);
                P = __ffma2_rn(xr[p], yr[p], P);  // z = x^H y : re = xr yr + xi yi ; im = xr yi - xi yr
                P = __ffma2_rn(xi[p], yi[p], P);
                Q = __ffma2_rn(xr[p], yi[p], Q);
                R = __ffma2_rn(xi[p], yr[p], R);
            }
            float zr = P.x + P.y, zi = (Q.x + Q.y) - (R.x + R.y);
            warp_sum2(zr, zi, lane);
            const float ay = ysc[2 * j], ray = ysc[2 * j + 1];
            const float sxy = ax * ay;  // z = x^H y = ax ay (x~^H y~)
            zr *= sxy;
            zi *= sxy;
            const float bn = nrm[j];
            const float zz = zr * zr + zi * zi;
            float rel2 = 0.f;
            if (a > 0.f && bn > 0.f) rel2 = __fdividef(zz, a * bn);
            mymax = fmaxf(mymax, rel2);
            if (rel2 > tol2_rot && zz > 0.f) {
                float c, invc, tr, ti, taz;
                rotation_params_fast(a, bn, zr, zi, zz, c, invc, tr, ti, taz);
                const float rho = ay * rax, sig = ax * ray;  // ay / ax , ax / ay
                // x~' = x~ - kappa y~ , kappa = conj(tau) rho ;  y~' = y~ + lam x~ , lam = tau sig
                const float2 nkr = make_float2(-tr * rho, -tr * rho), pki = make_float2(ti * rho, ti * rho),
                             nki = make_float2(-ti * rho, -ti * rho);
                const float2 plr = make_float2(tr * sig, tr * sig), pli = make_float2(ti * sig, ti * sig),
                             nli = make_float2(-ti * sig, -ti * sig);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float2 oxr = xr[p], oxi = xi[p], oyr = yr[p], oyi = yi[p];
                    // (kappa y)_r = rho (tr yr + ti yi) ; (kappa y)_i = rho (tr yi - ti yr)
                    xr[p] = __ffma2_rn(nkr, oyr, __ffma2_rn(nki, oyi, oxr));
                    xi[p] = __ffma2_rn(nkr, oyi, __ffma2_rn(pki, oyr, oxi));
                    // (lam x)_r = sig (tr xr - ti xi) ; (lam x)_i = sig (tr xi + ti xr)
                    *reinterpret_cast<float2*>(yrp + 64 * p) = __ffma2_rn(plr, oxr, __ffma2_rn(nli, oxi, oyr));
                    *reinterpret_cast<float2*>(yip + 64 * p) = __ffma2_rn(plr, oxi, __ffma2_rn(pli, oxr, oyi));
                }
                ax *= c;
                rax *= invc;
                a = fmaxf(a - taz, 0.f);
                if (lane == 0) {
                    nrm[j] = fmaxf(bn + taz, 0.f);
                    ysc[2 * j] = ay * c;
                    ysc[2 * j + 1] = ray * invc;
                }
            }
        }
        __syncthreads();
    }

    if (validx) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int t = 2 * lane + 64 * p;
            if (t < r) Wb[(size_t)gx * ld + t] = make_float2(ax * xr[p].x, ax * xi[p].x);
            if (t + 1 < r) Wb[(size_t)gx * ld + t + 1] = make_float2(ax * xr[p].y, ax * xi[p].y);
        }
    }
    {
        const int gy = gidx[warp];
        if (gy >= 0) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int t = 2 * lane + 64 * p;
                const float2 vr = *reinterpret_cast<const float2*>(YR + (size_t)warp * L + t);
                const float2 vi = *reinterpret_cast<const float2*>(YI + (size_t)warp * L + t);
                const float ay = ysc[2 * warp];
                if (t < r) Wb[(size_t)gy * ld + t] = make_float2(ay * vr.x, ay * vi.x);
                if (t + 1 < r) Wb[(size_t)gy * ld + t + 1] = make_float2(ay * vr.y, ay * vi.y);
            }
        }
    }
    mymax = warp_max(mymax);
    if (lane == 0) atomicMax(&offmax[b], __float_as_uint(mymax));
}

template <int EPL>
int launch_cross(vk_context* h, cudaStream_t st, float2* W, size_t mat_stride, const JacobiPlan& p, int round,
                 unsigned nblocks, float tol2_rot, unsigned* offmax, const int32_t* done) {
    const size_t smem = (size_t)p.bsz * 32 * EPL * sizeof(float2) + (size_t)p.bsz * 16;
    if (smem > 48 * 1024)  // per-device attribute; cheap enough to set on every launch
        VK_CUDA(h, cudaFuncSetAttribute(jacobi_cross_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_cross_kernel<EPL><<<nblocks, 32 * p.bsz, smem, st>>>(W, mat_stride, p.ld, p.r, p.bsz, p.nb, round,
                                                                        tol2_rot, offmax, done);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

int cross_epl(const JacobiPlan& p) {
    if (p.ldot != p.ltot || p.ltot != p.r || p.nb <= 2 || p.bsz != 16) return 0;
    const int need = (p.r + 31) / 32;
    const int opts[] = {4, 6, 8, 12, 16};
    for (int e : opts)
        if (need <= e) return e;
    return 0;
}

int launch_cross_dispatch(vk_context* h, cudaStream_t st, int epl, float2* W, size_t mat_stride, const JacobiPlan& p,
                          int round, unsigned nblocks, float tol2_rot, unsigned* offmax, const int32_t* done) {
    switch (epl) {
        case 4: return launch_cross<4>(h, st, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 6: return launch_cross<6>(h, st, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 8: return launch_cross<8>(h, st, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 12: return launch_cross<12>(h, st, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
        case 16: return launch_cross<16>(h, st, W, mat_stride, p, round, nblocks, tol2_rot, offmax, done);
    }
    return vk_fail(h, VK_EINVAL, "jacobi: no cross kernel for this size");
}

// Whole-problem-in-one-CTA Jacobi for small vector sets (r <= 64: the small-matrix SVD path and small Gram matrices).
// A pair is handled by a GROUP of LPP lanes (LPP = 4..32), so a warp rotates 32/LPP pairs at once and the scalar
// rotation parameters, the reductions and the loop overhead are shared by them; with one warp per pair and only four
// elements per lane (64 x 64 matrices) that overhead was ~95 % of the instruction stream. Vectors live in shared
// memory as separate re / im planes (rows padded by 8 floats to spread the groups over the banks); inner products and
// rotations use packed fma.rn.f32x2 on element pairs. Sweeps iterate inside the kernel until convergence.
template <int LPP>
__global__ void __launch_bounds__(1024)
jacobi_small_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int ldot, int ltot, int r, int nslots, int lpad,
                    int max_sweeps, float tol2_rot, float tol2_stop, int32_t* __restrict__ done,
                    int32_t* __restrict__ sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GPW = 32 / LPP;
    float* XR = reinterpret_cast<float*>(smem_raw);      // [nslots][lpad]
    float* XI = XR + (size_t)nslots * lpad;               // [nslots][lpad]
    float* nrm = XI + (size_t)nslots * lpad;              // [nslots]
    __shared__ unsigned cta_max;
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane % LPP;
    const int pslot = warp * GPW + lane / LPP;
    const int npairs = nslots >> 1;
    const bool act = pslot < npairs;
    float2* Wb = W + (size_t)b * mat_stride;

    for (int v = warp; v < nslots; v += nwarps) {
        float s = 0.f;
        for (int t = lane; t < lpad; t += 32) {
            float2 w = make_float2(0.f, 0.f);
            if (v < r && t < ltot) w = Wb[(size_t)v * ld + t];
            XR[(size_t)v * lpad + t] = w.x;
            XI[(size_t)v * lpad + t] = w.y;
            if (t < ldot) s = fmaf(w.x, w.x, fmaf(w.y, w.y, s));
        }
        s = warp_sum(s);
        if (lane == 0) nrm[v] = s;
    }
    if (threadIdx.x == 0) cta_max = 0u;
    __syncthreads();

    const int nd = act ? ldot : 0, nt = act ? ltot : 0;
    int it = 0;
    bool converged = false;
    for (; it < max_sweeps; ++it) {
        float mymax = 0.f;
        for (int q = 0; q < nslots - 1; ++q) {
            int s1 = 0, s2 = 1;
            if (act) rr_pair(nslots, q, pslot, s1, s2);
            float* xr = XR + (size_t)s1 * lpad;
            float* xi = XI + (size_t)s1 * lpad;
            float* yr = XR + (size_t)s2 * lpad;
            float* yi = XI + (size_t)s2 * lpad;
            float2 P = make_float2(0.f, 0.f), Q = make_float2(0.f, 0.f), R = make_float2(0.f, 0.f);
            for (int e = 2 * g; e < nd; e += 2 * LPP) {
                float2 a_r = *reinterpret_cast<const float2*>(xr + e), a_i = *reinterpret_cast<const float2*>(xi + e);
                float2 b_r = *reinterpret_cast<const float2*>(yr + e), b_i = *reinterpret_cast<const float2*>(yi + e);
                if (e + 1 >= nd) {  // odd ldot: the second element of the pair is not part of the inner product
                    a_r.y = 0.f;
                    a_i.y = 0.f;
                }
                P = __ffma2_rn(a_r, b_r, P);
                P = __ffma2_rn(a_i, b_i, P);
                Q = __ffma2_rn(a_r, b_i, Q);
                R = __ffma2_rn(a_i, b_r, R);
            }
            float zr = P.x + P.y, zi = (Q.x + Q.y) - (R.x + R.y);
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) {
                zr += __shfl_xor_sync(0xffffffffu, zr, o);
                zi += __shfl_xor_sync(0xffffffffu, zi, o);
            }
            const float an = nrm[s1], bn = nrm[s2];
            __syncwarp();
            const float zz = zr * zr + zi * zi;
            float rel2 = 0.f;
            if (act && an > 0.f && bn > 0.f) rel2 = __fdividef(zz, an * bn);
            mymax = fmaxf(mymax, rel2);
            if (rel2 > tol2_rot && zz > 0.f) {
                float c, wr, wi, taz;
                rotation_params(an, bn, zr, zi, zz, c, wr, wi, taz);
                const float2 cc = make_float2(c, c), pwr = make_float2(wr, wr), nwr = make_float2(-wr, -wr),
                             pwi = make_float2(wi, wi), nwi = make_float2(-wi, -wi);
                for (int e = 2 * g; e < nt; e += 2 * LPP) {
                    const float2 a_r = *reinterpret_cast<const float2*>(xr + e), a_i = *reinterpret_cast<const float2*>(xi + e);
                    const float2 b_r = *reinterpret_cast<const float2*>(yr + e), b_i = *reinterpret_cast<const float2*>(yi + e);
                    *reinterpret_cast<float2*>(xr + e) = __ffma2_rn(cc, a_r, __ffma2_rn(nwr, b_r, __fmul2_rn(nwi, b_i)));
                    *reinterpret_cast<float2*>(xi + e) = __ffma2_rn(cc, a_i, __ffma2_rn(nwr, b_i, __fmul2_rn(pwi, b_r)));
                    *reinterpret_cast<float2*>(yr + e) = __ffma2_rn(cc, b_r, __ffma2_rn(pwr, a_r, __fmul2_rn(nwi, a_i)));
                    *reinterpret_cast<float2*>(yi + e) = __ffma2_rn(cc, b_i, __ffma2_rn(pwr, a_i, __fmul2_rn(pwi, a_r)));
                }
                if (g == 0) {
                    nrm[s1] = fmaxf(an - taz, 0.f);
                    nrm[s2] = fmaxf(bn + taz, 0.f);
                }
            }
            __syncthreads();
        }
        mymax = warp_max(mymax);
        if (lane == 0) atomicMax(&cta_max, __float_as_uint(mymax));
        __syncthreads();
        const float sweep_max = __uint_as_float(cta_max);
        __syncthreads();
        if (threadIdx.x == 0) cta_max = 0u;
        // refresh the cached norms from the data once per sweep (they are updated by formula in between)
        for (int v = warp; v < r; v += nwarps) {
            float s = 0.f;
            for (int t = lane; t < ldot; t += 32) {
                const float a_ = XR[(size_t)v * lpad + t], b_ = XI[(size_t)v * lpad + t];
                s = fmaf(a_, a_, fmaf(b_, b_, s));
            }
            s = warp_sum(s);
            if (lane == 0) nrm[v] = s;
        }
        __syncthreads();
        if (sweep_max <= tol2_stop) {
            converged = true;
            ++it;
            break;
        }
    }
    for (int v = warp; v < r; v += nwarps)
        for (int t = lane; t < ltot; t += 32)
            Wb[(size_t)v * ld + t] = make_float2(XR[(size_t)v * lpad + t], XI[(size_t)v * lpad + t]);
    if (threadIdx.x == 0) {
        sweeps[b] = it;
        done[b] = converged ? 1 : 0;
    }
}

template <int LPP>
int launch_small(vk_context* h, float2* W, int B, const JacobiPlan& p, float tol2_rot, float tol2_stop, int32_t* done,
                 int32_t* sweeps) {
    const int nslots = p.r + (p.r & 1);
    const int lpad = ((p.ltot + 1) / 2) * 2 + 8;
    const int npairs = nslots / 2;
    constexpr int GPW = 32 / LPP;
    int warps = (npairs + GPW - 1) / GPW;
    if (warps < 2) warps = 2;
    const size_t smem = (size_t)nslots * lpad * 8 + (size_t)nslots * 4;
    if (smem > 220 * 1024) return vk_fail(h, VK_EINVAL, "jacobi(small): problem does not fit shared memory");
    VK_CUDA(h, cudaFuncSetAttribute(jacobi_small_kernel<LPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_small_kernel<LPP><<<B, 32 * warps, smem, h->stream>>>(W, (size_t)p.r * p.ld, p.ld, p.ldot, p.ltot, p.r, nslots,
                                                                lpad, h->max_sweeps, tol2_rot, tol2_stop, done, sweeps);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// Register-resident variant of the small kernel for power-of-two slot counts. A sweep is a RECURSIVE block tournament:
// stage h = nslots/2, nslots/4, ..., 1 pairs the two halves of every block of 2h vectors. Within a stage group g keeps
// x = vector (blk*2h + pos) in registers and meets the h vectors of the other half one per round, (pos + q) mod h, read
// from and written back to shared memory. 32 + 16 + ... + 1 = nslots - 1 rounds per sweep as before, every group busy in
// every round, but only y moves through shared memory: ~2.2 KB per pair instead of 5 KB (x read twice + written, y read
// twice + written) in jacobi_small_kernel, which was bound by exactly that traffic.
template <int LPP, int NPL>
__global__ void __launch_bounds__(512)
jacobi_small_reg_kernel(float2* __restrict__ W, size_t mat_stride, int ld, int ldot, int ltot, int r, int nslots, int lpad,
                        int max_sweeps, float tol2_rot, float tol2_stop, int32_t* __restrict__ done,
                        int32_t* __restrict__ sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* XR = reinterpret_cast<float*>(smem_raw);      // [nslots][lpad]
    float* XI = XR + (size_t)nslots * lpad;               // [nslots][lpad]
    float* nrm = XI + (size_t)nslots * lpad;              // [nslots]
    __shared__ unsigned cta_max;
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int g = lane % LPP;
    const int grp = (threadIdx.x / LPP);                  // group index, nslots / 2 groups in the CTA
    float2* Wb = W + (size_t)b * mat_stride;

    for (int v = warp; v < nslots; v += nwarps) {
        float s = 0.f;
        for (int t = lane; t < lpad; t += 32) {
            float2 w = make_float2(0.f, 0.f);
            if (v < r && t < ltot) w = Wb[(size_t)v * ld + t];
            XR[(size_t)v * lpad + t] = w.x;
            XI[(size_t)v * lpad + t] = w.y;
            if (t < ldot) s = fmaf(w.x, w.x, fmaf(w.y, w.y, s));
        }
        s = warp_sum(s);
        if (lane == 0) nrm[v] = s;
    }
    if (threadIdx.x == 0) cta_max = 0u;
    __syncthreads();

    int it = 0;
    bool converged = false;
    for (; it < max_sweeps; ++it) {
        float mymax = 0.f;
        for (int h = nslots >> 1; h >= 1; h >>= 1) {
            const int blk = grp / h, pos = grp - blk * h;
            const int xs = blk * 2 * h + pos;
            float2 xr[NPL], xi[NPL];
            {
                const float* pr = XR + (size_t)xs * lpad + 2 * g;
                const float* pi = XI + (size_t)xs * lpad + 2 * g;
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    xr[p] = *reinterpret_cast<const float2*>(pr + 2 * LPP * p);
                    xi[p] = *reinterpret_cast<const float2*>(pi + 2 * LPP * p);
                }
            }
            float an = nrm[xs];
            for (int q = 0; q < h; ++q) {
                int yo = pos + q;
                if (yo >= h) yo -= h;
                const int ys = blk * 2 * h + h + yo;
                float* yrp = XR + (size_t)ys * lpad + 2 * g;
                float* yip = XI + (size_t)ys * lpad + 2 * g;
                float2 yr[NPL], yi[NPL];
                float2 P = make_float2(0.f, 0.f), Q = make_float2(0.f, 0.f), R = make_float2(0.f, 0.f);
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    yr[p] = *reinterpret_cast<const float2*>(yrp + 2 * LPP * p);
                    yi[p] = *reinterpret_cast<const float2*>(yip + 2 * LPP * p);
                    const int e = 2 * g + 2 * LPP * p;
                    if (e < ldot) {  // entries beyond ldot carry the accumulated rotations, not the vectors
                        float2 a_r = xr[p], a_i = xi[p];
                        if (e + 1 >= ldot) {
                            a_r.y = 0.f;
                            a_i.y = 0.f;
                        }
                        P = __ffma2_rn(a_r, yr[p], P);
                        P = __ffma2_rn(a_i, yi[p], P);
                        Q = __ffma2_rn(a_r, yi[p], Q);
                        R = __ffma2_rn(a_i, yr[p], R);
                    }
                }
                float zr = P.x + P.y, zi = (Q.x + Q.y) - (R.x + R.y);
#pragma unroll
                for (int o = LPP / 2; o > 0; o >>= 1) {
                    zr += __shfl_xor_sync(0xffffffffu, zr, o);
                    zi += __shfl_xor_sync(0xffffffffu, zi, o);
                }
                const float bn = nrm[ys];
                const float zz = zr * zr + zi * zi;
                float rel2 = 0.f;
                if (an > 0.f && bn > 0.f) rel2 = __fdividef(zz, an * bn);
                mymax = fmaxf(mymax, rel2);
                if (rel2 > tol2_rot && zz > 0.f) {
                    float c, wr, wi, taz;
                    rotation_params(an, bn, zr, zi, zz, c, wr, wi, taz);
                    const float2 cc = make_float2(c, c), pwr = make_float2(wr, wr), nwr = make_float2(-wr, -wr),
                                 pwi = make_float2(wi, wi), nwi = make_float2(-wi, -wi);
#pragma unroll
                    for (int p = 0; p < NPL; ++p) {
                        const float2 oxr = xr[p], oxi = xi[p];
                        xr[p] = __ffma2_rn(cc, oxr, __ffma2_rn(nwr, yr[p], __fmul2_rn(nwi, yi[p])));
                        xi[p] = __ffma2_rn(cc, oxi, __ffma2_rn(nwr, yi[p], __fmul2_rn(pwi, yr[p])));
                        *reinterpret_cast<float2*>(yrp + 2 * LPP * p) = __ffma2_rn(cc, yr[p], __ffma2_rn(pwr, oxr, __fmul2_rn(nwi, oxi)));
                        *reinterpret_cast<float2*>(yip + 2 * LPP * p) = __ffma2_rn(cc, yi[p], __ffma2_rn(pwr, oxi, __fmul2_rn(pwi, oxr)));
                    }
                    an = fmaxf(an - taz, 0.f);
                    if (g == 0) nrm[ys] = fmaxf(bn + taz, 0.f);
                }
                __syncthreads();
            }
            {
                float* pr = XR + (size_t)xs * lpad + 2 * g;
                float* pi = XI + (size_t)xs * lpad + 2 * g;
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    *reinterpret_cast<float2*>(pr + 2 * LPP * p) = xr[p];
                    *reinterpret_cast<float2*>(pi + 2 * LPP * p) = xi[p];
                }
                if (g == 0) nrm[xs] = an;
            }
            __syncthreads();
        }
        mymax = warp_max(mymax);
        if (lane == 0) atomicMax(&cta_max, __float_as_uint(mymax));
        __syncthreads();
        const float sweep_max = __uint_as_float(cta_max);
        __syncthreads();
        if (threadIdx.x == 0) cta_max = 0u;
        for (int v = warp; v < r; v += nwarps) {  // refresh the cached norms from the data once per sweep
            float s = 0.f;
            for (int t = lane; t < ldot; t += 32) {
                const float a_ = XR[(size_t)v * lpad + t], b_ = XI[(size_t)v * lpad + t];
                s = fmaf(a_, a_, fmaf(b_, b_, s));
            }
            s = warp_sum(s);
            if (lane == 0) nrm[v] = s;
        }
        __syncthreads();
        if (sweep_max <= tol2_stop) {
            converged = true;
            ++it;
            break;
        }
    }
    for (int v = warp; v < r; v += nwarps)
        for (int t = lane; t < ltot; t += 32)
            Wb[(size_t)v * ld + t] = make_float2(XR[(size_t)v * lpad + t], XI[(size_t)v * lpad + t]);
    if (threadIdx.x == 0) {
        sweeps[b] = it;
        done[b] = converged ? 1 : 0;
    }
}

template <int LPP, int NPL>
int launch_small_reg(vk_context* h, float2* W, int B, const JacobiPlan& p, int nslots, float tol2_rot, float tol2_stop,
                     int32_t* done, int32_t* sweeps) {
    const int lpad = 2 * LPP * NPL + 8;  // every lane's NPL element pairs exist (zero padded); +8 floats spreads banks
    const int threads = (nslots / 2) * LPP;
    const size_t smem = (size_t)nslots * lpad * 8 + (size_t)nslots * 4;
    VK_CUDA(h, cudaFuncSetAttribute(jacobi_small_reg_kernel<LPP, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_small_reg_kernel<LPP, NPL><<<B, threads, smem, h->stream>>>(W, (size_t)p.r * p.ld, p.ld, p.ldot, p.ltot, p.r, nslots,
                                                                       lpad, h->max_sweeps, tol2_rot, tol2_stop, done, sweeps);
    VK_LAUNCH_CHECK(h);
    return VK_OK;
}

// picks (lanes per pair, element pairs per lane) for the register kernel; returns false when it does not apply
bool small_reg_dispatch(vk_context* h, float2* W, int B, const JacobiPlan& p, float tol2_rot, float tol2_stop, int32_t* done,
                        int32_t* sweeps, int* rc) {
    int nslots = 2;
    while (nslots < p.r) nslots <<= 1;
    if (2 * nslots > 3 * (p.r + (p.r & 1))) return false;  // too many phantom vectors: the plain tournament is cheaper
    const int L = p.ltot;
    int lpp, npl;
    if (L <= 32) lpp = 4, npl = 4;
    else if (L <= 64) lpp = 4, npl = 8;
    else if (L <= 128) lpp = 8, npl = 8;
    else if (L <= 256) lpp = 16, npl = 8;
    else if (L <= 512) lpp = 32, npl = 8;
    else return false;
    const int threads = (nslots / 2) * lpp;
    if (threads > 512 || threads < 32) return false;
    if ((size_t)nslots * (2 * lpp * npl + 8) * 8 + 1024 > 220 * 1024) return false;
#define VK_SMALL_REG(LP, NP) *rc = launch_small_reg<LP, NP>(h, W, B, p, nslots, tol2_rot, tol2_stop, done, sweeps)
    if (lpp == 4 && npl == 4) VK_SMALL_REG(4, 4);
    else if (lpp == 4) VK_SMALL_REG(4, 8);
    else if (lpp == 8) VK_SMALL_REG(8, 8);
    else if (lpp == 16) VK_SMALL_REG(16, 8);
    else VK_SMALL_REG(32, 8);
#undef VK_SMALL_REG
    return true;
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

__global__ void count_active_kernel(int B, const int32_t* __restrict__ done, int32_t* __restrict__ active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && !done[b]) atomicAdd(active, 1);
}

int vk_launch_jacobi(vk_context* h, float2* W, int B, const JacobiPlan& p, int32_t* sweeps_dev, int32_t* done_dev,
                     unsigned* offmax_dev, int32_t* active_dev, bool preset_done) {
    if (B <= 0) return VK_OK;
    if (p.smem > VK_SMEM_BUDGET + 4096)
        return vk_fail(h, VK_EINVAL, "jacobi: a pair of vectors does not fit shared memory (matrix too large)");
    cudaStream_t st = h->stream;
    VK_CUDA(h, cudaFuncSetAttribute(jacobi_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    if (!preset_done) {
        VK_CUDA(h, cudaMemsetAsync(sweeps_dev, 0, sizeof(int32_t) * B, st));
        VK_CUDA(h, cudaMemsetAsync(done_dev, 0, sizeof(int32_t) * B, st));
    }
    VK_CUDA(h, cudaMemsetAsync(offmax_dev, 0, sizeof(unsigned) * B, st));
    if (preset_done) {
        // the fast path may have solved everything: one poll instead of a sweep of empty launches
        VK_CUDA(h, cudaMemsetAsync(active_dev, 0, sizeof(int32_t), st));
        count_active_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, done_dev, active_dev);
        VK_LAUNCH_CHECK(h);
        VK_CUDA(h, cudaMemcpyAsync(h->h_poll, active_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VK_CUDA(h, cudaStreamSynchronize(st));
        if (h->h_poll[0] == 0) return VK_OK;
    }
    // fp32 inner products of length ldot carry relative noise ~ sqrt(ldot) * 2^-24 (LAPACK xGESVJ uses the same
    // scale for its threshold): never rotate below it, and call a sweep converged when every off-diagonal it met
    // was below max(user tol, 4 * noise) — the rotations of that sweep then leave the matrix at the noise floor.
    const float noise = sqrtf((float)p.ldot) * 5.9604645e-8f;
    const float tol_stop = fmaxf(h->jacobi_tol, 4.f * noise);
    const float tol_rot = noise;  // the rotation threshold never follows the (looser) stop level
    const float tol2_stop = tol_stop * tol_stop;
    const float tol2_rot = tol_rot * tol_rot;
    const size_t mat_stride = (size_t)p.r * p.ld;
    const int threads = 32 * p.bsz;
    const long long nblocks = (long long)B * (p.nb / 2);
    if (nblocks > 0x7fffffffLL) return vk_fail(h, VK_EINVAL, "jacobi: batch too large");
    if (p.nb == 2 && !h->jacobi_generic) {
        int rc_small = VK_OK;
        if (h->small_reg && small_reg_dispatch(h, W, B, p, tol2_rot, tol2_stop, done_dev, sweeps_dev, &rc_small)) return rc_small;
        // elements per lane ~ 8..16: lanes per pair from the vector length
        const int want = (p.ltot + 15) / 16;
        if (want <= 4) return launch_small<4>(h, W, B, p, tol2_rot, tol2_stop, done_dev, sweeps_dev);
        if (want <= 8) return launch_small<8>(h, W, B, p, tol2_rot, tol2_stop, done_dev, sweeps_dev);
        if (want <= 16) return launch_small<16>(h, W, B, p, tol2_rot, tol2_stop, done_dev, sweeps_dev);
        return launch_small<32>(h, W, B, p, tol2_rot, tol2_stop, done_dev, sweeps_dev);
    }
    if (p.nb == 2) {
        jacobi_pairs_kernel<<<(unsigned)nblocks, threads, p.smem, st>>>(W, mat_stride, p.ld, p.ldot, p.ltot, p.r, p.bsz,
                                                                         p.nb, 0, 1, h->max_sweeps, tol2_rot, tol2_stop,
                                                                         offmax_dev, done_dev, sweeps_dev);
        VK_LAUNCH_CHECK(h);
        return VK_OK;
    }
    const int epl = (h->jacobi_generic ? 0 : cross_epl(p));
    const int npairs = p.nb / 2;

    // One launch covers B * npairs CTAs, usually a non-integer number of waves (KAT-7: 896 CTAs on 296 slots = 3.03
    // waves, i.e. a 4th, almost empty wave per launch). The matrices are therefore split into G independent groups,
    // each with its own stream: a group's launches stay ordered, different groups overlap, and no launch boundary
    // drains the whole GPU. Each group always has exactly one sweep in flight; the host polls a group's convergence
    // word while the other groups keep the SMs busy.
    const int slots = h->num_sms * ((epl && epl <= 8) ? 2 : 1);
    int G = h->jacobi_groups > 0 ? h->jacobi_groups : (int)(nblocks / (slots > 0 ? slots : 1));
    if (G > VK_MAX_GROUPS) G = VK_MAX_GROUPS;
    if (G > B) G = B;
    if (G < 1) G = 1;
    cudaStream_t gs[VK_MAX_GROUPS];
    int gb0[VK_MAX_GROUPS + 1];
    for (int g = 0; g <= G; ++g) gb0[g] = (int)((long long)B * g / G);
    if (G > 1) {
        for (int g = 0; g < G; ++g) {
            if (!h->sub[g]) VK_CUDA(h, cudaStreamCreateWithFlags(&h->sub[g], cudaStreamNonBlocking));
            if (!h->sub_ev[g]) VK_CUDA(h, cudaEventCreateWithFlags(&h->sub_ev[g], cudaEventDisableTiming));
            gs[g] = h->sub[g];
        }
        if (!h->fork_ev) VK_CUDA(h, cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
        VK_CUDA(h, cudaEventRecord(h->fork_ev, st));
        for (int g = 0; g < G; ++g) VK_CUDA(h, cudaStreamWaitEvent(gs[g], h->fork_ev, 0));
    } else {
        gs[0] = st;
    }

    auto issue_sweep = [&](int g) -> int {
        const int b0 = gb0[g], nbg = gb0[g + 1] - gb0[g];
        cudaStream_t sg = gs[g];
        float2* Wg = W + (size_t)b0 * mat_stride;
        const unsigned nblk = (unsigned)((long long)nbg * npairs);
        if (epl) {
            // pairs inside each block (generic kernel, intra-only), then every block pair with the register kernel
            jacobi_pairs_kernel<<<nblk, threads, p.smem, sg>>>(Wg, mat_stride, p.ld, p.ldot, p.ltot, p.r, p.bsz, p.nb, 0, 2,
                                                                1, tol2_rot, tol2_stop, offmax_dev + b0, done_dev + b0,
                                                                sweeps_dev + b0);
            VK_LAUNCH_CHECK(h);
            for (int round = 0; round < p.nb - 1; ++round) {
                const int rc = launch_cross_dispatch(h, sg, epl, Wg, mat_stride, p, round, nblk, tol2_rot, offmax_dev + b0,
                                                     done_dev + b0);
                if (rc) return rc;
            }
        } else {
            for (int round = 0; round < p.nb - 1; ++round) {
                jacobi_pairs_kernel<<<nblk, threads, p.smem, sg>>>(Wg, mat_stride, p.ld, p.ldot, p.ltot, p.r, p.bsz, p.nb,
                                                                    round, round == 0 ? 1 : 0, 1, tol2_rot, tol2_stop,
                                                                    offmax_dev + b0, done_dev + b0, sweeps_dev + b0);
                VK_LAUNCH_CHECK(h);
            }
        }
        VK_CUDA(h, cudaMemsetAsync(active_dev + 16 * g, 0, sizeof(int32_t), sg));
        sweep_check_kernel<<<(nbg + 255) / 256, 256, 0, sg>>>(nbg, tol2_stop, offmax_dev + b0, done_dev + b0,
                                                              sweeps_dev + b0, active_dev + 16 * g);
        VK_LAUNCH_CHECK(h);
        VK_CUDA(h, cudaMemcpyAsync(h->h_poll + 8 + g, active_dev + 16 * g, sizeof(int32_t), cudaMemcpyDeviceToHost, sg));
        return VK_OK;
    };

    int issued[VK_MAX_GROUPS];
    bool alive[VK_MAX_GROUPS];
    int nalive = G;
    for (int g = 0; g < G; ++g) {
        alive[g] = true;
        issued[g] = 1;
        const int rc = issue_sweep(g);
        if (rc) return rc;
    }
    while (nalive > 0) {
        for (int g = 0; g < G; ++g) {
            if (!alive[g]) continue;
            VK_CUDA(h, cudaStreamSynchronize(gs[g]));
            if (h->h_poll[8 + g] == 0 || issued[g] >= h->max_sweeps) {
                alive[g] = false;
                --nalive;
                continue;
            }
            ++issued[g];
            const int rc = issue_sweep(g);
            if (rc) return rc;
        }
    }
    if (G > 1) {
        for (int g = 0; g < G; ++g) {
            VK_CUDA(h, cudaEventRecord(h->sub_ev[g], gs[g]));
            VK_CUDA(h, cudaStreamWaitEvent(st, h->sub_ev[g], 0));
        }
    }
    return VK_OK;
}
