// xrerank.cuh — batched re-rank of MANY queries against ONE shared candidate list (BASELINE.json
// config 5: 4096 queries x 100k candidates, d = 768), bit-exact.
//
// D::built_distance (src/reader.rs:381-391) for every (query, candidate) pair is a dense
// Q x C x d contraction. Tensor cores cannot be used for it: the result contract is "top-k ids
// identical to the reference", and the reference's value for each pair is defined by its AVX+FMA
// summation order (src/spaces/simple_avx.rs:6-110) — 32 accumulator lanes (k mod 32), each a
// sequential FMA chain over the 32-chunks, then the hsum256 tree. So this is a register-tiled
// CUDA-core kernel that keeps exactly that order per pair:
//   * lane l of a warp IS accumulator lane l; a thread holds an 8 x 8 tile of pairs (64 chains),
//     so one chunk costs 16 shared-memory loads for 64 FMAs;
//   * operands (16 query rows x 32 candidate rows per CTA, 64 floats of k per stage) are staged with
//     cp.async, double buffered;
//   * the final reduction is a transposing butterfly (xor 4, 2, 1: each step halves the number of
//     pairs a lane keeps, 56 shuffles instead of 192), then ((h1+h2)+h3)+h4 across the four
//     accumulator groups — the same additions, in the same order, as hsum256 + the final sum.
// Manhattan (strictly sequential scalar sum) and d < 32 (SSE / scalar paths) use the generic kernels.
#pragma once
#include "kernels.cuh"

namespace ab {

constexpr int XQ = 8, XC = 8;        // register tile per warp (queries x candidates)
constexpr int XWQ = 2, XWC = 4;      // warps per CTA along q and c
constexpr int XQB = XQ * XWQ;        // 16 queries per CTA
constexpr int XCB = XC * XWC;        // 32 candidates per CTA
constexpr int XSLAB = 64;            // floats of k per pipeline stage (2 chunks)
constexpr int XTHREADS = 32 * XWQ * XWC;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// dist[q * nc + c] = built_distance(query q, item rows[c]); grid = (ceil(nc / XCB), ceil(nq / XQB))
template <bool EUCLID>
__global__ void __launch_bounds__(XTHREADS)
xrerank_kernel(const float* __restrict__ items, const float* __restrict__ ih0, uint32_t d, uint32_t ld, int metric,
               const float* __restrict__ queries /* nq x ld */, const float* __restrict__ qh0, uint32_t nq,
               const uint32_t* __restrict__ rows, uint32_t nc, float* __restrict__ dist) {
    __shared__ __align__(16) float sQ[2][XQB][XSLAB];
    __shared__ __align__(16) float sC[2][XCB][XSLAB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wq = warp / XWC, wc = warp % XWC;
    const uint32_t qb = blockIdx.y * XQB, cb = blockIdx.x * XCB;
    const uint32_t m = d & ~31u;                       // AVX main part
    const int nstage = (int)((m + XSLAB - 1) / XSLAB);

    // this thread's cp.async sources: 1 float4 of Q, 2 float4 of C per stage
    const int qr = tid / (XSLAB / 4), qc4 = tid % (XSLAB / 4);          // 256 threads -> 16 rows x 16 float4
    const uint32_t qrow = min(qb + (uint32_t)qr, nq - 1);
    const float* qsrc = queries + (size_t)qrow * ld;
    const float* csrc[2];
    int cr[2], cc4[2];
    for (int h = 0; h < 2; ++h) {
        int idx = tid + h * XTHREADS;
        cr[h] = idx / (XSLAB / 4); cc4[h] = idx % (XSLAB / 4);
        uint32_t crow = min(cb + (uint32_t)cr[h], nc - 1);
        csrc[h] = items + (size_t)rows[crow] * ld;
    }
    auto issue = [&](int st, int buf) {
        const uint32_t k0 = (uint32_t)st * XSLAB;
        if (k0 + qc4 * 4 < m) cp_async16(&sQ[buf][qr][qc4 * 4], qsrc + k0 + qc4 * 4);
        for (int h = 0; h < 2; ++h)
            if (k0 + cc4[h] * 4 < m) cp_async16(&sC[buf][cr[h]][cc4[h] * 4], csrc[h] + k0 + cc4[h] * 4);
        cp_async_commit();
    };

    float acc[XQ][XC];
#pragma unroll
    for (int i = 0; i < XQ; ++i)
#pragma unroll
        for (int j = 0; j < XC; ++j) acc[i][j] = 0.f;

    if (nstage > 0) issue(0, 0);
    for (int st = 0; st < nstage; ++st) {
        const int buf = st & 1;
        if (st + 1 < nstage) { issue(st + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const uint32_t k0 = (uint32_t)st * XSLAB;
#pragma unroll
        for (int kk = 0; kk < XSLAB / 32; ++kk) {
            if (k0 + kk * 32 < m) {
                float qv[XQ], cv[XC];
#pragma unroll
                for (int i = 0; i < XQ; ++i) qv[i] = sQ[buf][wq * XQ + i][kk * 32 + lane];
#pragma unroll
                for (int j = 0; j < XC; ++j) cv[j] = sC[buf][wc * XC + j][kk * 32 + lane];
#pragma unroll
                for (int i = 0; i < XQ; ++i)
#pragma unroll
                    for (int j = 0; j < XC; ++j) {
                        if (EUCLID) { float t = __fsub_rn(qv[i], cv[j]); acc[i][j] = fmaf(t, t, acc[i][j]); }
                        else acc[i][j] = fmaf(qv[i], cv[j], acc[i][j]);
                    }
            }
        }
        __syncthreads();
    }

    // ---- hsum256 for 64 pairs at once: transposing butterfly --------------------------------------
    const unsigned full = 0xffffffffu;
    const bool b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
    float v1[XQ][4];
#pragma unroll
    for (int i = 0; i < XQ; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // x128[u] = x[u+4] + x[u]; this lane keeps pairs (i, j + 4*b2)
            float send = b2 ? acc[i][j] : acc[i][j + 4];
            float mine = b2 ? acc[i][j + 4] : acc[i][j];
            v1[i][j] = __fadd_rn(mine, __shfl_xor_sync(full, send, 4));
        }
    float v2[XQ][2];
#pragma unroll
    for (int i = 0; i < XQ; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {   // x64[0] = x128[0]+x128[2], x64[1] = x128[1]+x128[3]
            float send = b1 ? v1[i][j] : v1[i][j + 2];
            float mine = b1 ? v1[i][j + 2] : v1[i][j];
            v2[i][j] = __fadd_rn(mine, __shfl_xor_sync(full, send, 2));
        }
    float h[XQ];
#pragma unroll
    for (int i = 0; i < XQ; ++i) {      // x64[0] + x64[1]
        float send = b0 ? v2[i][0] : v2[i][1];
        float mine = b0 ? v2[i][1] : v2[i][0];
        h[i] = __fadd_rn(mine, __shfl_xor_sync(full, send, 1));
    }
    // lane (g, t) now holds accumulator g's horizontal sum for the pairs (i = 0..7, j = t)
    const int t = lane & 7;
    float res[XQ];
#pragma unroll
    for (int i = 0; i < XQ; ++i) {
        float s1 = __shfl_sync(full, h[i], t + 8), s2 = __shfl_sync(full, h[i], t + 16), s3 = __shfl_sync(full, h[i], t + 24);
        res[i] = __fadd_rn(__fadd_rn(__fadd_rn(h[i], s1), s2), s3);   // ((h1 + h2) + h3) + h4
    }
    if (lane < 8) {
        const uint32_t c = cb + wc * XC + t;
        if (c < nc) {
            const uint32_t crow = rows[c];
            const float* cptr = items + (size_t)crow * ld;
            const float ch = (metric == COSINE) ? ih0[crow] : 0.f;
#pragma unroll
            for (int i = 0; i < XQ; ++i) {
                const uint32_t q = qb + wq * XQ + i;
                if (q < nq) {
                    float r = res[i];
                    const float* qptr = queries + (size_t)q * ld;
                    for (uint32_t k = m; k < d; ++k) {   // len % 32 tail: separately rounded
                        if (EUCLID) { float tt = __fsub_rn(qptr[k], cptr[k]); r = __fadd_rn(r, __fmul_rn(tt, tt)); }
                        else r = __fadd_rn(r, __fmul_rn(qptr[k], cptr[k]));
                    }
                    dist[(size_t)q * nc + c] = built_finish(metric, r, qh0 ? qh0[q] : 0.f, ch);
                }
            }
        }
    }
}

// top-k per query straight from a dense distance row: key = ordered_key(dist) << 32 | position
__global__ void __launch_bounds__(TOPK_THREADS)
topk_dense_kernel(const float* __restrict__ dist, const uint32_t* __restrict__ rows, uint32_t nc, uint32_t k, int metric,
                  uint32_t* __restrict__ out_rows, float* __restrict__ out_dist, uint32_t* __restrict__ out_len) {
    __shared__ unsigned long long buf[TOPK_CAP];
    __shared__ uint32_t fill;
    const uint32_t q = blockIdx.x;
    const float* dq = dist + (size_t)q * nc;
    const uint32_t kk = nc < k ? nc : k;
    unsigned long long threshold = ~0ull;
    uint32_t pos = 0;
    bool have = false;
    while (pos < nc) {
        const uint32_t base = have ? kk : 0u;
        const uint32_t room = TOPK_CAP - base;
        const uint32_t take = nc - pos < room ? nc - pos : room;
        __syncthreads();
        if (threadIdx.x == 0) fill = base;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < take; i += blockDim.x) {
            unsigned long long key = ((unsigned long long)ordered_key(dq[pos + i]) << 32) | (unsigned long long)(pos + i);
            if (key < threshold) { uint32_t s = atomicAdd(&fill, 1u); buf[s] = key; }
        }
        pos += take;
        __syncthreads();
        const uint32_t f = fill;
        int mm = 2;
        while ((uint32_t)mm < f) mm <<= 1;
        for (int i = (int)f + threadIdx.x; i < mm; i += blockDim.x) buf[i] = ~0ull;
        bitonic_sort_shared(buf, mm);
        have = true;
        threshold = (f >= kk && kk > 0) ? buf[kk - 1] : ~0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) out_len[q] = kk;
    for (uint32_t i = threadIdx.x; i < kk; i += blockDim.x) {
        uint32_t p = (uint32_t)(buf[i] & 0xffffffffull);
        out_rows[(size_t)q * k + i] = rows[p];
        out_dist[(size_t)q * k + i] = normalized_distance_dev(metric, dq[p]);
    }
}

}  // namespace ab
