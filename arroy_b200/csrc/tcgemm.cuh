// tcgemm.cuh — S = Q . C^T on the 5th-generation tensor cores (tcgen05, TF32 inputs, FP32
// accumulation in TMEM), hand-written for sm_100a. It is the first stage of the shared-candidate
// re-rank (xrerank.cuh, "tensor-core pre-filter"): S only has to *bound* the reference's distance
// of every (query, candidate) pair, the survivors are re-scored in the reference's exact order.
//
//   Q : m  x K  fp32, row-major, pitch ld (K = ld, padding is zero)      -> UMMA operand A, K-major
//   C : nc x K  fp32, row-major, pitch ld (item rows, in place)          -> UMMA operand B, K-major
//   S : m  x nc fp32, row-major, pitch lds
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2D, 128B-swizzled boxes of 32 floats of K
//              (128 x 32 of Q, 256 x 32 of C) into a 4-stage shared-memory ring, mbarrier tx-counts
//   warp 1     allocates TMEM (512 columns = two 128 x 256 FP32 accumulators) and issues
//              tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 256, K = 8), four per stage;
//              tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld 32x32b.x32 (a warp may read the 32 TMEM lanes = 32 queries of
//              its quarter; two warps per quarter split the 256 columns), per-column constants of
//              the tile staged in shared memory, distance estimate, 16-byte stores; the other
//              accumulator is being filled meanwhile
// Tiles are ordered candidates-major so that the CTAs running at the same time share the same
// candidate rows in L2 and C streams from HBM once.
#pragma once
#include <cuda.h>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ab {

constexpr int TG_BM = 128;      // queries per tile   (UMMA M)
constexpr int TG_BN = 256;      // candidates per tile (UMMA N)
constexpr int TG_BK = 32;       // floats of K per stage = one 128-byte swizzle span
constexpr int TG_UK = 8;        // K of one tcgen05.mma.kind::tf32
constexpr int TG_STAGES = 4;
constexpr int TG_EPI_WARPS = 8;                // two per TMEM lane quarter, 128 columns each
constexpr int TG_THREADS = 64 + 32 * TG_EPI_WARPS;
constexpr uint32_t TG_A_BYTES = TG_BM * TG_BK * 4;
constexpr uint32_t TG_B_BYTES = TG_BN * TG_BK * 4;
constexpr uint32_t TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr size_t TG_SMEM = (size_t)TG_STAGES * TG_STAGE_BYTES + 1024 /* 1024-byte alignment */ + 256 /* barriers */ + 2 * 2 * TG_BN * 4 /* column constants */;
constexpr uint32_t TG_TMEM_COLS = 512;

__device__ __forceinline__ uint32_t tg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tg_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on a phase parity. A protocol bug must not hang the GPU: trap after ~2 s.
__device__ __forceinline__ void tg_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if ((spin & 0xffffu) == 0xffffu) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void tg_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// the same load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tg_tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tg_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tg_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tg_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 256
constexpr uint32_t TG_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TG_BN >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);

__device__ __forceinline__ void tg_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(TG_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tg_commit(uint32_t bar) {   // arrives on `bar` once every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tg_commit_mc(uint32_t bar, uint16_t mask) {   // ... on the barrier at this offset in every CTA of `mask`
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tg_tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// What the epilogue writes for pair (q, c) from the raw contraction s = Q[q] . C[c]:
//   TG_RAW     s
//   TG_NEG     -s                                             (DotProduct built_distance, dot_product.rs:52-56)
//   TG_EUCLID  (qa[q] + ca[c]) - 2 s                          (qa, ca = squared norms; euclidean.rs:45-47)
//   TG_COSINE  qb[q] * cb[c] > f32::EPSILON ? (1 - clamp(s * qa[q] * ca[c])) / 2 : 0     (cosine.rs:43-59;
//              qa, ca = reciprocal header norms, qb, cb = header norms)
// i.e. an FP32 *estimate* of the reference's built_distance; xrerank.cuh bounds its error.
enum { TG_RAW = 0, TG_NEG = 1, TG_EUCLID = 2, TG_COSINE = 3 };

struct TgEpilogue {
    int mode;
    const float* qa; const float* qb;   // per query (row)
    const float* ca; const float* cb;   // per candidate (column)
};

__device__ __forceinline__ float tg_finish(int mode, float s, float qa, float qb, float ca, float cb) {
    if (mode == TG_RAW) return s;
    if (mode == TG_NEG) return -s;
    if (mode == TG_EUCLID) return __fsub_rn(__fadd_rn(qa, ca), __fmul_rn(2.0f, s));
    const float pnqn = __fmul_rn(qb, cb);
    if (pnqn > 1.1920928955078125e-07f) {
        float c = __fmul_rn(s, __fmul_rn(qa, ca));
        c = c < -1.0f ? -1.0f : (c > 1.0f ? 1.0f : c);   // NaN stays NaN
        return __fmul_rn(0.5f, __fsub_rn(1.0f, c));
    }
    return pnqn == pnqn ? 0.0f : pnqn;
}

// MC = 1: independent CTAs. MC = 2: clusters of two CTAs that work on the same 256 candidates and two
// neighbouring blocks of 128 queries; each CTA fetches one half of the candidate tile and TMA-multicasts it
// into both CTAs' shared memory, so the operand traffic L2 -> SM per CTA drops from 48 to 32 KB per stage (that
// traffic, not the tensor pipe, limits the MC = 1 kernel). A stage is released to both producers by both MMA
// warps (tcgen05.commit.multicast on the `empty` barriers, which therefore count two arrivals).
template <int MC>
__global__ void __launch_bounds__(TG_THREADS, 1)
tcgemm_tf32_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                   float* __restrict__ S, uint32_t m, uint32_t nc, uint32_t lds, uint32_t nk, TgEpilogue ep) {
    extern __shared__ uint8_t tg_raw[];
    const uint32_t raw = tg_smem_u32(tg_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                      // swizzle-128B tiles need 1024-byte alignment
    const uint32_t bars = base + TG_STAGES * TG_STAGE_BYTES;           // full[4], empty[4], tmem_full[2], tmem_empty[2], tmem ptr
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (TG_STAGES + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * TG_STAGES + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * TG_STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * TG_STAGES + 4);
    float* cst = reinterpret_cast<float*>(tg_raw + (bars + 256u - raw));   // [2 accumulators][ca, cb][TG_BN]
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(tg_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = MC > 1 ? cooperative_groups::this_cluster().block_rank() : 0u;
    const uint32_t num_mb = ((m + TG_BM - 1) / TG_BM + MC - 1) / MC, num_n = (nc + TG_BN - 1) / TG_BN;   // m-blocks per cluster step
    const uint32_t tiles = num_mb * num_n, first = blockIdx.x / MC, stride = gridDim.x / MC;
    auto tile_m0 = [&](uint32_t t) { return ((t % num_mb) * MC + crank) * TG_BM; };   // may lie beyond m: zero rows, nothing stored
    auto tile_n0 = [&](uint32_t t) { return (t / num_mb) * TG_BN; };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TG_STAGES; ++s) { tg_mbar_init(full(s), 1); tg_mbar_init(empty(s), MC); }
        for (int a = 0; a < 2; ++a) { tg_mbar_init(tfull(a), 1); tg_mbar_init(tempty(a), TG_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tg_fence_before();
    __syncthreads();
    if (MC > 1) cooperative_groups::this_cluster().sync();   // the peer's barriers exist before anything is multicast to them
    tg_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (uint32_t t = first; t < tiles; t += stride) {
                const int m0 = (int)tile_m0(t), n0 = (int)tile_n0(t);
                for (uint32_t kb = 0; kb < nk; ++kb) {
                    tg_mbar_wait(empty(stage), phase ^ 1u);
                    tg_mbar_expect_tx(full(stage), TG_STAGE_BYTES);
                    const uint32_t sa = base + stage * TG_STAGE_BYTES, sb = sa + TG_A_BYTES;
                    tg_tma_load_2d(sa, &map_q, full(stage), (int)(kb * TG_BK), m0);
                    if (MC == 1) tg_tma_load_2d(sb, &map_c, full(stage), (int)(kb * TG_BK), n0);
                    else tg_tma_load_2d_mc(sb + crank * (TG_B_BYTES / MC), &map_c, full(stage), (int)(kb * TG_BK), n0 + (int)(crank * (TG_BN / MC)), (uint16_t)((1u << MC) - 1u));
                    if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, it = 0;
            for (uint32_t t = first; t < tiles; t += stride, ++it) {
                const uint32_t a = it & 1u, aphase = (it >> 1) & 1u;
                tg_mbar_wait(tempty(a), aphase ^ 1u);                 // epilogue has drained this accumulator
                tg_fence_after();
                const uint32_t d = tmem_base + a * TG_BN;
                for (uint32_t kb = 0; kb < nk; ++kb) {
                    tg_mbar_wait(full(stage), phase);                  // TMA has landed this stage
                    tg_fence_after();
                    const uint32_t sa = base + stage * TG_STAGE_BYTES, sb = sa + TG_A_BYTES;
                    const uint64_t da = tg_smem_desc(sa), db = tg_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < TG_BK / TG_UK; ++k)            // +32 bytes of K inside the swizzle span per step
                        tg_mma_tf32(d, da + (uint64_t)(k * TG_UK * 4 / 16), db + (uint64_t)(k * TG_UK * 4 / 16), (kb | (uint32_t)k) != 0u);
                    if (MC == 1) tg_commit(empty(stage));              // frees the stage when these MMAs retire
                    else tg_commit_mc(empty(stage), (uint16_t)((1u << MC) - 1u));   // ... in both CTAs: both producers write into both
                    if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
                }
                tg_commit(tfull(a));                                   // accumulator complete
            }
        }
    } else {
        const uint32_t qtr = (uint32_t)warp & 3u;                      // TMEM lane quarter this warp may read
        const uint32_t half = (uint32_t)(warp - 2) >> 2;               // which 128 columns of the tile
        const uint32_t et = threadIdx.x - 64u;                         // 0 .. 255 over the epilogue warps
        uint32_t it = 0;
        for (uint32_t t = first; t < tiles; t += stride, ++it) {
            const uint32_t a = it & 1u, aphase = (it >> 1) & 1u;
            const uint32_t m0 = tile_m0(t), n0 = tile_n0(t);
            float* ca_s = cst + a * 2 * TG_BN;
            float* cb_s = ca_s + TG_BN;
            if (ep.mode >= TG_EUCLID) {                                // this tile's column constants -> shared memory
                const uint32_t col = n0 + et;
                ca_s[et] = col < nc ? __ldg(ep.ca + col) : 0.f;
                cb_s[et] = (ep.mode == TG_COSINE && col < nc) ? __ldg(ep.cb + col) : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TG_EPI_WARPS) : "memory");
            }
            const uint32_t row = m0 + qtr * 32u + (uint32_t)lane;
            float qa = 0.f, qb = 0.f;
            if (row < m && ep.mode >= TG_EUCLID) { qa = ep.qa[row]; if (ep.mode == TG_COSINE) qb = ep.qb[row]; }
            float* out = S + (size_t)row * lds + n0 + half * 128u;
            tg_mbar_wait(tfull(a), aphase);
            tg_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < 4; ++c) {
                uint32_t v[32];
                tg_tmem_ld32(tmem_base + ((qtr * 32u) << 16) + a * TG_BN + half * 128u + c * 32u, v);
                const uint32_t col = n0 + half * 128u + c * 32u;
                if (ep.mode == TG_NEG) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] ^= 0x80000000u;
                } else if (ep.mode >= TG_EUCLID) {
                    const float4* ca4 = reinterpret_cast<const float4*>(ca_s + half * 128u + c * 32u);
                    const float4* cb4 = reinterpret_cast<const float4*>(cb_s + half * 128u + c * 32u);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = ca4[j];
                        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ep.mode == TG_COSINE) y = cb4[j];
                        v[4 * j + 0] = __float_as_uint(tg_finish(ep.mode, __uint_as_float(v[4 * j + 0]), qa, qb, x.x, y.x));
                        v[4 * j + 1] = __float_as_uint(tg_finish(ep.mode, __uint_as_float(v[4 * j + 1]), qa, qb, x.y, y.y));
                        v[4 * j + 2] = __float_as_uint(tg_finish(ep.mode, __uint_as_float(v[4 * j + 2]), qa, qb, x.z, y.z));
                        v[4 * j + 3] = __float_as_uint(tg_finish(ep.mode, __uint_as_float(v[4 * j + 3]), qa, qb, x.w, y.w));
                    }
                }
                if (row < m) {
                    if (col + 32u <= lds) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4*>(out + c * 32u + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (col + j < nc) out[c * 32u + j] = __uint_as_float(v[j]);
                    }
                }
            }
            tg_fence_before();
            __syncwarp();
            if (lane == 0) tg_mbar_arrive(tempty(a));
        }
    }
    tg_fence_before();
    __syncthreads();
    if (MC > 1) cooperative_groups::this_cluster().sync();   // no CTA leaves while its peer can still multicast into it
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TG_TMEM_COLS) : "memory");
}

// the same epilogue as a separate pass (used after the cuBLAS cross-check engine)
__global__ void tg_finish_kernel(float* __restrict__ S, uint32_t m, uint32_t nc, uint32_t lds, TgEpilogue ep) {
    const uint64_t total = (uint64_t)m * nc;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / nc), c = (uint32_t)(i - (uint64_t)r * nc);
        float qa = 0.f, qb = 0.f, ca = 0.f, cb = 0.f;
        if (ep.mode >= TG_EUCLID) { qa = ep.qa[r]; ca = ep.ca[c]; if (ep.mode == TG_COSINE) { qb = ep.qb[r]; cb = ep.cb[c]; } }
        float* p = S + (size_t)r * lds + c;
        *p = tg_finish(ep.mode, *p, qa, qb, ca, cb);
    }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*tg_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tg_encode_fn tg_encoder() {
    static tg_encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<tg_encode_fn>(p);
    }();
    return fn;
}

// rows x ld fp32 matrix, boxes of 32 floats x box_rows rows, 128-byte swizzle, zero fill outside
inline bool tg_make_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint32_t ld, uint32_t box_rows) {
    tg_encode_fn enc = tg_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {ld, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TG_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// S[m x nc] (pitch lds, lds % 4 == 0) = Q[m x ld] . C[nc x ld]^T; false if the tensor maps cannot be built
inline bool tcgemm_tf32(const float* Q, uint32_t m, const float* C, uint32_t nc, uint32_t ld, float* S, uint32_t lds, TgEpilogue ep, int sm_count, cudaStream_t stream, int mc = 2) {
    if (mc != 1 && mc != 2) mc = 2;
    CUtensorMap mq, mcand;
    if (!tg_make_map(&mq, Q, m, ld, TG_BM) || !tg_make_map(&mcand, C, nc, ld, TG_BN / mc)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tcgemm_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(tcgemm_tf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM) != cudaSuccess) return false;
        configured = true;
    }
    const uint32_t num_mb = ((m + TG_BM - 1) / TG_BM + mc - 1) / mc;
    const uint32_t tiles = num_mb * ((nc + TG_BN - 1) / TG_BN);
    const uint32_t units = (uint32_t)(sm_count / mc);
    const int grid = (int)((tiles < units ? tiles : units) * mc);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TG_THREADS); cfg.dynamicSmemBytes = TG_SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)mc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const uint32_t nk = ld / TG_BK;
    cudaError_t e = mc == 1 ? cudaLaunchKernelEx(&cfg, tcgemm_tf32_kernel<1>, mq, mcand, S, m, nc, lds, nk, ep)
                            : cudaLaunchKernelEx(&cfg, tcgemm_tf32_kernel<2>, mq, mcand, S, m, nc, lds, nk, ep);
    return e == cudaSuccess;
}

}  // namespace ab
