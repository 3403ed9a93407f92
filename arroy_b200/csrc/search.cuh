// search.cuh — batched query path on the device (SURVEY.md §8f "next" #1).
//
// walk_kernel restates the candidate-collecting loop of Reader::nns_by_leaf
// (src/reader.rs:328-374): a max-heap of (OrderedFloat(dist), NodeId), pop the greatest, a
// Descendants node appends its items, a SplitPlaneNormal node pushes both children with
// D::pq_distance(dist, D::margin(normal, query), side) (src/distance/mod.rs:63-68). One warp per
// query: lane 0 owns the binary heap (global memory, per query), all 32 lanes compute the margin
// in the reference's exact summation order. Items are de-duplicated with a per-query bitmap while
// they are appended (the reference sorts + dedups afterwards, reader.rs:378-379; the *count* that
// stops the walk includes duplicates, as in the reference). The unique candidates are then sorted
// (CUB segmented sort) and go through distance_kernel + topk_kernel unchanged.
#pragma once
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace ab {

struct DevForest {
    const uint8_t* kind;        // per node id: 0 = missing, 1 = Descendants, 2 = SplitPlaneNormal
    const uint32_t* left;
    const uint32_t* right;
    const uint32_t* normal_idx; // index into normals, 0xffffffff = "normal: none"
    const float* nh0;           // normal header 0 (bias / extra_dim)
    const uint32_t* desc_off;
    const uint32_t* desc_len;
    const float* normals;       // n_normals x ld
    const uint32_t* desc_rows;  // concatenated descendant lists, as ROW indices
    const uint32_t* roots;
    uint32_t n_roots, n_nodes;
};

__device__ __forceinline__ float key_to_dist(uint32_t k) {  // inverse of ordered_key (canonical +0 / NaN)
    if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
__device__ __forceinline__ float f32_min_dev(float a, float b) {  // Rust f32::min: NaN loses
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}

__device__ __forceinline__ void heap_push(unsigned long long* h, uint32_t& size, unsigned long long v) {
    uint32_t i = size++;
    while (i > 0) {
        uint32_t p = (i - 1) >> 1;
        unsigned long long pv = h[p];
        if (pv >= v) break;
        h[i] = pv;
        i = p;
    }
    h[i] = v;
}
__device__ __forceinline__ unsigned long long heap_pop(unsigned long long* h, uint32_t& size) {
    unsigned long long top = h[0];
    unsigned long long last = h[--size];
    uint32_t i = 0;
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1;
        if (l >= size) break;
        unsigned long long lv = h[l];
        uint32_t c = l;
        unsigned long long cv = lv;
        if (r < size) { unsigned long long rv = h[r]; if (rv > lv) { c = r; cv = rv; } }
        if (cv <= last) break;
        h[i] = cv;
        i = c;
    }
    if (size > 0) h[i] = last;
    return top;
}

constexpr int WALK_WARPS = 8;
constexpr uint32_t WALK_SHEAP = 512;    // heap entries per query kept in shared memory (spills to the global heap beyond)

// query q: vector = qrows ? items[qrows[q]] : queries[q] (ld floats); qh0 = extra_dim for DotProduct margins
__global__ void __launch_bounds__(WALK_WARPS * 32)
walk_kernel(DevForest F, const float* __restrict__ items, uint32_t d, uint32_t ld, int metric, uint32_t nq,
            const uint32_t* __restrict__ qrows, const float* __restrict__ queries, const float* __restrict__ qh0,
            unsigned long long search_k, unsigned long long* __restrict__ heaps, uint32_t heap_cap,
            uint32_t* __restrict__ cand, uint32_t cand_cap, uint32_t* __restrict__ cand_count,
            uint32_t* __restrict__ bitmap, uint32_t bitmap_words, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * WALK_WARPS + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float* qv = qrows ? items + (size_t)qrows[q] * ld : queries + (size_t)q * ld;
    const float qhdr = qh0 ? qh0[q] : 0.f;
    // The heap lives in shared memory (a walk pushes two entries per pop: a few hundred in practice) and moves to
    // its global-memory slot only if it outgrows WALK_SHEAP: every pop / push is a chain of dependent accesses.
    extern __shared__ unsigned long long walk_sheap[];
    unsigned long long* gheap = heaps + (size_t)q * heap_cap;
    unsigned long long* heap = walk_sheap + (size_t)(threadIdx.x >> 5) * WALK_SHEAP;
    uint32_t cap = WALK_SHEAP < heap_cap ? WALK_SHEAP : heap_cap;
    uint32_t* out = cand + (size_t)q * cand_cap;
    uint32_t* bm = bitmap + (size_t)q * bitmap_words;
    uint32_t size = 0;
    if (F.n_roots > cap) { heap = gheap; cap = heap_cap; }
    if (lane == 0) {
        const unsigned long long inf_key = (unsigned long long)ordered_key(__uint_as_float(0x7f800000u)) << 32;
        for (uint32_t r = 0; r < F.n_roots && size < cap; ++r) heap_push(heap, size, inf_key | F.roots[r]);
    }
    unsigned long long total = 0;   // nns.len() of the reference (duplicates included)
    uint32_t unique = 0;
    int st = 0;
    for (;;) {
        size = __shfl_sync(0xffffffffu, size, 0);
        if (total >= search_k || size == 0) break;
        unsigned long long top = 0;
        if (lane == 0) top = heap_pop(heap, size);
        top = __shfl_sync(0xffffffffu, top, 0);
        const uint32_t node = (uint32_t)top;
        const float dist = key_to_dist((uint32_t)(top >> 32));
        const int kind = node < F.n_nodes ? F.kind[node] : 0;
        if (kind == 1) {
            const uint32_t off = F.desc_off[node], len = F.desc_len[node];
            for (uint32_t i0 = 0; i0 < len; i0 += 32) {
                uint32_t i = i0 + lane;
                bool fresh = false;
                uint32_t row = 0;
                if (i < len) {
                    row = F.desc_rows[off + i];
                    uint32_t bit = 1u << (row & 31);
                    fresh = (atomicOr(&bm[row >> 5], bit) & bit) == 0;
                }
                unsigned m = __ballot_sync(0xffffffffu, fresh);
                if (fresh) {
                    uint32_t p = unique + __popc(m & ((1u << lane) - 1u));
                    if (p < cand_cap) out[p] = row; else st = 1;
                }
                unique += __popc(m);
            }
            total += len;
        } else if (kind == 2) {
            const uint32_t ni = F.normal_idx[node];
            float mg = 0.0f;
            if (ni != 0xffffffffu) {
                const float* nv = F.normals + (size_t)ni * ld;
                float dt = exact_warp<false>(nv, qv, (int)d);
                mg = margin_finish(metric, dt, F.nh0[node], qhdr);
            }
            // outgrown the shared-memory heap: move it (the array *is* the heap) to the global slot
            const int spill = __shfl_sync(0xffffffffu, (int)(size + 2 > cap && heap != gheap), 0);
            if (spill) {
                const uint32_t sz = __shfl_sync(0xffffffffu, size, 0);
                for (uint32_t i = lane; i < sz; i += 32) gheap[i] = heap[i];
                __syncwarp();
                heap = gheap; cap = heap_cap;
            }
            if (lane == 0) {
                if (size + 2 > cap) st = 2;
                else {
                    heap_push(heap, size, ((unsigned long long)ordered_key(f32_min_dev(-mg, dist)) << 32) | F.left[node]);
                    heap_push(heap, size, ((unsigned long long)ordered_key(f32_min_dev(mg, dist)) << 32) | F.right[node]);
                }
            }
        } else {
            st = 3;  // missing node (Error::MissingKey)
        }
        st = __reduce_max_sync(0xffffffffu, st);
        if (st) break;
    }
    if (lane == 0) { cand_count[q] = unique < cand_cap ? unique : cand_cap; status[q] = st; }
}

// offsets[q] = q * cand_cap (begin), ends[q] = begin + count — segment descriptors for the sort
__global__ void walk_segments_kernel(const uint32_t* __restrict__ cand_count, uint32_t nq, uint32_t cand_cap, uint64_t* __restrict__ begins, uint64_t* __restrict__ ends) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) { begins[q] = (uint64_t)q * cand_cap; ends[q] = (uint64_t)q * cand_cap + cand_count[q]; }
    if (q == nq) { begins[q] = (uint64_t)q * cand_cap; }
}

__global__ void gather_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

}  // namespace ab
