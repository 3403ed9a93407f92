// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
// C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline leg
// can drive it through ctypes. Not linked into, nor loaded by, the product library.
#include <cstring>

#include "build.hpp"

using namespace oracle;

extern "C" {

// ---- RNG ---------------------------------------------------------------------------
void* oracle_rng_from_seed(const uint8_t* seed32) { return new StdRng(StdRng::from_seed(seed32)); }
void* oracle_rng_seed_from_u64(uint64_t s) { return new StdRng(StdRng::seed_from_u64(s)); }
void* oracle_rng_clone(void* r) { return new StdRng(*static_cast<StdRng*>(r)); }
void oracle_rng_free(void* r) { delete static_cast<StdRng*>(r); }
uint32_t oracle_rng_next_u32(void* r) { return static_cast<StdRng*>(r)->next_u32(); }
uint64_t oracle_rng_next_u64(void* r) { return static_cast<StdRng*>(r)->next_u64(); }
float oracle_rng_gen_f32(void* r) { return static_cast<StdRng*>(r)->gen_f32(); }
void oracle_rng_fill_f32(void* r, float* out, uint64_t n) { auto* g = static_cast<StdRng*>(r); for (uint64_t i = 0; i < n; ++i) out[i] = g->gen_f32(); }
int oracle_rng_gen_bool(void* r) { return static_cast<StdRng*>(r)->gen_bool() ? 1 : 0; }
void oracle_rng_gen_seed(void* r, uint8_t* out32) { static_cast<StdRng*>(r)->gen_seed(out32); }
uint32_t oracle_rng_gen_range_u32_incl(void* r, uint32_t lo, uint32_t hi) { return static_cast<StdRng*>(r)->gen_range_u32_incl(lo, hi); }
uint64_t oracle_rng_gen_range_u64(void* r, uint64_t lo, uint64_t hi) { return static_cast<StdRng*>(r)->gen_range_u64(lo, hi); }
void oracle_rng_sample2(void* r, uint32_t length, uint32_t* out2) { static_cast<StdRng*>(r)->sample2(length, out2); }

// Synthetic benchmark matrix: element (i,j) = n-th gen::<f32>() of
// StdRng::from_seed(seed), n = i*d + j, minus `centre` (0 or 0.5). Counter based, so
// rows [row0, row0+rows) are generated without walking the stream from 0.
void oracle_synth_rows(const uint8_t* seed32, uint64_t d, uint64_t row0, uint64_t rows, float centre, float* out, int n_threads) {
    StdRng base = StdRng::from_seed(seed32);
    auto work = [&](uint64_t r0, uint64_t r1) {
        uint32_t blk[16];
        uint64_t w = (row0 + r0) * d, wend = (row0 + r1) * d;
        float* o = out + r0 * d;
        while (w < wend) {
            chacha12_block(base.key, w / 16, blk);
            for (uint64_t k = w % 16; k < 16 && w < wend; ++k, ++w) *o++ = (float)(blk[k] >> 8) * (1.0f / 16777216.0f) - centre;
        }
    };
    if (n_threads <= 1) { work(0, rows); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, rows * t / n_threads, rows * (t + 1) / n_threads);
    for (auto& x : th) x.join();
}

// ---- arithmetic ----------------------------------------------------------------------
float oracle_dot(const float* a, const float* b, uint64_t n) { return dot_product(a, b, n); }
float oracle_euclid(const float* a, const float* b, uint64_t n) { return euclidean_distance(a, b, n); }
float oracle_margin(int metric, const float* nv, float nh0, float nh1, const float* qv, float qh0, float qh1, uint64_t d) {
    return margin(metric, Leaf{nh0, nh1, nv}, Leaf{qh0, qh1, qv}, d);
}
float oracle_built_distance(int metric, const float* pv, float ph0, float ph1, const float* qv, float qh0, float qh1, uint64_t d) {
    return built_distance(metric, Leaf{ph0, ph1, pv}, Leaf{qh0, qh1, qv}, d);
}
float oracle_normalized_distance(int metric, float dist) { return normalized_distance(metric, dist); }
void oracle_new_header(int metric, const float* v, uint64_t d, float* out2) { new_header(metric, v, d, out2[0], out2[1]); }

// side()/margin over a row list — the loop of src/writer.rs:1201-1207
void oracle_side_batch(int metric, const float* nv, float nh0, float nh1, const float* vectors, const float* h0, const float* h1,
                       uint64_t d, const uint32_t* rows, uint64_t n_rows, uint8_t* out_side, float* out_margin) {
    Leaf nl{nh0, nh1, nv};
    for (uint64_t i = 0; i < n_rows; ++i) {
        uint32_t r = rows[i];
        float mg = margin(metric, nl, Leaf{h0 ? h0[r] : 0.f, h1 ? h1[r] : 0.f, vectors + (size_t)r * d}, d);
        if (out_margin) out_margin[i] = mg;
        out_side[i] = side_is_right(mg) ? 1 : 0;
    }
}

// D::create_split on a row subset; consumes the RNG exactly like the reference.
void oracle_create_split(int metric, void* rng, const float* vectors, const float* h0, const float* h1, uint64_t d,
                         const uint32_t* rows, uint32_t n_rows, float* out_normal, float* out_hdr2) {
    SubsetView sv{rows, n_rows, vectors, h0, h1, (size_t)d};
    OwnedLeaf nl;
    create_split(metric, *static_cast<StdRng*>(rng), sv, nl);
    memcpy(out_normal, nl.v.data(), 4 * d);
    out_hdr2[0] = nl.h0; out_hdr2[1] = nl.h1;
}

static uint64_t g_rerank_dims = 0;
// re-rank loop + top-k + normalized_distance — src/reader.rs:381-399
uint32_t oracle_rerank(int metric, const float* qv, float qh0, float qh1, const float* vectors, const float* h0, const float* h1,
                       uint64_t d, const uint32_t* rows, uint64_t n_rows, uint32_t count, uint32_t* out_rows, float* out_dist) {
    std::vector<std::pair<float, uint32_t>> dists;
    Leaf q{qh0, qh1, qv};
    for (uint64_t i = 0; i < n_rows; ++i) {
        uint32_t r = rows[i];
        dists.push_back({built_distance(metric, q, Leaf{h0 ? h0[r] : 0.f, h1 ? h1[r] : 0.f, vectors + (size_t)r * d}, d), r});
    }
    size_t k = std::min<size_t>(count, dists.size());
    auto top = Db::median_based_top_k(std::move(dists), k);
    for (size_t i = 0; i < top.size(); ++i) { out_rows[i] = top[i].second; out_dist[i] = normalized_distance(metric, top[i].first, g_rerank_dims ? g_rerank_dims : d); }
    return (uint32_t)top.size();
}
// binary-quantized indexes: the `dimensions` normalized_distance divides by (reader.rs:398) for the next oracle_rerank calls (0 = d)
void oracle_set_rerank_dims(uint64_t dims) { g_rerank_dims = dims; }
// BinaryQuantized::from_slice + ::iter: the +-1.0 values of the quantized vector, 64 * ceil(d / 64) of them
uint64_t oracle_bq_quantize(const float* v, uint64_t d, float* out) { bq_quantize_pm1(v, d, out); return bq_padded_dims(d); }
void oracle_db_set_user_dims(void* db, uint64_t dims) { static_cast<Db*>(db)->user_dims = dims; }

void oracle_dot_preprocess(const float* vectors, uint64_t n, uint64_t d, float* out_extra_dim, float* out_norm) {
    Db db(DOT_PRODUCT, d);
    std::vector<uint32_t> ids(n);
    for (uint64_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    db.ids = ids.data(); db.vec = vectors; db.n = n; db.borrowed = true;
    db.freeze(); db.preprocess();
    for (uint64_t i = 0; i < n; ++i) { out_extra_dim[i] = db.h0_own[i]; out_norm[i] = db.h1_own[i]; }
}

uint64_t oracle_target_n_trees(int64_t n_trees_opt, uint64_t dims, uint64_t n_items, uint64_t n_roots) { return target_n_trees(n_trees_opt, dims, n_items, n_roots); }
double oracle_split_imbalance(uint64_t l, uint64_t r) { return split_imbalance(l, r); }

// ---- database --------------------------------------------------------------------------
void* oracle_db_new(int metric, uint64_t d) { return new Db(metric, d); }
void oracle_db_free(void* db) { delete static_cast<Db*>(db); }
void oracle_db_add_item(void* db, uint32_t id, const float* v) { auto* D = static_cast<Db*>(db); D->staged[id] = std::vector<float>(v, v + D->d); D->mark_updated(id); }
int oracle_db_del_item(void* db, uint32_t id) { auto* D = static_cast<Db*>(db); int ex = (int)D->staged.erase(id); if (ex) D->mark_updated(id); return ex; }
int oracle_db_build_incremental(void* db, void* rng, int64_t n_trees_opt, uint64_t split_after) {
    try { static_cast<Db*>(db)->build_incremental(*static_cast<StdRng*>(rng), n_trees_opt, split_after); return 0; } catch (...) { return -1; }
}
// borrow caller-owned arrays (ids ascending); they must outlive the db
void oracle_db_set_items(void* db, uint64_t n, const uint32_t* ids, const float* vectors) {
    auto* D = static_cast<Db*>(db); D->staged.clear(); D->ids = ids; D->vec = vectors; D->n = n; D->borrowed = true;
}
int oracle_db_build(void* db, void* rng, int64_t n_trees_opt, uint64_t split_after, int n_threads) {
    try { static_cast<Db*>(db)->build(*static_cast<StdRng*>(rng), n_trees_opt, split_after, n_threads); return 0; } catch (...) { return -1; }
}
int oracle_db_build_memory_limited(void* db, void* rng, int64_t n_trees_opt, uint64_t split_after, uint64_t available_memory) {
    try { static_cast<Db*>(db)->build_memory_limited(*static_cast<StdRng*>(rng), n_trees_opt, split_after, available_memory); return 0; } catch (...) { return -1; }
}
uint64_t oracle_db_n_nodes(void* db) { return static_cast<Db*>(db)->nodes.size(); }
uint64_t oracle_db_n_roots(void* db) { return static_cast<Db*>(db)->roots.size(); }
void oracle_db_roots(void* db, uint32_t* out) { auto* D = static_cast<Db*>(db); memcpy(out, D->roots.data(), 4 * D->roots.size()); }
uint64_t oracle_db_scanned_rows(void* db) { return static_cast<Db*>(db)->scanned_rows.load(); }
void oracle_db_item_header(void* db, uint32_t id, float* out2) {
    auto* D = static_cast<Db*>(db); int64_t r = D->row_of(id); out2[0] = out2[1] = 0.f;
    if (r >= 0) { Leaf l = D->leaf((size_t)r); out2[0] = l.h0; out2[1] = l.h1; }
}
typedef void (*oracle_node_sink)(void* arg, uint32_t node_id, const uint8_t* bytes, uint64_t len);
void oracle_db_emit_nodes(void* db, oracle_node_sink sink, void* arg) {
    auto* D = static_cast<Db*>(db);
    std::vector<uint8_t> buf;
    for (size_t id = 0; id < D->nodes.size(); ++id) { if (D->nodes[id].kind == 0) continue; encode_tree_node(D->metric, D->d, D->nodes[id], buf); sink(arg, (uint32_t)id, buf.data(), buf.size()); }
}
// returns number of results, -1 if the item does not exist. out_cand (optional, cap
// cand_cap) receives the deduplicated candidate ids that entered the re-rank loop.
static int64_t nns_common(Db* D, const Leaf& q, uint64_t count, uint64_t search_k, uint64_t oversampling, const uint32_t* cand, int64_t n_cand,
                          uint32_t* out_ids, float* out_dist, uint32_t* out_cand, uint64_t cand_cap, uint64_t* out_n_cand) {
    std::vector<uint32_t> cv, used;
    if (n_cand >= 0) { cv.assign(cand, cand + n_cand); std::sort(cv.begin(), cv.end()); }
    auto res = D->nns_by_leaf(q, count, search_k, oversampling, n_cand >= 0 ? &cv : nullptr, &used);
    for (size_t i = 0; i < res.size(); ++i) { out_ids[i] = res[i].first; out_dist[i] = res[i].second; }
    if (out_n_cand) *out_n_cand = used.size();
    if (out_cand) memcpy(out_cand, used.data(), 4 * std::min<uint64_t>(cand_cap, used.size()));
    return (int64_t)res.size();
}
int64_t oracle_db_nns_by_item(void* db, uint32_t item, uint64_t count, uint64_t search_k, uint64_t oversampling, const uint32_t* cand, int64_t n_cand,
                              uint32_t* out_ids, float* out_dist, uint32_t* out_cand, uint64_t cand_cap, uint64_t* out_n_cand) {
    auto* D = static_cast<Db*>(db);
    int64_t r = D->row_of(item);
    if (r < 0) return -1;
    return nns_common(D, D->leaf((size_t)r), count, search_k, oversampling, cand, n_cand, out_ids, out_dist, out_cand, cand_cap, out_n_cand);
}
int64_t oracle_db_nns_by_vector(void* db, const float* v, uint64_t count, uint64_t search_k, uint64_t oversampling, const uint32_t* cand, int64_t n_cand,
                                uint32_t* out_ids, float* out_dist, uint32_t* out_cand, uint64_t cand_cap, uint64_t* out_n_cand) {
    auto* D = static_cast<Db*>(db);
    Leaf q{0.f, 0.f, v};
    new_header(D->metric, v, D->d, q.h0, q.h1);  // reader.rs:72-73
    return nns_common(D, q, count, search_k, oversampling, cand, n_cand, out_ids, out_dist, out_cand, cand_cap, out_n_cand);
}

}  // extern "C"
