// ORACLE — TEST INFRASTRUCTURE ONLY (see rng.hpp header).
//
// CPU restatement of the reference's host control flow around the distance kernels:
//   src/writer.rs:487-629     Writer::build (fresh-build path)
//   src/writer.rs:964-976 + src/distance/dot_product.rs:119-165   DotProduct::preprocess
//   src/writer.rs:474-477     fit_in_descendant
//   src/writer.rs:1167-1261   make_tree_in_file
//   src/writer.rs:1310-1326   randomly_split_children
//   src/writer.rs:1348-1353   split_imbalance
//   src/writer.rs:1358-1394   target_n_trees
//   src/parallel.rs:207-255   ConcurrentNodeIds (fresh build: plain counter)
//   src/reader.rs:317-401     Reader::nns_by_leaf
//   src/reader.rs:607-640     median_based_top_k
//   src/node.rs:218-241       NodeCodec::bytes_encode   (+ roaring 0.10.9 serialize_into,
//                             published portable format, cookie 12346, no run containers)
// Storage (LMDB/heed) is replaced by in-memory tables; node ids follow the order a
// 1-thread rayon pool produces (tree tasks run LIFO: SURVEY.md Appendix B.4), which
// is what the reference's snapshot tests pin (src/tests/mod.rs:94).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <functional>
#include <map>
#include <queue>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "distance.hpp"

namespace oracle {

struct TreeNode {
    uint8_t kind = 0;  // 1 = Descendants, 2 = SplitPlaneNormal
    uint32_t left = 0, right = 0;
    bool has_normal = false;
    OwnedLeaf normal;
    std::vector<uint32_t> descendants;  // item ids, ascending
};

static inline double split_imbalance(uint64_t l, uint64_t r) {  // writer.rs:1348-1353
    double ls = (double)l, rs = (double)r;
    double f = ls / (ls + rs + 2.220446049250313e-16);
    return f > 1.0 - f ? f : 1.0 - f;  // f64::max
}

static inline uint64_t target_n_trees(int64_t n_trees_opt, uint64_t dimensions, uint64_t n_items, uint64_t n_roots) {
    if (n_trees_opt >= 0) return (uint64_t)n_trees_opt;  // writer.rs:1364-1365
    double nb_vec = (double)n_items;
    double nb_trees;
    if (nb_vec < 10000.0) nb_trees = std::pow(2.0, std::log2(nb_vec) - 6.0);
    else nb_trees = std::pow(2.0, std::log10(nb_vec) + std::log10((double)dimensions) + std::pow(768.0 / (double)dimensions, 4.0));
    double c = std::ceil(nb_trees);
    uint64_t n;  // Rust `as u64` saturates, NaN -> 0
    if (!(c == c) || c <= 0.0) n = 0; else if (c >= 18446744073709551615.0) n = UINT64_MAX; else n = (uint64_t)c;
    if (n_roots > n) {
        uint64_t to_remove = n_roots - n;
        if ((double)to_remove / (double)n < 0.20) n = n_roots;
    }
    return n;
}

// roaring 0.10.9 RoaringBitmap::serialize_into for an ascending id list.
static inline void roaring_serialize(const std::vector<uint32_t>& ids, std::vector<uint8_t>& out) {
    struct C { uint16_t key; size_t begin, end; };
    std::vector<C> cs;
    for (size_t i = 0; i < ids.size();) {
        uint16_t key = (uint16_t)(ids[i] >> 16);
        size_t j = i;
        while (j < ids.size() && (uint16_t)(ids[j] >> 16) == key) ++j;
        cs.push_back({key, i, j});
        i = j;
    }
    auto w16 = [&](uint16_t v) { out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); };
    auto w32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) out.push_back((uint8_t)(v >> (8 * k))); };
    w32(12346u);
    w32((uint32_t)cs.size());
    for (auto& c : cs) { w16(c.key); w16((uint16_t)(c.end - c.begin - 1)); }
    uint32_t offset = 8 + 8 * (uint32_t)cs.size();
    for (auto& c : cs) {
        w32(offset);
        size_t len = c.end - c.begin;
        offset += len > 4096 ? 8192u : (uint32_t)len * 2u;
    }
    for (auto& c : cs) {
        size_t len = c.end - c.begin;
        if (len > 4096) {
            std::vector<uint64_t> bits(1024, 0);
            for (size_t i = c.begin; i < c.end; ++i) { uint16_t lo = (uint16_t)ids[i]; bits[lo >> 6] |= 1ull << (lo & 63); }
            for (uint64_t b : bits) for (int k = 0; k < 8; ++k) out.push_back((uint8_t)(b >> (8 * k)));
        } else {
            for (size_t i = c.begin; i < c.end; ++i) w16((uint16_t)ids[i]);
        }
    }
}

// NodeCodec::bytes_encode for tree nodes — src/node.rs:229-241
static inline void encode_tree_node(int metric, size_t d, const TreeNode& n, std::vector<uint8_t>& out) {
    out.clear();
    if (n.kind == 1) {
        out.push_back(1);
        roaring_serialize(n.descendants, out);
        return;
    }
    out.push_back(2);
    for (int k = 3; k >= 0; --k) out.push_back((uint8_t)(n.left >> (8 * k)));
    for (int k = 3; k >= 0; --k) out.push_back((uint8_t)(n.right >> (8 * k)));
    if (n.has_normal) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&n.normal.h0);
        out.insert(out.end(), p, p + 4);
        if (header_floats(metric) == 2) { p = reinterpret_cast<const uint8_t*>(&n.normal.h1); out.insert(out.end(), p, p + 4); }
        if (is_bq(metric)) {   // the vector part is the bit string: 64-bit words, bit i of word w = element 64 w + i (binary_quantized.rs:80-92)
            for (size_t w = 0; w < d / 64; ++w) {
                uint64_t word = 0;
                for (size_t i = 0; i < 64; ++i) if (n.normal.v[64 * w + i] > 0.f) word |= 1ull << i;
                p = reinterpret_cast<const uint8_t*>(&word);
                out.insert(out.end(), p, p + 8);
            }
            return;
        }
        p = reinterpret_cast<const uint8_t*>(n.normal.v.data());
        out.insert(out.end(), p, p + 4 * d);
    }
}

struct Db {
    int metric;
    size_t d;
    size_t user_dims = 0;   // binary-quantized indexes: Reader::dimensions (d is the padded bit count); 0 = d
    // staging area of add_item (small tests)
    std::map<uint32_t, std::vector<float>> staged;
    // frozen item table (rows ascending by id)
    std::vector<uint32_t> ids_own;
    std::vector<float> vec_own, h0_own, h1_own;
    const uint32_t* ids = nullptr;
    const float* vec = nullptr;
    bool borrowed = false;   // set_items(): caller-owned arrays instead of the add_item staging area
    size_t n = 0;
    // forest
    std::vector<TreeNode> nodes;  // index = tree node id
    std::vector<uint32_t> roots;
    bool built = false;
    std::atomic<uint64_t> scanned_rows{0};  // rows that went through side() during build

    Db(int m, size_t dim) : metric(m), d(dim) {}

    int64_t row_of(uint32_t id) const {
        const uint32_t* p = std::lower_bound(ids, ids + n, id);
        if (p == ids + n || *p != id) return -1;
        return p - ids;
    }
    Leaf leaf(size_t row) const { return Leaf{h0_own[row], h1_own.empty() ? 0.f : h1_own[row], vec + row * d}; }

    void freeze() {
        if (!borrowed) {
            ids_own.clear(); vec_own.clear();
            for (auto& kv : staged) { ids_own.push_back(kv.first); vec_own.insert(vec_own.end(), kv.second.begin(), kv.second.end()); }
            ids = ids_own.data(); vec = vec_own.data(); n = ids_own.size();
        }
        // Writer::add_item: header = D::new_header(vector) — writer.rs:388-390
        h0_own.assign(n, 0.f);
        h1_own.assign(metric == DOT_PRODUCT ? n : 0, 0.f);
        if (metric == COSINE) for (size_t r = 0; r < n; ++r) h0_own[r] = norm_no_header(vec + r * d, d);
        if (metric == BQ_COSINE) for (size_t r = 0; r < n; ++r) { float h1; new_header(metric, vec + r * d, d, h0_own[r], h1); }
    }

    // DotProduct::preprocess — dot_product.rs:119-165
    void preprocess() {
        if (metric != DOT_PRODUCT) return;
        float max_norm = 0.0f;
        for (size_t r = 0; r < n; ++r) {
            float nm = norm_no_header(vec + r * d, d);
            max_norm = (nm != nm) ? max_norm : (max_norm != max_norm ? nm : (max_norm > nm ? max_norm : nm));  // f32::max
        }
        for (size_t r = 0; r < n; ++r) {
            float node_norm = norm_no_header(vec + r * d, d);
            float diff = (max_norm * max_norm) - (node_norm * node_norm);
            h1_own[r] = max_norm * max_norm;
            h0_own[r] = std::sqrt(diff);
        }
    }

    struct LocalNode {  // tree-local, post-order; children are local indices; the root is last
        uint8_t kind; uint32_t left, right; bool has_normal; OwnedLeaf normal; std::vector<uint32_t> rows;
    };

    // make_tree_in_file — writer.rs:1167-1261. Returns the local index of the node.
    // is_root: the root id was allocated up front (next_id = Some(..)), so it takes no
    // post-order number; it is appended last by the caller's convention.
    uint32_t make_tree(StdRng& rng, const std::vector<uint32_t>& rows, size_t K, std::vector<LocalNode>& out, uint64_t& scanned) {
        if (rows.size() <= K) {
            out.push_back(LocalNode{1, 0, 0, false, OwnedLeaf{}, rows});
            return (uint32_t)out.size() - 1;
        }
        SubsetView children{rows.data(), (uint32_t)rows.size(), vec, h0_own.data(), h1_own.empty() ? nullptr : h1_own.data(), d};
        std::vector<uint32_t> left, right;
        left.reserve(rows.size()); right.reserve(rows.size());
        int remaining_attempts = 3;
        OwnedLeaf normal;
        for (;;) {
            left.clear(); right.clear();
            create_split(metric, rng, children, normal);
            Leaf nl = normal.view();
            for (uint32_t r : rows) {
                if (side_is_right(margin(metric, nl, leaf(r), d))) right.push_back(r); else left.push_back(r);
            }
            scanned += rows.size();
            if (split_imbalance(left.size(), right.size()) < 0.95 || remaining_attempts == 0) break;
            --remaining_attempts;
        }
        bool has_normal = true;
        if (split_imbalance(left.size(), right.size()) > 0.99) {
            left.clear(); right.clear();  // randomly_split_children — writer.rs:1310-1326
            for (uint32_t r : rows) { if (rng.gen_bool()) left.push_back(r); else right.push_back(r); }
            has_normal = false;
        }
        uint32_t l = make_tree(rng, left, K, out, scanned);
        uint32_t r = make_tree(rng, right, K, out, scanned);
        LocalNode nd{2, l, r, has_normal, has_normal ? normal : OwnedLeaf{}, {}};
        out.push_back(std::move(nd));
        return (uint32_t)out.size() - 1;
    }

    // Writer::build, fresh-build path (no pre-existing trees) — writer.rs:487-629.
    // n_trees_opt < 0 => target_n_trees formula; split_after == 0 => dimensions.
    void build(StdRng& user_rng, int64_t n_trees_opt, size_t split_after, int n_threads) {
        freeze();
        preprocess();
        nodes.clear(); roots.clear(); scanned_rows = 0;
        const size_t K = split_after ? split_after : d;
        if (n <= K) {  // clear_db_and_create_a_single_leaf — writer.rs:916-962
            if (n > 0) {
                TreeNode t; t.kind = 1; t.descendants.assign(ids, ids + n);
                nodes.push_back(std::move(t));
                roots.push_back(0);
            }
            built = true; has_metadata = true; updated.clear();
            return;
        }
        uint64_t T = target_n_trees(n_trees_opt, d, n, 0);
        for (uint64_t t = 0; t < T; ++t) roots.push_back((uint32_t)t);  // writer.rs:556-561
        StdRng rng1 = user_rng.fork();                                   // writer.rs:575
        std::vector<StdRng> tree_rng;
        for (uint64_t t = 0; t < T; ++t) tree_rng.push_back(rng1.fork());  // writer.rs:795 (IntMap order = ascending roots)
        std::vector<uint32_t> all(n);
        for (size_t r = 0; r < n; ++r) all[r] = (uint32_t)r;
        std::vector<std::vector<LocalNode>> local(T);
        std::atomic<uint64_t> next{0};
        auto worker = [&]() {
            for (;;) {
                uint64_t t = next.fetch_add(1);
                if (t >= T) return;
                uint64_t sc = 0;
                make_tree(tree_rng[t], all, K, local[t], sc);
                scanned_rows += sc;
            }
        };
        if (n_threads <= 1) worker();
        else {
            std::vector<std::thread> th;
            for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
            for (auto& x : th) x.join();
        }
        // id assignment of a 1-thread rayon pool: spawned tree tasks run LIFO, ids come
        // from one counter starting after the roots, post-order inside each tree.
        std::vector<uint64_t> base(T);
        uint64_t counter = T;
        for (uint64_t k = 0; k < T; ++k) { uint64_t t = T - 1 - k; base[t] = counter; counter += local[t].size() - 1; }
        nodes.resize(counter);
        for (uint64_t t = 0; t < T; ++t) {
            auto& L = local[t];
            const uint32_t root_local = (uint32_t)L.size() - 1;
            auto gid = [&](uint32_t li) { return li == root_local ? (uint32_t)t : (uint32_t)(base[t] + li); };
            for (uint32_t li = 0; li < L.size(); ++li) {
                TreeNode& o = nodes[gid(li)];
                o.kind = L[li].kind;
                if (o.kind == 1) { o.descendants.resize(L[li].rows.size()); for (size_t i = 0; i < L[li].rows.size(); ++i) o.descendants[i] = ids[L[li].rows[i]]; }
                else { o.left = gid(L[li].left); o.right = gid(L[li].right); o.has_normal = L[li].has_normal; o.normal = std::move(L[li].normal); }
            }
            L.clear(); L.shrink_to_fit();
        }
        built = true; has_metadata = true; updated.clear();
    }

    // ---------------------------------------------------------------------------------------------
    // Memory-limited build (available_memory set) on a 1-thread rayon pool — restated only to pin the
    // oracle against the reference's `..._with_little_memory.snap` (src/tests/writer.rs:1377-1391),
    // the one golden that exercises Cosine two_means / create_split / margin:
    //   src/writer.rs:660-739    incremental_index_large_descendant
    //   src/writer.rs:744-844    insert_descendants_in_file_and_spawn_tasks
    //   src/writer.rs:1463-1531  insert_items_in_descendants_from_tmpfile
    //   src/writer.rs:1536-1584  fit_in_memory
    // plus the third-party behaviour the node ids depend on: hashbrown (via
    // nohash::IntMap) insertion / growth / iteration order, rayon's LIFO order of scope.spawn.
    struct IntMapEmu {  // std HashMap<u32, _, BuildNoHashHasher>: identity hash, SwissTable slots
        std::vector<int64_t> keys;               // -1 = empty
        std::vector<std::vector<uint32_t>> vals;
        size_t items = 0;
        static size_t capacity_of(size_t buckets) { return buckets < 8 ? buckets - 1 : buckets / 8 * 7; }
        size_t find(uint32_t k) const {
            if (keys.empty()) return SIZE_MAX;
            size_t mask = keys.size() - 1, pos = k & mask;
            for (size_t i = 0; i < keys.size(); ++i) { size_t b = (pos + i) & mask; if (keys[b] == (int64_t)k) return b; if (keys[b] < 0) return SIZE_MAX; }
            return SIZE_MAX;
        }
        void raw_insert(uint32_t k, std::vector<uint32_t>&& v) {
            size_t mask = keys.size() - 1, pos = k & mask;
            for (size_t i = 0;; ++i) { size_t b = (pos + i) & mask; if (keys[b] < 0) { keys[b] = k; vals[b] = std::move(v); return; } }
        }
        std::vector<uint32_t>& entry(uint32_t k) {  // insert-or-get
            size_t f = find(k);
            if (f != SIZE_MAX) return vals[f];
            if (keys.empty() || items == capacity_of(keys.size())) {  // reserve_rehash(1)
                size_t cap = std::max(items + 1, keys.empty() ? (size_t)0 : capacity_of(keys.size()) + 1);
                size_t nb = cap < 4 ? 4 : (cap < 8 ? 8 : 1);
                if (nb == 1) { size_t adj = cap * 8 / 7; nb = 1; while (nb < adj) nb <<= 1; }
                std::vector<int64_t> ok = std::move(keys);
                std::vector<std::vector<uint32_t>> ov = std::move(vals);
                keys.assign(nb, -1); vals.assign(nb, {});
                for (size_t b = 0; b < ok.size(); ++b) if (ok[b] >= 0) raw_insert((uint32_t)ok[b], std::move(ov[b]));
            }
            raw_insert(k, {});
            ++items;
            return vals[find(k)];
        }
    };

    struct LmTask { StdRng rng; uint32_t id; std::vector<uint32_t> rows; };
    struct LmState {
        size_t K = 0, memory = 0;
        uint32_t counter = 0;
        // ConcurrentNodeIds::next — src/parallel.rs:238-254: exhaust the free list first, then count up
        std::vector<uint32_t> available;
        size_t select_in_bitmap = 0;
        bool look_into_bitmap = false;
        uint32_t next_id() {
            if (look_into_bitmap) {
                if (select_in_bitmap < available.size()) return available[select_in_bitmap++];
                look_into_bitmap = false;
            }
            return counter++;
        }
        std::map<uint32_t, TreeNode> out;                 // final nodes (by id), rows still as row indices
        std::map<uint32_t, LocalNode> tmp_splits;         // the (single) thread's TmpNodes file: split nodes
        std::vector<LmTask> stack;                         // rayon local deque, popped LIFO
    };

    // fit_in_memory — writer.rs:1536-1584 (page_size = 4096)
    bool lm_fit_in_memory(LmState& S, std::vector<uint32_t>& to_insert, StdRng& rng, std::vector<uint32_t>& out) {
        out.clear();
        if (to_insert.empty()) return false;
        if (to_insert.size() <= d) { out.swap(to_insert); return true; }
        const size_t page_size = 4096;
        size_t nb_page_allowed = (size_t)std::floor((double)S.memory / (double)page_size);
        size_t largest_item_size = 4 * (size_t)header_floats(metric) + 4 * d;   // D::size_of_item
        size_t nb_items_per_page = page_size / largest_item_size;
        size_t nb_page_per_item = (size_t)std::ceil((double)largest_item_size / (double)page_size);
        size_t nb_items = nb_items_per_page > 1 ? nb_page_allowed * nb_items_per_page : (nb_page_per_item > 1 ? nb_page_allowed / nb_page_per_item : nb_page_allowed);
        if (nb_items <= d) nb_items = d + 1;
        if (nb_items >= to_insert.size()) { out.swap(to_insert); return true; }
        for (size_t i = 0; i < nb_items; ++i) {
            uint64_t idx = rng.gen_range_u64(0, to_insert.size());
            uint32_t item = to_insert[idx];                  // RoaringBitmap::select(idx)
            out.insert(std::lower_bound(out.begin(), out.end(), item), item);
            to_insert.erase(to_insert.begin() + idx);
        }
        return true;
    }

    uint32_t lm_make_tree(LmState& S, StdRng& rng, const std::vector<uint32_t>& rows, IntMapEmu& descendants, int64_t next_id) {
        if (rows.size() <= S.K) {
            uint32_t id = next_id >= 0 ? (uint32_t)next_id : S.next_id();
            descendants.entry(id) = rows;   // descendants.insert(item_id, item_indices.clone())
            return id;
        }
        SubsetView children{rows.data(), (uint32_t)rows.size(), vec, h0_own.data(), h1_own.empty() ? nullptr : h1_own.data(), d};
        std::vector<uint32_t> left, right;
        int remaining_attempts = 3;
        OwnedLeaf normal;
        for (;;) {
            left.clear(); right.clear();
            create_split(metric, rng, children, normal);
            Leaf nl = normal.view();
            for (uint32_t r : rows) { if (side_is_right(margin(metric, nl, leaf(r), d))) right.push_back(r); else left.push_back(r); }
            if (split_imbalance(left.size(), right.size()) < 0.95 || remaining_attempts == 0) break;
            --remaining_attempts;
        }
        bool has_normal = true;
        if (split_imbalance(left.size(), right.size()) > 0.99) {
            left.clear(); right.clear();
            for (uint32_t r : rows) { if (rng.gen_bool()) left.push_back(r); else right.push_back(r); }
            has_normal = false;
        }
        uint32_t l = lm_make_tree(S, rng, left, descendants, -1);
        uint32_t r = lm_make_tree(S, rng, right, descendants, -1);
        uint32_t id = next_id >= 0 ? (uint32_t)next_id : S.next_id();
        S.tmp_splits[id] = LocalNode{2, l, r, has_normal, has_normal ? normal : OwnedLeaf{}, {}};
        return id;
    }

    // insert_items_in_descendants_from_tmpfile — writer.rs:1463-1531
    void lm_route(LmState& S, StdRng& rng, uint32_t node, const std::vector<uint32_t>& to_insert, IntMapEmu& descendants) {
        auto it = S.tmp_splits.find(node);
        if (it == S.tmp_splits.end()) {   // tmp_nodes.get(..) == None: a pending descendants entry
            std::vector<uint32_t>& dst = descendants.vals[descendants.find(node)];
            std::vector<uint32_t> merged;
            std::set_union(dst.begin(), dst.end(), to_insert.begin(), to_insert.end(), std::back_inserter(merged));
            dst.swap(merged);
            return;
        }
        const LocalNode& sp = it->second;
        std::vector<uint32_t> left, right;
        if (!sp.has_normal) { for (uint32_t r : to_insert) { if (rng.gen_bool()) left.push_back(r); else right.push_back(r); } }
        else { Leaf nl = sp.normal.view(); for (uint32_t r : to_insert) { if (side_is_right(margin(metric, nl, leaf(r), d))) right.push_back(r); else left.push_back(r); } }
        const uint32_t lid = sp.left, rid = sp.right;
        if (!left.empty()) lm_route(S, rng, lid, left, descendants);
        if (!right.empty()) lm_route(S, rng, rid, right, descendants);
    }

    // insert_descendants_in_file_and_spawn_tasks — writer.rs:744-844
    void lm_process_descendants(LmState& S, StdRng& rng, IntMapEmu& descendants) {
        for (size_t b = 0; b < descendants.keys.size(); ++b) {   // hashbrown iteration: ascending bucket
            if (descendants.keys[b] < 0) continue;
            uint32_t id = (uint32_t)descendants.keys[b];
            std::vector<uint32_t>& rows = descendants.vals[b];
            if (rows.size() <= S.K) { TreeNode t; t.kind = 1; t.descendants = rows; S.out[id] = std::move(t); }
            else { LmTask task{rng.fork(), id, rows}; S.stack.push_back(std::move(task)); }
        }
    }

    // incremental_index_large_descendant — writer.rs:660-739
    void lm_run_task(LmState& S, LmTask& task) {
        IntMapEmu descendants;
        std::vector<uint32_t> to_insert = task.rows, chunk;
        lm_fit_in_memory(S, to_insert, task.rng, chunk);
        lm_make_tree(S, task.rng, chunk, descendants, (int64_t)task.id);
        while (lm_fit_in_memory(S, to_insert, task.rng, chunk)) lm_route(S, task.rng, task.id, chunk, descendants);
        lm_process_descendants(S, task.rng, descendants);
    }

    void build_memory_limited(StdRng& user_rng, int64_t n_trees_opt, size_t split_after, size_t available_memory) {
        freeze();
        preprocess();
        nodes.clear(); roots.clear(); scanned_rows = 0;
        LmState S;
        S.K = split_after ? split_after : d;
        S.memory = available_memory;   // / current_num_threads() == 1
        if (n <= S.K) throw std::runtime_error("single-leaf case: use build()");
        uint64_t T = target_n_trees(n_trees_opt, d, n, 0);
        IntMapEmu top;
        std::vector<uint32_t> all(n);
        for (size_t r = 0; r < n; ++r) all[r] = (uint32_t)r;
        for (uint64_t t = 0; t < T; ++t) { uint32_t nid = S.next_id(); roots.push_back(nid); top.entry(nid) = all; }   // writer.rs:556-561
        StdRng rng1 = user_rng.fork();                                                                        // writer.rs:575
        lm_process_descendants(S, rng1, top);
        while (!S.stack.empty()) { LmTask t = std::move(S.stack.back()); S.stack.pop_back(); lm_run_task(S, t); }
        nodes.resize(S.counter);
        for (auto& kv : S.out) { TreeNode& o = nodes[kv.first]; o.kind = 1; o.descendants.resize(kv.second.descendants.size()); for (size_t i = 0; i < o.descendants.size(); ++i) o.descendants[i] = ids[kv.second.descendants[i]]; }
        for (auto& kv : S.tmp_splits) { TreeNode& o = nodes[kv.first]; o.kind = 2; o.left = kv.second.left; o.right = kv.second.right; o.has_normal = kv.second.has_normal; o.normal = kv.second.normal; }
        built = true;
    }

    // ---------------------------------------------------------------------------------------------
    // Writer::build on an index that already has trees (src/writer.rs:487-629), in the order a
    // 1-thread rayon pool executes it:
    //   delete_extra_trees :632-653, delete_tree :1263-1277
    //   delete_items_from_trees :978-1015, delete_items_in_file :1021-1114
    //   insert_items_in_current_trees :846-889, insert_items_in_tree :1119-1160,
    //   insert_items_in_descendants_from_frozen_reader :1398-1459
    //   ConcurrentNodeIds (free list first) src/parallel.rs:207-255
    // Pinned by the reference's incremental inline snapshots (tests/test_oracle_golden.py). With
    // several roots the reference merges the per-root results through rayon's `reduce`
    // (writer.rs:1148-1159); this restatement merges in ascending root order.
    std::set<uint32_t> updated;      // the `Updated` keys (src/node_id.rs:14-16)
    bool has_metadata = false;

    void mark_updated(uint32_t id) { updated.insert(id); }

    void inc_delete_tree(uint32_t node) {
        if (node >= nodes.size() || nodes[node].kind == 0) return;
        if (nodes[node].kind == 2) { inc_delete_tree(nodes[node].left); inc_delete_tree(nodes[node].right); }
        nodes[node] = TreeNode{};
    }

    struct TmpOps { std::vector<std::pair<uint32_t, TreeNode>> puts; std::set<uint32_t> deleted; };

    // returns (new id, Some(items) / None)
    std::pair<uint32_t, std::pair<bool, std::vector<uint32_t>>> inc_delete_items(uint32_t current, TmpOps& tmp, const std::set<uint32_t>& to_delete, size_t K) {
        const TreeNode& nd = nodes[current];
        if (nd.kind == 1) {
            std::vector<uint32_t> nw;
            for (uint32_t id : nd.descendants) if (!to_delete.count(id)) nw.push_back(id);
            if (nw.size() != nd.descendants.size()) { TreeNode t; t.kind = 1; t.descendants = nw; tmp.puts.push_back({current, std::move(t)}); }
            return {current, {true, nw}};
        }
        const uint32_t left = nd.left, right = nd.right;
        auto L = inc_delete_items(left, tmp, to_delete, K);
        auto R = inc_delete_items(right, tmp, to_delete, K);
        const uint32_t new_left = L.first, new_right = R.first;
        const bool ls = L.second.first, rs = R.second.first;
        auto put_split = [&]() {
            if (new_left != left || new_right != right) { TreeNode t = nodes[current]; t.left = new_left; t.right = new_right; tmp.puts.push_back({current, std::move(t)}); }
        };
        if (ls && L.second.second.empty()) { tmp.deleted.insert(new_left); tmp.deleted.insert(current); return {new_right, R.second}; }
        if (rs && R.second.second.empty()) { tmp.deleted.insert(new_right); tmp.deleted.insert(current); return {new_left, L.second}; }
        if (ls && rs) {
            size_t total = L.second.second.size() + R.second.second.size();
            if (total <= K) {
                std::vector<uint32_t> all;
                std::set_union(L.second.second.begin(), L.second.second.end(), R.second.second.begin(), R.second.second.end(), std::back_inserter(all));
                tmp.deleted.insert(new_left); tmp.deleted.insert(new_right);
                TreeNode t; t.kind = 1; t.descendants = all;
                tmp.puts.push_back({current, std::move(t)});
                return {current, {true, all}};
            }
            put_split();
            return {current, {false, {}}};
        }
        put_split();
        return {current, {false, {}}};
    }

    void inc_route(StdRng& rng, uint32_t node, const std::vector<uint32_t>& to_insert /* item ids */, IntMapEmu& out) {
        const TreeNode& nd = nodes[node];
        if (nd.kind == 1) {
            std::vector<uint32_t> merged;
            std::set_union(nd.descendants.begin(), nd.descendants.end(), to_insert.begin(), to_insert.end(), std::back_inserter(merged));
            out.entry(node) = merged;   // descendants_to_update.insert(current_node, descendants | to_insert)
            return;
        }
        std::vector<uint32_t> left, right;
        if (!nd.has_normal) { for (uint32_t id : to_insert) { if (rng.gen_bool()) left.push_back(id); else right.push_back(id); } }
        else { Leaf nl = nd.normal.view(); for (uint32_t id : to_insert) { if (side_is_right(margin(metric, nl, leaf((size_t)row_of(id)), d))) right.push_back(id); else left.push_back(id); } }
        const uint32_t l = nd.left, r = nd.right;
        if (!left.empty()) inc_route(rng, l, left, out);
        if (!right.empty()) inc_route(rng, r, right, out);
    }

    void build_incremental(StdRng& user_rng, int64_t n_trees_opt, size_t split_after) {
        freeze();
        preprocess();
        const size_t K = split_after ? split_after : d;
        std::set<uint32_t> updated_items;
        updated_items.swap(updated);                       // reset_and_retrieve_updated_items
        if (n <= K) {                                      // clear_db_and_create_a_single_leaf
            nodes.clear(); roots.clear();
            if (n > 0) { TreeNode t; t.kind = 1; t.descendants.assign(ids, ids + n); nodes.push_back(std::move(t)); roots.push_back(0); }
            has_metadata = true; built = true;
            return;
        }
        const std::set<uint32_t>& to_delete = updated_items;
        std::vector<uint32_t> to_insert;
        for (uint32_t id : updated_items) if (row_of(id) >= 0) to_insert.push_back(id);
        if (!has_metadata) roots.clear();
        LmState S;
        S.K = K; S.memory = SIZE_MAX;
        {   // ConcurrentNodeIds::new(used_tree_node) — before anything is deleted
            uint32_t last = 0; bool any = false;
            for (uint32_t i = 0; i < nodes.size(); ++i) if (nodes[i].kind) { last = i; any = true; }
            uint32_t last_id = any ? last + 1 : 0;
            for (uint32_t i = 0; i < last_id; ++i) if (!nodes[i].kind) S.available.push_back(i);
            S.counter = last_id;
            S.look_into_bitmap = !S.available.empty();
        }
        uint64_t target = target_n_trees(n_trees_opt, d, n, roots.size());
        {   // delete_extra_trees
            size_t extraneous = roots.size() > target ? roots.size() - (size_t)target : 0;
            for (size_t i = 0; i < extraneous && !roots.empty(); ++i) { uint32_t r0 = roots[0]; roots[0] = roots.back(); roots.pop_back(); inc_delete_tree(r0); }
        }
        {   // delete_items_from_trees
            TmpOps tmp;
            for (uint32_t& root : roots) { auto res = inc_delete_items(root, tmp, to_delete, K); root = res.first; }
            std::sort(roots.begin(), roots.end());
            for (uint32_t id : tmp.deleted) nodes[id] = TreeNode{};
            for (auto& pr : tmp.puts) if (!tmp.deleted.count(pr.first)) nodes[pr.first] = pr.second;
        }
        // insert_items_in_current_trees. The per-root results are combined by rayon's
        // `repeat_n(..).zip(roots).map(..).reduce(..)` (writer.rs:1128-1159); on a 1-thread pool the
        // length splitter splits the roots exactly once, at len / 2: each half is folded left to
        // right into its own map, then the right map is merged into the left one, and the result is
        // re-inserted into a fresh map (writer.rs:879-887) — every step in hashbrown iteration order.
        IntMapEmu top_ids;
        if (!roots.empty() && !to_insert.empty()) {
            uint64_t seed = user_rng.next_u64();                       // repeat_n(rng.next_u64(), roots.len())
            auto merge_into = [](IntMapEmu& dst, IntMapEmu& src) {
                for (size_t b = 0; b < src.keys.size(); ++b) {
                    if (src.keys[b] < 0) continue;
                    std::vector<uint32_t>& dv = dst.entry((uint32_t)src.keys[b]);
                    std::vector<uint32_t> merged;
                    std::set_union(dv.begin(), dv.end(), src.vals[b].begin(), src.vals[b].end(), std::back_inserter(merged));
                    dv.swap(merged);
                }
            };
            auto fold = [&](size_t a, size_t b) {
                IntMapEmu acc;
                for (size_t i = a; i < b; ++i) {
                    StdRng rr = StdRng::seed_from_u64(seed + (uint64_t)roots[i]);
                    IntMapEmu per_root;
                    inc_route(rr, roots[i], to_insert, per_root);
                    merge_into(acc, per_root);
                }
                return acc;
            };
            IntMapEmu reduced;
            if (roots.size() >= 2) { size_t mid = roots.size() / 2; reduced = fold(0, mid); IntMapEmu right = fold(mid, roots.size()); merge_into(reduced, right); }
            else reduced = fold(0, roots.size());
            merge_into(top_ids, reduced);
        }
        uint64_t nb_missing = target > roots.size() ? target - roots.size() : 0;
        std::vector<uint32_t> all_ids(ids, ids + n);
        for (uint64_t i = 0; i < nb_missing; ++i) { uint32_t nid = S.next_id(); roots.push_back(nid); top_ids.entry(nid) = all_ids; }
        // ids -> rows for the tree builder
        IntMapEmu top = top_ids;
        for (auto& v : top.vals) for (auto& x : v) x = (uint32_t)row_of(x);
        StdRng rng1 = user_rng.fork();                                  // writer.rs:575
        lm_process_descendants(S, rng1, top);
        while (!S.stack.empty()) { LmTask t = std::move(S.stack.back()); S.stack.pop_back(); lm_run_task(S, t); }
        if (nodes.size() < S.counter) nodes.resize(S.counter);
        for (auto& kv : S.out) { TreeNode& o = nodes[kv.first]; o = TreeNode{}; o.kind = 1; o.descendants.resize(kv.second.descendants.size()); for (size_t i = 0; i < o.descendants.size(); ++i) o.descendants[i] = ids[kv.second.descendants[i]]; }
        for (auto& kv : S.tmp_splits) { TreeNode& o = nodes[kv.first]; o = TreeNode{}; o.kind = 2; o.left = kv.second.left; o.right = kv.second.right; o.has_normal = kv.second.has_normal; o.normal = kv.second.normal; }
        has_metadata = true; built = true;
    }

    // Total order of (OrderedFloat<f32>, u32): NaN greatest & all NaN equal, -0 == +0.
    static bool less_key(float a, uint32_t ia, float b, uint32_t ib) {
        bool an = a != a, bn = b != b;
        if (an || bn) { if (an && bn) return ia < ib; return bn; }
        if (a < b) return true;
        if (a > b) return false;
        return ia < ib;
    }

    // median_based_top_k — reader.rs:607-640 (restated literally; equals sort+truncate)
    static std::vector<std::pair<float, uint32_t>> median_based_top_k(std::vector<std::pair<float, uint32_t>> v, size_t k) {
        auto lt = [](const std::pair<float, uint32_t>& x, const std::pair<float, uint32_t>& y) { return less_key(x.first, x.second, y.first, y.second); };
        std::pair<float, uint32_t> threshold{FLT_MAX, UINT32_MAX};
        std::vector<std::pair<float, uint32_t>> buffer;
        if (k == 0) return buffer;  // (the reference would panic in select_nth_unstable; unreachable from nns_by_leaf unless count == 0)
        buffer.reserve(2 * std::max<size_t>(k, 1));
        size_t i = 0;
        for (; i < v.size() && i < 2 * k; ++i) buffer.push_back(v[i]);
        for (; i < v.size(); ++i) {
            if (!lt(v[i], threshold)) continue;
            if (buffer.size() == 2 * k) {
                std::nth_element(buffer.begin(), buffer.begin() + (k - 1), buffer.end(), lt);
                threshold = buffer[k - 1];
                buffer.resize(k);
            }
            buffer.push_back(v[i]);
        }
        std::sort(buffer.begin(), buffer.end(), lt);
        if (buffer.size() > k) buffer.resize(k);
        return buffer;
    }

    // Reader::nns_by_leaf — reader.rs:317-401.  candidates == nullptr => no filter.
    // If out_candidates is given, the deduplicated candidate id list that goes into the
    // re-rank loop (reader.rs:378-379) is stored there.
    std::vector<std::pair<uint32_t, float>> nns_by_leaf(const Leaf& q, size_t count, size_t search_k_opt, size_t oversampling_opt,
                                                        const std::vector<uint32_t>* candidates, std::vector<uint32_t>* out_candidates = nullptr) const {
        std::vector<std::pair<uint32_t, float>> output;
        if (n == 0) return output;
        size_t search_k = search_k_opt ? search_k_opt : count * roots.size();
        {   // saturating_mul
            size_t mul = oversampling_opt ? oversampling_opt : 1;  // DEFAULT_OVERSAMPLING = 1 (mod.rs:41)
            unsigned __int128 p = (unsigned __int128)search_k * mul;
            search_k = p > (unsigned __int128)SIZE_MAX ? SIZE_MAX : (size_t)p;
        }
        struct QE { float dist; uint32_t node; };
        auto cmp = [](const QE& a, const QE& b) { return less_key(a.dist, a.node, b.dist, b.node); };  // max-heap
        std::priority_queue<QE, std::vector<QE>, decltype(cmp)> queue(cmp);
        for (uint32_t r : roots) queue.push(QE{INFINITY, r});
        std::vector<uint32_t> nns;
        while (nns.size() < search_k) {
            if (queue.empty()) break;
            QE top = queue.top(); queue.pop();
            const TreeNode& node = nodes[top.node];
            if (node.kind == 1) {
                if (candidates) {
                    for (uint32_t id : node.descendants) if (std::binary_search(candidates->begin(), candidates->end(), id)) nns.push_back(id);
                } else nns.insert(nns.end(), node.descendants.begin(), node.descendants.end());
            } else {
                float mg = node.has_normal ? margin(metric, node.normal.view(), q, d) : 0.0f;
                queue.push(QE{pq_distance(top.dist, mg, false), node.left});
                queue.push(QE{pq_distance(top.dist, mg, true), node.right});
            }
        }
        std::sort(nns.begin(), nns.end());
        nns.erase(std::unique(nns.begin(), nns.end()), nns.end());
        if (out_candidates) *out_candidates = nns;
        std::vector<std::pair<float, uint32_t>> dists;
        dists.reserve(nns.size());
        for (uint32_t id : nns) {
            int64_t row = row_of(id);
            dists.push_back({built_distance(metric, q, leaf((size_t)row), d), id});
        }
        size_t k = std::min(count, dists.size());
        auto top = median_based_top_k(std::move(dists), k);
        for (auto& pr : top) output.push_back({pr.second, normalized_distance(metric, pr.first, user_dims ? user_dims : d)});
        return output;
    }
};

}  // namespace oracle
