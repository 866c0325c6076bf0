// Host-side construction of the device BVH and of the reference-order rank tables (see lgb_build.hpp).
#include "lgb_build.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

namespace lgb {
namespace {

struct Item { float lo[3], hi[3], c[3]; uint32_t type, index, pad; };   // 48 B

struct Box3 {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -std::numeric_limits<float>::infinity(); } }
    void grow(const float* l, const float* h) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); } }
    void grow(const Box3& b) { grow(b.lo, b.hi); }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dx * dz + dy * dz;
    }
};

constexpr int NB = 16;
struct Bins { Box3 box[3][NB]; uint32_t cnt[3][NB]; void reset() { for (int a = 0; a < 3; a++) for (int b = 0; b < NB; b++) { box[a][b].reset(); cnt[a][b] = 0; } } };

struct Builder {
    Item* items;
    HostNode* nodes;
    std::atomic<uint32_t> node_count{0};
    std::atomic<uint32_t> type_count[3];
    std::vector<uint32_t>* order;
    std::atomic<int> spare_threads{0};
    std::atomic<uint32_t> max_depth{0};
    int total_threads = 1;

    static void bin_range(const Item* it, size_t n, const float cmin[3], const float scale[3], Bins& bins) {
        for (size_t i = 0; i < n; i++) {
            for (int a = 0; a < 3; a++) {
                int k = (int)((it[i].c[a] - cmin[a]) * scale[a]);
                k = k < 0 ? 0 : (k >= NB ? NB - 1 : k);
                bins.box[a][k].grow(it[i].lo, it[i].hi);
                bins.cnt[a][k]++;
            }
        }
    }

    uint32_t make_leaf(size_t b, size_t e) {
        const uint32_t type = items[b].type, n = (uint32_t)(e - b);
        const uint32_t first = type_count[type].fetch_add(n);
        for (uint32_t i = 0; i < n; i++) order[type][first + i] = items[b + i].index;
        return kLeafBit | (type << 29) | ((n - 1) << 24) | first;
    }

    // Builds the subtree over items[b, e); returns its child word, `box` = its bounds.
    uint32_t build(size_t b, size_t e, Box3& box, uint32_t depth) {
        const size_t n = e - b;
        box.reset();
        Box3 cb; cb.reset();
        bool mixed = false;
        for (size_t i = b; i < e; i++) {
            box.grow(items[i].lo, items[i].hi);
            cb.grow(items[i].c, items[i].c);
            mixed |= items[i].type != items[b].type;
        }
        if (n <= (size_t)kMaxLeaf && !mixed) {
            uint32_t d = max_depth.load(std::memory_order_relaxed);
            while (depth > d && !max_depth.compare_exchange_weak(d, depth)) {}
            return make_leaf(b, e);
        }
        size_t mid = b;
        if (n <= (size_t)kMaxLeaf) {   // small but mixed types: split at the first type boundary
            std::stable_sort(items + b, items + e, [](const Item& x, const Item& y) { return x.type < y.type; });
            mid = b + 1;
            while (items[mid].type == items[b].type) mid++;
        } else {
            float scale[3]; bool any_axis = false;
            for (int a = 0; a < 3; a++) {
                float ext = cb.hi[a] - cb.lo[a];
                scale[a] = ext > 0.0f ? (float)NB * (1.0f - 1e-6f) / ext : 0.0f;
                any_axis |= ext > 0.0f;
            }
            int best_axis = -1, best_bin = -1;
            if (any_axis && depth < 40) {
                Bins bins; bins.reset();
                const int want = (n > (1u << 17)) ? std::min<int>(total_threads, 8) : 1;
                if (want > 1) {
                    std::vector<Bins> part(want);
                    std::vector<std::thread> th;
                    const size_t chunk = (n + want - 1) / want;
                    for (int t = 0; t < want; t++)
                        th.emplace_back([&, t] { part[t].reset(); size_t s = b + t * chunk, f = std::min(e, s + chunk); if (s < f) bin_range(items + s, f - s, cb.lo, scale, part[t]); });
                    for (auto& t : th) t.join();
                    for (int t = 0; t < want; t++)
                        for (int a = 0; a < 3; a++) for (int k = 0; k < NB; k++) { bins.box[a][k].grow(part[t].box[a][k]); bins.cnt[a][k] += part[t].cnt[a][k]; }
                } else {
                    bin_range(items + b, n, cb.lo, scale, bins);
                }
                float best_cost = std::numeric_limits<float>::infinity();
                for (int a = 0; a < 3; a++) {
                    if (scale[a] == 0.0f) continue;
                    float la[NB]; uint32_t lc[NB];
                    Box3 acc; acc.reset(); uint32_t c = 0;
                    for (int k = 0; k < NB; k++) { acc.grow(bins.box[a][k]); c += bins.cnt[a][k]; la[k] = c ? acc.half_area() : 0.0f; lc[k] = c; }
                    acc.reset(); c = 0;
                    for (int k = NB - 1; k >= 1; k--) {
                        acc.grow(bins.box[a][k]); c += bins.cnt[a][k];
                        if (!c || !lc[k - 1]) continue;
                        float cost = la[k - 1] * (float)lc[k - 1] + acc.half_area() * (float)c;
                        if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k - 1; }
                    }
                }
            }
            if (best_axis >= 0) {
                const int a = best_axis; const float cm = cb.lo[a], sc = scale[a];
                Item* m = std::partition(items + b, items + e, [&](const Item& it) {
                    int k = (int)((it.c[a] - cm) * sc);
                    k = k < 0 ? 0 : (k >= NB ? NB - 1 : k);
                    return k <= best_bin;
                });
                mid = (size_t)(m - items);
            }
            if (mid == b || mid == e) {      // coincident centroids or depth guard: median split along the widest axis
                int a = 0;
                for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
                mid = b + n / 2;
                std::nth_element(items + b, items + mid, items + e, [a](const Item& x, const Item& y) { return x.c[a] < y.c[a]; });
            }
        }
        const uint32_t me = node_count.fetch_add(1);
        Box3 lb, rb; uint32_t lw, rw;
        const bool spawn = (mid - b) > 8192 && (e - mid) > 8192 && spare_threads.fetch_sub(1) > 0;
        if (spawn) {
            std::thread t([&] { lw = build(b, mid, lb, depth + 1); });
            rw = build(mid, e, rb, depth + 1);
            t.join();
            spare_threads.fetch_add(1);
        } else {
            if ((mid - b) > 8192 && (e - mid) > 8192) spare_threads.fetch_add(1);   // undo the failed reservation
            lw = build(b, mid, lb, depth + 1);
            rw = build(mid, e, rb, depth + 1);
        }
        HostNode& nd = nodes[me];
        for (int k = 0; k < 3; k++) { nd.v[k] = lb.lo[k]; nd.v[3 + k] = lb.hi[k]; nd.v[6 + k] = rb.lo[k]; nd.v[9 + k] = rb.hi[k]; }
        nd.c0 = lw; nd.c1 = rw; nd.pad0 = nd.pad1 = 0;
        return me;
    }
};

}  // namespace

int build_sah(std::vector<PrimBox>& prims, float pad, int threads, BuiltBVH& out) {
    auto t0 = std::chrono::steady_clock::now();
    const size_t n = prims.size();
    if (n == 0) return -1;
    std::vector<Item> items(n);
    size_t per_type[3] = {0, 0, 0};
    for (size_t i = 0; i < n; i++) {
        Item& it = items[i];
        for (int k = 0; k < 3; k++) {
            it.lo[k] = std::nextafterf(prims[i].lo[k] - pad, -INFINITY);
            it.hi[k] = std::nextafterf(prims[i].hi[k] + pad, INFINITY);
            it.c[k] = 0.5f * prims[i].lo[k] + 0.5f * prims[i].hi[k];
        }
        it.type = prims[i].type; it.index = prims[i].index; it.pad = 0;
        if (it.type > 2) return -1;
        per_type[it.type]++;
    }
    for (int t = 0; t < 3; t++) { if (per_type[t] > kLeafFirstMask) return -2; out.order[t].assign(per_type[t], 0); }
    out.nodes.assign(std::max<size_t>(n, 2), HostNode{});
    Builder B;
    B.items = items.data(); B.nodes = out.nodes.data(); B.order = out.order;
    for (auto& c : B.type_count) c.store(0);
    B.total_threads = std::max(1, threads);
    B.spare_threads.store(std::max(0, threads - 1));
    Box3 box;
    bool homogeneous = true;
    for (size_t i = 1; i < n && homogeneous; i++) homogeneous = items[i].type == items[0].type;
    if (n <= (size_t)kMaxLeaf && homogeneous) {      // tiny scene: the root's two children are the same leaf (testing it twice changes nothing)
        B.node_count.store(1);
        uint32_t w = B.build(0, n, box, 1);
        HostNode& nd = out.nodes[0];
        for (int k = 0; k < 3; k++) { nd.v[k] = box.lo[k]; nd.v[3 + k] = box.hi[k]; nd.v[6 + k] = box.lo[k]; nd.v[9 + k] = box.hi[k]; }
        nd.c0 = w; nd.c1 = w; nd.pad0 = nd.pad1 = 0;
    } else {
        uint32_t w = B.build(0, n, box, 0);
        if (w != 0) return -3;                       // the root must be node 0
    }
    out.nodes.resize(B.node_count.load());
    out.max_depth = B.max_depth.load();
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

bool build_rank_tables(const lgb_scene_desc* d, uint32_t prim_count, int threads, std::vector<uint32_t>& rank) {
    rank.assign((size_t)8 * prim_count, 0xFFFFFFFFu);
    std::atomic<bool> ok{true};
    auto one = [&](int oct) {
        uint32_t* r = rank.data() + (size_t)oct * prim_count;
        uint32_t counter = 0;
        // explicit DFS: entries are node indices; a leaf expands its primitives in order, recursing into
        // nested BVHs immediately (bvh.rs:483-488 calls the nested intersect before the next primitive).
        struct Frame { uint32_t node; uint32_t next_ref; };
        std::vector<Frame> st;
        st.push_back({0, 0});
        while (!st.empty()) {
            Frame fr = st.back(); st.pop_back();
            const lgb_node& n = d->nodes[fr.node];
            if (n.b & LGB_LEAF_FLAG) {
                const uint32_t cnt = n.b & ~LGB_LEAF_FLAG;
                uint32_t i = fr.next_ref;
                for (; i < cnt; i++) {
                    const uint32_t ref = d->prim_refs[n.a + i], type = ref >> 30, idx = ref & 0x3FFFFFFFu;
                    uint32_t id;
                    if (type == LGB_PRIM_INSTANCE) {
                        st.push_back({fr.node, i + 1});                       // resume this leaf afterwards
                        st.push_back({d->instances[idx].root_node, 0});
                        break;
                    }
                    id = type == LGB_PRIM_SPHERE ? d->sphere_id[idx] : type == LGB_PRIM_CUBOID ? d->cuboid_id[idx] : d->triangle_id[idx];
                    if (id >= prim_count || r[id] != 0xFFFFFFFFu) { ok.store(false); return; }
                    r[id] = counter++;
                }
            } else {
                const bool neg = (oct >> n.b) & 1;                             // bvh.rs:496-503
                const uint32_t first = neg ? n.a : fr.node + 1, second = neg ? fr.node + 1 : n.a;
                st.push_back({second, 0});
                st.push_back({first, 0});
            }
        }
        if (counter != prim_count) ok.store(false);
    };
    if (threads > 1) {
        std::vector<std::thread> th;
        for (int o = 0; o < 8; o++) th.emplace_back(one, o);
        for (auto& t : th) t.join();
    } else {
        for (int o = 0; o < 8; o++) one(o);
    }
    return ok.load();
}

}  // namespace lgb
