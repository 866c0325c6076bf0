// Host-side construction of the device BVH and of the reference-order rank tables (see lgb_build.hpp).
#include "lgb_build.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <immintrin.h>

#include <array>
#include <thread>

#include "lgb_parallel.hpp"

namespace lgb {
namespace {

// One primitive during the build: padded f32 box (conservative for f32 rays, see lgb_api.cu) and what it is.
struct alignas(16) Item { float lo[3]; uint32_t type; float hi[3]; uint32_t index; };   // 32 B, two 16-byte lanes
static_assert(sizeof(Item) == 32, "Item layout");

// Boxes are kept as two SSE lanes {x, y, z, 0}.
struct Box3 {
    __m128 lo, hi;
    void reset() { lo = _mm_set1_ps(std::numeric_limits<float>::infinity()); hi = _mm_set1_ps(-std::numeric_limits<float>::infinity()); }
    void grow(__m128 l, __m128 h) { lo = _mm_min_ps(lo, l); hi = _mm_max_ps(hi, h); }
    void grow(const Box3& b) { grow(b.lo, b.hi); }
    void grow_pt(__m128 p) { grow(p, p); }
    void get(float* l, float* h) const { alignas(16) float a[4], b[4]; _mm_store_ps(a, lo); _mm_store_ps(b, hi); for (int k = 0; k < 3; k++) { l[k] = a[k]; h[k] = b[k]; } }
    float extent(int k) const { alignas(16) float a[4]; _mm_store_ps(a, _mm_sub_ps(hi, lo)); return a[k]; }
    float half_area() const {
        alignas(16) float d[4]; _mm_store_ps(d, _mm_sub_ps(hi, lo));
        return d[0] * d[1] + d[0] * d[2] + d[1] * d[2];
    }
};
// lane 3 (the type / index word) is masked to +0: as float bits it would be a denormal, and denormal operands
// send mulps / addps through a microcode assist (measured: 25x slower binning)
inline __m128 lane3_mask() { return _mm_castsi128_ps(_mm_set_epi32(0, -1, -1, -1)); }
inline __m128 item_lo(const Item& it) { return _mm_and_ps(_mm_load_ps(it.lo), lane3_mask()); }
inline __m128 item_hi(const Item& it) { return _mm_and_ps(_mm_load_ps(it.hi), lane3_mask()); }
inline __m128 centroid(const Item& it) { const __m128 h = _mm_set1_ps(0.5f); return _mm_add_ps(_mm_mul_ps(h, item_lo(it)), _mm_mul_ps(h, item_hi(it))); }
inline float centroid_axis(const Item& it, int a) { return 0.5f * it.lo[a] + 0.5f * it.hi[a]; }

#ifndef LGB_SAH_BINS
#define LGB_SAH_BINS 16
#endif
constexpr int NB = LGB_SAH_BINS;
// Per axis and bin: bounds of the primitives and their count.
struct Bins {
    Box3 box[3][NB]; uint32_t cnt[3][NB];
    void reset() { for (int a = 0; a < 3; a++) for (int b = 0; b < NB; b++) { box[a][b].reset(); cnt[a][b] = 0; } }
    void merge(const Bins& o) { for (int a = 0; a < 3; a++) for (int b = 0; b < NB; b++) { box[a][b].grow(o.box[a][b]); cnt[a][b] += o.cnt[a][b]; } }
};
struct BinMap {                        // centroid -> bin along each axis of a node's centroid box
    __m128 cmin4, scale4;
    float cmin[3], scale[3]; bool any;
    void set(const Box3& cb) {
        any = false;
        float l[3], h[3]; cb.get(l, h);
        for (int a = 0; a < 3; a++) {
            const float ext = h[a] - l[a];
            cmin[a] = l[a];
            scale[a] = ext > 0.0f ? (float)NB * (1.0f - 1e-6f) / ext : 0.0f;
            any |= ext > 0.0f;
        }
        cmin4 = _mm_set_ps(0.f, cmin[2], cmin[1], cmin[0]); scale4 = _mm_set_ps(0.f, scale[2], scale[1], scale[0]);
    }
    // bin() and bin3() are the same SSE operations in the same operand order (clamp in float, then truncate):
    // the bin counts of one pass and the partition of the next must agree item by item.
    int bin(float c, int a) const {
        const __m128 x = _mm_mul_ss(_mm_sub_ss(_mm_set_ss(c), _mm_set_ss(cmin[a])), _mm_set_ss(scale[a]));
        return _mm_cvttss_si32(_mm_min_ss(_mm_max_ss(x, _mm_setzero_ps()), _mm_set_ss((float)(NB - 1))));
    }
    void bin3(__m128 c, int k[3]) const {
        const __m128 x = _mm_mul_ps(_mm_sub_ps(c, cmin4), scale4);
        const __m128i q = _mm_cvttps_epi32(_mm_min_ps(_mm_max_ps(x, _mm_setzero_ps()), _mm_set1_ps((float)(NB - 1))));
        alignas(16) int v[4]; _mm_store_si128((__m128i*)v, q);
        k[0] = v[0]; k[1] = v[1]; k[2] = v[2];
    }
};
inline void bin_range(const Item* it, size_t n, const BinMap& bm, Bins& bins) {
    for (size_t i = 0; i < n; i++) {
        const __m128 lo = item_lo(it[i]), hi = item_hi(it[i]);
        int k[3]; bm.bin3(centroid(it[i]), k);
        bins.box[0][k[0]].grow(lo, hi); bins.cnt[0][k[0]]++;
        bins.box[1][k[1]].grow(lo, hi); bins.cnt[1][k[1]]++;
        bins.box[2][k[2]].grow(lo, hi); bins.cnt[2][k[2]]++;
    }
}
struct Split { int axis = -1, bin = -1; Box3 lbox, rbox; uint32_t nl = 0; };
// 16-bin SAH over the three axes (cost = area x count per side); axis = -1 if no plane separates the centroids.
inline Split choose_split(const Bins& bins, const BinMap& bm) {
    Split sp;
    float best = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) {
        if (bm.scale[a] == 0.0f) continue;
        float la[NB]; uint32_t lc[NB];
        Box3 acc; acc.reset(); uint32_t c = 0;
        for (int k = 0; k < NB; k++) { acc.grow(bins.box[a][k]); c += bins.cnt[a][k]; la[k] = c ? acc.half_area() : 0.0f; lc[k] = c; }
        acc.reset(); c = 0;
        for (int k = NB - 1; k >= 1; k--) {
            acc.grow(bins.box[a][k]); c += bins.cnt[a][k];
            if (!c || !lc[k - 1]) continue;
            const float cost = la[k - 1] * (float)lc[k - 1] + acc.half_area() * (float)c;
            if (cost < best) { best = cost; sp.axis = a; sp.bin = k - 1; }
        }
    }
    if (sp.axis >= 0) {
        sp.lbox.reset(); sp.rbox.reset();
        for (int k = 0; k < NB; k++) {
            if (k <= sp.bin) { sp.lbox.grow(bins.box[sp.axis][k]); sp.nl += bins.cnt[sp.axis][k]; }
            else sp.rbox.grow(bins.box[sp.axis][k]);
        }
    }
    return sp;
}
inline void bounds_of(const Item* it, size_t n, Box3& box, Box3& cb) {
    box.reset(); cb.reset();
    for (size_t i = 0; i < n; i++) { box.grow(item_lo(it[i]), item_hi(it[i])); cb.grow_pt(centroid(it[i])); }
}
// In-place partition by `bin(centroid[a]) <= split_bin`; also returns the centroid bounds of both sides.
inline size_t partition_items(Item* it, size_t n, const BinMap& bm, int a, int split_bin, Box3& lcb, Box3& rcb) {
    lcb.reset(); rcb.reset();
    size_t i = 0, j = n;
    auto left = [&](const Item& x) { return bm.bin(centroid_axis(x, a), a) <= split_bin; };
    for (;;) {
        while (i < j && left(it[i])) { lcb.grow_pt(centroid(it[i])); i++; }
        while (i < j && !left(it[j - 1])) { rcb.grow_pt(centroid(it[j - 1])); j--; }
        if (i >= j) break;
        std::swap(it[i], it[j - 1]);
        lcb.grow_pt(centroid(it[i])); rcb.grow_pt(centroid(it[j - 1]));
        i++; j--;
    }
    return i;
}
inline void write_node_boxes(HostNode& nd, const Box3& l, const Box3& r) { l.get(nd.v, nd.v + 3); r.get(nd.v + 6, nd.v + 9); }

// A sub-tree built by one thread into its own node / leaf-order arrays; indices are rebased when the pieces
// are assembled, so the result does not depend on thread timing.
struct SubTree {
    Item* items = nullptr; size_t n = 0;
    Box3 box, cb; uint32_t depth = 0;
    uint32_t parent = 0; int which = 0;                // slot of the top tree that points at this sub-tree
    std::vector<HostNode> nodes;                       // local indices; node 0 = sub-tree root (if interior)
    std::vector<uint32_t> order[4];
    uint32_t root_word = 0, max_depth = 0;
    uint32_t node_base = 0, type_base[4] = {0, 0, 0, 0};

    uint32_t make_leaf(Item* it, size_t cnt, uint32_t depth_) {
        const uint32_t type = it[0].type, first = (uint32_t)order[type].size();
        for (size_t i = 0; i < cnt; i++) order[type].push_back(it[i].index);
        max_depth = std::max(max_depth, depth_);
        return kLeafBit | (type << 29) | ((uint32_t)(cnt - 1) << 24) | first;
    }
    // Child word of the sub-tree over it[0, cnt) whose bounds / centroid bounds are already known.
    uint32_t build(Item* it, size_t cnt, const Box3& cbox, uint32_t depth_) {
        bool mixed = false;
        if (cnt <= (size_t)kMaxLeaf) {
            for (size_t i = 1; i < cnt; i++) mixed |= it[i].type != it[0].type;
            if (!mixed) return make_leaf(it, cnt, depth_);
        }
        size_t mid = 0;
        Box3 lbox, rbox, lcb, rcb;
        bool have_child_bounds = false;
        if (cnt <= (size_t)kMaxLeaf) {             // small but mixed types: split at the first type boundary
            std::stable_sort(it, it + cnt, [](const Item& x, const Item& y) { return x.type < y.type; });
            mid = 1;
            while (it[mid].type == it[0].type) mid++;
        } else {
            BinMap bm; bm.set(cbox);
            Split sp;
            if (bm.any && depth_ < 40) {
                Bins bins; bins.reset();
                bin_range(it, cnt, bm, bins);
                sp = choose_split(bins, bm);
            }
            if (sp.axis >= 0) {
                mid = partition_items(it, cnt, bm, sp.axis, sp.bin, lcb, rcb);
                lbox = sp.lbox; rbox = sp.rbox; have_child_bounds = true;
            }
            if (mid == 0 || mid == cnt) {          // coincident centroids or depth guard: median split along the widest axis
                int a = 0;
                for (int k = 1; k < 3; k++) if (cbox.extent(k) > cbox.extent(a)) a = k;
                mid = cnt / 2;
                std::nth_element(it, it + mid, it + cnt, [a](const Item& x, const Item& y) { return centroid_axis(x, a) < centroid_axis(y, a); });
                have_child_bounds = false;
            }
        }
        if (!have_child_bounds) { bounds_of(it, mid, lbox, lcb); bounds_of(it + mid, cnt - mid, rbox, rcb); }
        const uint32_t me = (uint32_t)nodes.size();
        nodes.push_back(HostNode{});
        const uint32_t lw = build(it, mid, lcb, depth_ + 1);
        const uint32_t rw = build(it + mid, cnt - mid, rcb, depth_ + 1);
        HostNode& nd = nodes[me];
        write_node_boxes(nd, lbox, rbox);
        nd.c0 = lw; nd.c1 = rw; nd.pad0 = nd.pad1 = 0;
        return me;
    }
    uint32_t rebase(uint32_t w) const {
        if (w & kLeafBit) { const uint32_t type = (w >> 29) & 3u; return (w & ~kLeafFirstMask) | ((w & kLeafFirstMask) + type_base[type]); }
        return w + node_base;
    }
};

// Segment of the item array that is still being split by all threads together (top of the tree).
struct Seg { size_t b, e; Box3 cb; uint32_t node, depth; };
constexpr size_t kChunk = 8192;            // items per parallel work unit in the top phase
constexpr size_t kSubTree = 12288;         // segments at most this long are finished by one thread each

}  // namespace

// Binned-SAH build in two phases.  Top: while a segment is longer than kSubTree, all threads bin it, count and
// scatter it (level-synchronous over every such segment, ping-pong between two item buffers).  Bottom: the
// remaining segments are independent sub-trees, built one per task and assembled deterministically.
int build_sah(const PrimBox* prims, size_t n, float pad, int threads, BuiltBVH& out) {
    (void)threads;
    auto t0 = std::chrono::steady_clock::now();
    Pool& pool = Pool::get();
    if (n == 0) return -1;
    raw_vector<Item> buf0(n), buf1;
    std::atomic<bool> bad{false};
    size_t per_type[4] = {0, 0, 0, 0};
    {
        const size_t nc = pool.chunks_of(n, kChunk);
        std::vector<std::array<size_t, 4>> cnt(nc, {0, 0, 0, 0});
        pool.for_range(n, kChunk, [&](size_t b, size_t e, size_t c) {
            for (size_t i = b; i < e; i++) {
                Item& it = buf0[i];
                for (int k = 0; k < 3; k++) {
                    it.lo[k] = std::nextafterf(prims[i].lo[k] - pad, -INFINITY);
                    it.hi[k] = std::nextafterf(prims[i].hi[k] + pad, INFINITY);
                }
                it.type = prims[i].type; it.index = prims[i].index;
                if (it.type > 3) { bad.store(true); continue; }
                cnt[c][it.type]++;
            }
        });
        if (bad.load()) return -1;
        for (auto& c : cnt) for (int t = 0; t < 4; t++) per_type[t] += c[t];
    }
    for (int t = 0; t < 4; t++) { if (per_type[t] > kLeafFirstMask) return -2; out.order[t].resize(per_type[t]); }

    std::vector<HostNode> top;                 // nodes created by the top phase, in creation order (node 0 = root)
    std::vector<SubTree> subs;
    Item* cur = buf0.data();
    Item* other = nullptr;

    bool homogeneous = true;
    for (size_t i = 1; i < n && i <= (size_t)kMaxLeaf && homogeneous; i++) homogeneous = buf0[i].type == buf0[0].type;
    if (n <= (size_t)kMaxLeaf && homogeneous) {      // tiny scene: the root's two children are the same leaf (testing it twice changes nothing)
        SubTree st; Box3 box, cb; bounds_of(cur, n, box, cb);
        const uint32_t w = st.make_leaf(cur, n, 1);
        HostNode nd{};
        write_node_boxes(nd, box, box);
        nd.c0 = w; nd.c1 = w;
        out.nodes.assign(1, nd);
        for (int t = 0; t < 4; t++) out.order[t] = st.order[t];
        out.max_depth = 1;
        out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return 0;
    }

    // ---- top phase
    std::vector<Seg> active;
    {
        Seg root; root.b = 0; root.e = n; root.node = 0; root.depth = 0;
        const size_t nc = pool.chunks_of(n, kChunk);
        std::vector<Box3> cbs(nc); for (auto& b : cbs) b.reset();
        pool.for_range(n, kChunk, [&](size_t b, size_t e, size_t c) { Box3 bx; bounds_of(cur + b, e - b, bx, cbs[c]); });
        root.cb.reset(); for (auto& b : cbs) root.cb.grow(b);
        top.push_back(HostNode{});
        if (n <= kSubTree) {
            subs.emplace_back(); SubTree& st = subs.back();
            st.items = cur; st.n = n; st.cb = root.cb; st.depth = 0; st.parent = 0xFFFFFFFFu;    // the sub-tree IS the tree
            top.clear();
        } else active.push_back(root);
    }
    struct ChunkRef { uint32_t seg; size_t b, e; };
    while (!active.empty()) {
        if (!other) { buf1.resize(n); other = buf1.data(); }
        std::vector<ChunkRef> chunks;
        std::vector<size_t> first_chunk(active.size() + 1);
        for (size_t s = 0; s < active.size(); s++) {
            first_chunk[s] = chunks.size();
            for (size_t b = active[s].b; b < active[s].e; b += kChunk) chunks.push_back({(uint32_t)s, b, std::min(active[s].e, b + kChunk)});
        }
        first_chunk[active.size()] = chunks.size();
        std::vector<BinMap> bm(active.size());
        for (size_t s = 0; s < active.size(); s++) bm[s].set(active[s].cb);
        std::vector<Bins> cbins(chunks.size());
        pool.run(chunks.size(), [&](size_t c) { cbins[c].reset(); bin_range(cur + chunks[c].b, chunks[c].e - chunks[c].b, bm[chunks[c].seg], cbins[c]); });
        std::vector<Split> split(active.size());
        for (size_t s = 0; s < active.size(); s++) {
            if (!bm[s].any || active[s].depth >= 40) continue;
            Bins total = cbins[first_chunk[s]];
            for (size_t c = first_chunk[s] + 1; c < first_chunk[s + 1]; c++) total.merge(cbins[c]);
            split[s] = choose_split(total, bm[s]);
        }
        // left counts per chunk -> scatter offsets (stable: chunk order is kept on both sides)
        std::vector<size_t> nleft(chunks.size(), 0);
        std::vector<Box3> clcb(chunks.size()), crcb(chunks.size());      // centroid bounds of both sides, per chunk
        pool.run(chunks.size(), [&](size_t c) {
            const Split& sp = split[chunks[c].seg];
            clcb[c].reset(); crcb[c].reset();
            if (sp.axis < 0) return;
            const BinMap& m = bm[chunks[c].seg]; const int a = sp.axis; size_t k = 0;
            for (size_t i = chunks[c].b; i < chunks[c].e; i++) {
                const bool l = m.bin(centroid_axis(cur[i], a), a) <= sp.bin;
                (l ? clcb[c] : crcb[c]).grow_pt(centroid(cur[i]));
                k += l;
            }
            nleft[c] = k;
        });
        std::vector<size_t> loff(chunks.size()), roff(chunks.size());
        for (size_t s = 0; s < active.size(); s++) {
            if (split[s].axis < 0) continue;
            size_t l = active[s].b, total_left = 0;
            for (size_t c = first_chunk[s]; c < first_chunk[s + 1]; c++) total_left += nleft[c];
            size_t r = active[s].b + total_left;
            for (size_t c = first_chunk[s]; c < first_chunk[s + 1]; c++) { loff[c] = l; roff[c] = r; l += nleft[c]; r += (chunks[c].e - chunks[c].b) - nleft[c]; }
        }
        pool.run(chunks.size(), [&](size_t c) {
            const Split& sp = split[chunks[c].seg];
            if (sp.axis < 0) { std::copy(cur + chunks[c].b, cur + chunks[c].e, other + chunks[c].b); return; }
            const BinMap& m = bm[chunks[c].seg]; const int a = sp.axis; size_t l = loff[c], r = roff[c];
            for (size_t i = chunks[c].b; i < chunks[c].e; i++) {
                if (m.bin(centroid_axis(cur[i], a), a) <= sp.bin) other[l++] = cur[i]; else other[r++] = cur[i];
            }
        });
        std::swap(cur, other);
        std::vector<Seg> next;
        for (size_t s = 0; s < active.size(); s++) {
            const Seg& sg = active[s];
            size_t mid; Box3 lbox, rbox, lcb, rcb;
            if (split[s].axis >= 0) {
                mid = sg.b + split[s].nl; lbox = split[s].lbox; rbox = split[s].rbox;
                lcb.reset(); rcb.reset();
                for (size_t c = first_chunk[s]; c < first_chunk[s + 1]; c++) { lcb.grow(clcb[c]); rcb.grow(crcb[c]); }
            } else {                                   // no separating plane: median split along the widest centroid axis
                int a = 0;
                for (int k = 1; k < 3; k++) if (sg.cb.extent(k) > sg.cb.extent(a)) a = k;
                mid = sg.b + (sg.e - sg.b) / 2;
                std::nth_element(cur + sg.b, cur + mid, cur + sg.e, [a](const Item& x, const Item& y) { return centroid_axis(x, a) < centroid_axis(y, a); });
                bounds_of(cur + sg.b, mid - sg.b, lbox, lcb); bounds_of(cur + mid, sg.e - mid, rbox, rcb);
            }
            write_node_boxes(top[sg.node], lbox, rbox);
            for (int which = 0; which < 2; which++) {
                const size_t b = which ? mid : sg.b, e = which ? sg.e : mid;
                const Box3& cb = which ? rcb : lcb;
                if (e - b > kSubTree) {
                    Seg ch; ch.b = b; ch.e = e; ch.cb = cb; ch.depth = sg.depth + 1; ch.node = (uint32_t)top.size();
                    (which ? top[sg.node].c1 : top[sg.node].c0) = ch.node;
                    top.push_back(HostNode{});
                    next.push_back(ch);
                } else {
                    subs.emplace_back(); SubTree& st = subs.back();
                    st.items = cur + b; st.n = e - b; st.cb = cb; st.depth = sg.depth + 1; st.parent = sg.node; st.which = which;
                }
            }
        }
        active.swap(next);
    }

    const double t_top = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // ---- bottom phase: one task per sub-tree, largest first
    std::vector<uint32_t> by_size(subs.size());
    for (size_t i = 0; i < subs.size(); i++) by_size[i] = (uint32_t)i;
    std::sort(by_size.begin(), by_size.end(), [&](uint32_t x, uint32_t y) { return subs[x].n != subs[y].n ? subs[x].n > subs[y].n : x < y; });
    pool.run(subs.size(), [&](size_t k) {
        SubTree& st = subs[by_size[k]];
        st.nodes.reserve(st.n);
        st.root_word = st.build(st.items, st.n, st.cb, st.depth);
    });
    const double t_bottom = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // ---- assemble: top nodes first, then every sub-tree's nodes contiguously (depth-first inside each)
    uint32_t node_total = (uint32_t)top.size(), tb[4] = {0, 0, 0, 0}, max_depth = 0;
    for (SubTree& st : subs) {
        st.node_base = node_total; node_total += (uint32_t)st.nodes.size();
        for (int t = 0; t < 4; t++) { st.type_base[t] = tb[t]; tb[t] += (uint32_t)st.order[t].size(); }
        max_depth = std::max(max_depth, st.max_depth);
    }
    out.nodes.resize(node_total);
    std::copy(top.begin(), top.end(), out.nodes.begin());
    pool.run(subs.size(), [&](size_t k) {
        const SubTree& st = subs[k];
        for (size_t i = 0; i < st.nodes.size(); i++) {
            HostNode nd = st.nodes[i];
            nd.c0 = st.rebase(nd.c0); nd.c1 = st.rebase(nd.c1);
            out.nodes[st.node_base + i] = nd;
        }
        for (int t = 0; t < 4; t++) std::copy(st.order[t].begin(), st.order[t].end(), out.order[t].begin() + st.type_base[t]);
    });
    for (const SubTree& st : subs) {
        if (st.parent == 0xFFFFFFFFu) { if (st.root_word != 0 || st.node_base != 0) return -3; continue; }   // the root must be node 0
        (st.which ? out.nodes[st.parent].c1 : out.nodes[st.parent].c0) = st.rebase(st.root_word);
    }
    out.max_depth = max_depth;
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (std::getenv("LGB_TIMING")) std::fprintf(stderr, "[build_sah] %zu prims, %d threads: top %.1f ms (%zu nodes), sub-trees %.1f ms (%zu), assemble %.1f ms\n",
                                                 n, pool.threads(), t_top, top.size(), t_bottom - t_top, subs.size(), out.build_ms - t_bottom);
    return 0;
}

// ------------------------------------------------------------------ spaces
namespace {
inline float f32_down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
inline float f32_up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }
// Transform3::transform_bounds (transform.rs:219-240) of the box [lo, hi] by the column-major matrix m.
inline void xform_box(const double m[16], const double lo[3], const double hi[3], double olo[3], double ohi[3]) {
    for (int r = 0; r < 3; r++) {
        double l = m[12 + r], h = m[12 + r];
        for (int c = 0; c < 3; c++) {
            const double a = m[4 * c + r] * lo[c], b = m[4 * c + r] * hi[c];
            l += a < b ? a : b; h += a < b ? b : a;
        }
        olo[r] = l; ohi[r] = h;
    }
}
inline bool affine(const double m[16]) { return m[3] == 0.0 && m[7] == 0.0 && m[11] == 0.0 && m[15] == 1.0; }
}  // namespace

int build_scene(const lgb_scene_desc* d, const double world_lo[3], const double world_hi[3], BuiltScene& out, std::string& err) {
    auto t0 = std::chrono::steady_clock::now();
    Pool& pool = Pool::get();
    const size_t ns = d->n_spheres, nc = d->n_cuboids, nt = d->n_triangles, np = ns + nc + nt;
    // ---- discover the spaces (breadth-first over the nested levels: a child space has a larger index than its parent)
    out.spaces.clear(); out.inst_space.assign(d->n_instances, kNoSpace);
    {
        HostSpace s0; s0.ref_root = 0; s0.identity = d->root.identity; s0.swap_backface = d->root.swap_backface;
        std::memcpy(s0.m, d->root.m, sizeof s0.m); std::memcpy(s0.minv, d->root.minv, sizeof s0.minv);
        if (s0.identity) for (int k = 0; k < 16; k++) s0.m[k] = s0.minv[k] = (k % 5 == 0) ? 1.0 : 0.0;
        out.spaces.push_back(s0);
    }
    bool multi = false;
    for (uint64_t i = 0; i < d->n_instances; i++) multi |= !d->instances[i].identity || d->instances[i].swap_backface;
    std::vector<std::vector<uint32_t>> children(1);
    if (multi) {
        for (int t = 0; t < 3; t++) out.prim_space[t].resize(t == 0 ? ns : t == 1 ? nc : nt);
        std::vector<uint8_t> seen_inst(d->n_instances, 0);
        for (size_t sidx = 0; sidx < out.spaces.size(); sidx++) {
            std::vector<uint32_t> st{out.spaces[sidx].ref_root};
            while (!st.empty()) {
                const uint32_t ni = st.back(); st.pop_back();
                const lgb_node& n = d->nodes[ni];
                if (n.b & LGB_LEAF_FLAG) {
                    const uint32_t cnt = n.b & ~LGB_LEAF_FLAG;
                    for (uint32_t j = 0; j < cnt; j++) {
                        const uint32_t ref = d->prim_refs[n.a + j], type = ref >> 30, idx = ref & 0x3FFFFFFFu;
                        if (type != LGB_PRIM_INSTANCE) { out.prim_space[type][idx] = (uint32_t)sidx; continue; }
                        if (seen_inst[idx]++) { err = "instance referenced by more than one leaf"; return -1; }
                        const lgb_instance& in = d->instances[idx];
                        if (in.identity && !in.swap_backface) { st.push_back(in.root_node); continue; }
                        if (!affine(in.m) || !affine(in.minv)) { err = "instance: transform is not affine (row 3 must be 0 0 0 1)"; return -4; }
                        HostSpace c; c.parent = (uint32_t)sidx; c.depth = out.spaces[sidx].depth + 1; c.ref_root = in.root_node;
                        c.identity = in.identity; c.swap_backface = in.swap_backface;
                        std::memcpy(c.m, in.m, sizeof c.m); std::memcpy(c.minv, in.minv, sizeof c.minv);
                        out.inst_space[idx] = (uint32_t)out.spaces.size();
                        children[sidx].push_back((uint32_t)out.spaces.size());
                        out.spaces.push_back(c); children.emplace_back();
                    }
                } else { st.push_back(n.a); st.push_back(ni + 1); }
            }
        }
    }
    if (!out.spaces[0].identity && (!affine(out.spaces[0].m) || !affine(out.spaces[0].minv))) { err = "root transform is not affine (row 3 must be 0 0 0 1)"; return -4; }
    const size_t nsp = out.spaces.size();
    // ---- coordinate bound of every space: its own geometry and every ray origin taken into it
    {
        std::vector<std::array<double, 6>> obox(nsp);       // box that contains the ray origins, in the space's coordinates
        for (size_t sidx = 0; sidx < nsp; sidx++) {
            HostSpace& sp = out.spaces[sidx];
            const double* plo = sp.parent == kNoSpace ? world_lo : obox[sp.parent].data();
            const double* phi = sp.parent == kNoSpace ? world_hi : obox[sp.parent].data() + 3;
            if (sp.identity) { for (int k = 0; k < 3; k++) { obox[sidx][k] = plo[k]; obox[sidx][3 + k] = phi[k]; } }
            else xform_box(sp.minv, plo, phi, obox[sidx].data(), obox[sidx].data() + 3);
            double M = 0.0;
            auto upd = [&](double v) { const double a = std::fabs(v); if (a > M && std::isfinite(a)) M = a; };
            for (int k = 0; k < 6; k++) upd(obox[sidx][k]);
            for (int k = 0; k < 3; k++) { upd(d->nodes[sp.ref_root].lo[k]); upd(d->nodes[sp.ref_root].hi[k]); }
            sp.max_abs = M; sp.err_abs = (float)(M * std::ldexp(1.0, -20));
        }
    }
    // ---- primitive boxes (each in the coordinates of its own level), grouped by space
    raw_vector<PrimBox> prims(np);
    pool.for_range(ns, 1 << 14, [&](size_t b0, size_t e0, size_t) {
        for (size_t i = b0; i < e0; i++) {
            const lgb_sphere& sp = d->spheres[i]; PrimBox& b = prims[i]; b.type = LGB_PRIM_SPHERE; b.index = (uint32_t)i;
            for (int k = 0; k < 3; k++) { double lo = sp.center[k] - sp.radius, hi = sp.center[k] + sp.radius; b.lo[k] = f32_down(std::min(lo, hi)); b.hi[k] = f32_up(std::max(lo, hi)); }
        }
    });
    for (size_t i = 0; i < nc; i++) {
        const lgb_cuboid& c = d->cuboids[i]; PrimBox& b = prims[ns + i]; b.type = LGB_PRIM_CUBOID; b.index = (uint32_t)i;
        for (int k = 0; k < 3; k++) { b.lo[k] = f32_down(std::min(c.min[k], c.max[k])); b.hi[k] = f32_up(std::max(c.min[k], c.max[k])); }
    }
    pool.for_range(nt, 1 << 14, [&](size_t b0, size_t e0, size_t) {
        for (size_t i = b0; i < e0; i++) {
            const lgb_triangle& t = d->triangles[i]; PrimBox& b = prims[ns + nc + i]; b.type = LGB_PRIM_TRIANGLE; b.index = (uint32_t)i;
            for (int k = 0; k < 3; k++) { b.lo[k] = std::min(t.p0[k], std::min(t.p1[k], t.p2[k])); b.hi[k] = std::max(t.p0[k], std::max(t.p1[k], t.p2[k])); }
        }
    });
    std::vector<size_t> sp_begin(nsp + 1, 0);
    raw_vector<PrimBox> grouped;
    const PrimBox* by_space = prims.data();
    if (nsp > 1) {                               // stable counting sort of the primitives by space
        auto space_of = [&](const PrimBox& b) { return out.prim_space[b.type][b.index]; };
        for (size_t i = 0; i < np; i++) sp_begin[space_of(prims[i]) + 1]++;
        for (size_t k = 0; k < nsp; k++) sp_begin[k + 1] += sp_begin[k];
        std::vector<size_t> cursor(sp_begin.begin(), sp_begin.end() - 1);
        grouped.resize(np);
        for (size_t i = 0; i < np; i++) grouped[cursor[space_of(prims[i])]++] = prims[i];
        by_space = grouped.data();
    } else sp_begin[1] = np;
    // ---- one SAH BVH per space, children before parents (a parent needs the boxes of its child spaces)
    std::vector<BuiltBVH> bvh(nsp);
    std::vector<std::array<float, 6>> root_box(nsp);
    for (size_t k = nsp; k-- > 0;) {
        HostSpace& sp = out.spaces[k];
        raw_vector<PrimBox> items;
        const PrimBox* it = by_space + sp_begin[k];
        size_t n = sp_begin[k + 1] - sp_begin[k];
        if (!children[k].empty()) {
            items.resize(n + children[k].size());
            std::copy(it, it + n, items.begin());
            for (uint32_t c : children[k]) {
                double lo[3], hi[3], tlo[3], thi[3];
                for (int a = 0; a < 3; a++) { lo[a] = root_box[c][a]; hi[a] = root_box[c][3 + a]; }
                xform_box(out.spaces[c].m, lo, hi, tlo, thi);
                PrimBox& b = items[n++]; b.type = LGB_PRIM_INSTANCE; b.index = c;
                for (int a = 0; a < 3; a++) { b.lo[a] = f32_down(std::min(tlo[a], thi[a])); b.hi[a] = f32_up(std::max(tlo[a], thi[a])); }
            }
            it = items.data();
        }
        if (n == 0) { err = "a nested aggregate has no primitives"; return -1; }
        const int rc = build_sah(it, n, sp.err_abs, pool.threads(), bvh[k]);
        if (rc) { err = rc == -2 ? "more than 16.7M primitives of one type" : "BVH build failed"; return rc == -2 ? -2 : -1; }
        const HostNode& r = bvh[k].nodes[0];
        for (int a = 0; a < 3; a++) { root_box[k][a] = std::min(r.v[a], r.v[6 + a]); root_box[k][3 + a] = std::max(r.v[3 + a], r.v[9 + a]); }
        // traversal stack: the space's own tree, plus (resume entry + exit marker + the deepest child) under an instance leaf
        uint32_t need = bvh[k].max_depth + 1, worst = 0;
        for (uint32_t c : children[k]) worst = std::max(worst, out.spaces[c].stack_need);
        sp.stack_need = need + (children[k].empty() ? 0 : 2 + worst);
    }
    // ---- concatenate: nodes and leaf-ordered primitive lists of all spaces, child words made absolute
    uint32_t node_base = 0, tb[4] = {0, 0, 0, 0};
    size_t node_total = 0, tt[4] = {0, 0, 0, 0};
    for (size_t k = 0; k < nsp; k++) { node_total += bvh[k].nodes.size(); for (int t = 0; t < 4; t++) tt[t] += bvh[k].order[t].size(); }
    for (int t = 0; t < 4; t++) { if (tt[t] > kLeafFirstMask) { err = "more than 16.7M primitives of one type"; return -2; } out.order[t].resize(tt[t]); }
    if (nsp == 1) out.nodes.swap(bvh[0].nodes); else out.nodes.resize(node_total);
    for (size_t k = 0; k < nsp; k++) {
        out.spaces[k].root_node = node_base;
        if (nsp > 1) {
            const uint32_t nb = node_base; const uint32_t* tbp = tb;
            auto rebase = [nb, tbp](uint32_t w) {
                if (w & kLeafBit) { const uint32_t type = (w >> 29) & 3u; return (w & ~kLeafFirstMask) | ((w & kLeafFirstMask) + tbp[type]); }
                return w + nb;
            };
            for (size_t i = 0; i < bvh[k].nodes.size(); i++) { HostNode nd = bvh[k].nodes[i]; nd.c0 = rebase(nd.c0); nd.c1 = rebase(nd.c1); out.nodes[node_base + i] = nd; }
            node_base += (uint32_t)bvh[k].nodes.size();
        }
        for (int t = 0; t < 4; t++) { std::copy(bvh[k].order[t].begin(), bvh[k].order[t].end(), out.order[t].begin() + tb[t]); tb[t] += (uint32_t)bvh[k].order[t].size(); }
    }
    out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

bool build_rank_tables(const lgb_scene_desc* d, uint32_t prim_count, const BuiltScene& bs, uint32_t* rank) {
    Pool& pool = Pool::get();
    const size_t nsp = bs.spaces.size(), items = (size_t)prim_count + nsp;
    pool.for_range((size_t)8 * items, 1 << 16, [&](size_t b, size_t e, size_t) { std::fill(rank + b, rank + e, 0xFFFFFFFFu); });
    std::atomic<bool> ok{true};
    std::vector<uint64_t> counted(8 * nsp, 0);
    auto one = [&](size_t task) {
        const size_t oct = task & 7, sidx = task >> 3;
        uint32_t* r = rank + oct * items;
        uint32_t counter = 0;
        // explicit DFS: entries are node indices; a leaf expands its primitives in order, recursing into nested
        // identity levels immediately (bvh.rs:483-488 calls the nested intersect before the next primitive); a
        // nested level that opens its own space is ONE item here and is ranked again, inside, by its own octant.
        struct Frame { uint32_t node; uint32_t next_ref; };
        std::vector<Frame> st;
        st.push_back({bs.spaces[sidx].ref_root, 0});
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
                        if (bs.inst_space[idx] == kNoSpace) {
                            st.push_back({fr.node, i + 1});                       // resume this leaf afterwards
                            st.push_back({d->instances[idx].root_node, 0});
                            break;
                        }
                        id = prim_count + bs.inst_space[idx];
                    } else id = type == LGB_PRIM_SPHERE ? d->sphere_id[idx] : type == LGB_PRIM_CUBOID ? d->cuboid_id[idx] : d->triangle_id[idx];
                    if (id >= items || (type != LGB_PRIM_INSTANCE && id >= prim_count) || r[id] != 0xFFFFFFFFu) { ok.store(false); return; }
                    r[id] = counter++;
                }
            } else {
                const bool neg = (oct >> n.b) & 1;                             // bvh.rs:496-503
                const uint32_t first = neg ? n.a : fr.node + 1, second = neg ? fr.node + 1 : n.a;
                st.push_back({second, 0});
                st.push_back({first, 0});
            }
        }
        counted[task] = counter;
    };
    pool.run(8 * nsp, one);
    if (!ok.load()) return false;
    for (size_t oct = 0; oct < 8; oct++) {
        uint64_t total = 0;
        for (size_t sidx = 0; sidx < nsp; sidx++) total += counted[sidx * 8 + oct];
        if (total != (uint64_t)prim_count + nsp - 1) return false;
    }
    return true;
}

}  // namespace lgb
