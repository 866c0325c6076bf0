// Device-side construction of the device BVH: the same 16-bin, 3-axis SAH tree as csrc/lgb_build.cpp builds on
// the host (leaves <= kMaxLeaf, homogeneous type), level-synchronous on the GPU.  The host build costs ~25 ms of a
// 108 ms capture for 600 k primitives; here a level is four small kernels over the item array:
//   k_prep     per node of the level: centroid bounds -> bin map
//   k_bin      per item: box and count into the (axis, bin) slots of its node (warp-aggregated atomics on ordered keys)
//   k_split    per node: SAH sweep over the bins -> split, child ranges, node record; small nodes: leaf or type split
//   k_scatter  per item: move to its side (warp-aggregated cursors), grow the child's centroid bounds; items of
//              finished leaves go to their final position
// Items keep nested ranges [begin, end), so the final item array is in tree order and a leaf is a contiguous run.
// A node of at most kSub items leaves the level loop: its whole sub-tree is finished by ONE warp in shared memory (k_subtrees: same
// bins, same sweep, same float arithmetic), so the loop runs ~log2(n / kSub) levels instead of the tree's depth and the deep levels --
// a quarter of a million nodes, each with 1.3 KB of bins in global memory -- cost one launch.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#include "lgb_build.hpp"
#include "lgb_gpubuild.cuh"

namespace lgb {
namespace {

constexpr int NB = 16;
constexpr uint32_t kInvalid = 0xFFFFFFFFu;
constexpr uint32_t kKeyMax = 0xFFFFFFFFu;

// order-preserving float <-> uint map: atomicMin / atomicMax on the keys are min / max on the floats
__device__ __forceinline__ uint32_t fkey(float f) { const uint32_t b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ float kfloat(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }

struct Work {                      // one node of the current level
    uint32_t begin, end;           // its items
    uint32_t parent, which;        // where its child word goes (parent == kInvalid: it is the root, node 0)
    uint32_t depth;
    uint32_t cb_lo[3], cb_hi[3];   // centroid bounds (keys), grown by the scatter of the level above
    float cmin[3], scale[3];       // bin map (k_prep)
    // decision (k_split)
    uint32_t mode;                 // 0 SAH bin split, 1 split by type, 2 split by position, 3 leaf
    uint32_t axis, bin, tmin, nl;
    uint32_t child[2];             // index in the next level's work list, or kInvalid
    uint32_t cur_l, cur_r;         // scatter cursors
    uint32_t bins;                 // slot in the bin pool, or kInvalid for a node of <= kMaxLeaf items
};
// per node with more than kMaxLeaf items: [axis][bin] box lo (3 keys), box hi (3 keys), count
constexpr int kBinWords = 3 * NB * 7;

struct Ctl { uint32_t next_count, node_count, max_depth, bin_slots, count, levels, sub_count; };      // count: nodes of the level being split (k_advance)
constexpr uint32_t kSub = 256;         // a node of at most this many items is finished by one warp (k_subtrees)
struct SubRoot { uint32_t begin, end, parent, which, depth; };
// Start of a level: the nodes the previous level produced become the current ones.  The host does not read the count back every
// level (28 round trips for 600 k primitives): the kernels of a level take it from here, are launched for the largest level there can
// be, and a level past the last one does nothing.
__global__ void k_advance(Ctl* ctl) {
    ctl->count = ctl->next_count; ctl->next_count = 0; ctl->bin_slots = 0;
    if (ctl->count) ctl->levels++;
}

__device__ __forceinline__ int bin_of(float c, float cmin, float scale) {
    const float x = fminf(fmaxf((c - cmin) * scale, 0.0f), (float)(NB - 1));
    return (int)x;
}
__device__ __forceinline__ float centroid(const GItem& it, int a) { return 0.5f * it.lo[a] + 0.5f * it.hi[a]; }

__global__ void k_root(const GItem* items, uint32_t n, Work* work, Ctl* ctl) {
    // centroid bounds of everything
    uint32_t lo[3] = {kKeyMax, kKeyMax, kKeyMax}, hi[3] = {0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int a = 0; a < 3; a++) { const uint32_t k = fkey(centroid(items[i], a)); lo[a] = min(lo[a], k); hi[a] = max(hi[a], k); }
    for (int a = 0; a < 3; a++) {
        lo[a] = __reduce_min_sync(0xFFFFFFFFu, lo[a]); hi[a] = __reduce_max_sync(0xFFFFFFFFu, hi[a]);
        if ((threadIdx.x & 31) == 0) { atomicMin(&work[0].cb_lo[a], lo[a]); atomicMax(&work[0].cb_hi[a], hi[a]); }
    }
    (void)ctl;
}

__global__ void k_prep(Work* work, uint32_t* bins, Ctl* ctl) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ctl->count) return;
    Work& W = work[w];
    for (int a = 0; a < 3; a++) {
        const float lo = kfloat(W.cb_lo[a]), hi = kfloat(W.cb_hi[a]);
        const float ext = hi - lo;
        W.cmin[a] = lo;
        W.scale[a] = ext > 0.0f ? (float)NB * (1.0f - 1e-6f) / ext : 0.0f;
    }
    W.cur_l = W.cur_r = 0;
    W.bins = kInvalid;
    const uint32_t n = W.end - W.begin;
    if (n > kSub || (W.parent == kInvalid && n > (uint32_t)kMaxLeaf)) {         // (smaller nodes go to k_subtrees; the root never does)
        const uint32_t slot = atomicAdd(&ctl->bin_slots, 1u);
        W.bins = slot;
        uint32_t* b = bins + (size_t)slot * kBinWords;
        for (int i = 0; i < 3 * NB; i++) {
            b[7 * i + 0] = kKeyMax; b[7 * i + 1] = kKeyMax; b[7 * i + 2] = kKeyMax;
            b[7 * i + 3] = 0; b[7 * i + 4] = 0; b[7 * i + 5] = 0; b[7 * i + 6] = 0;
        }
    }
}

constexpr int kTopNodesDecl = 32;
// one atomic per distinct target among the lanes of a warp
__device__ __forceinline__ void agg_min(uint32_t* addr, uint32_t v, unsigned peers, bool leader) { v = __reduce_min_sync(peers, v); if (leader) atomicMin(addr, v); }
__device__ __forceinline__ void agg_max(uint32_t* addr, uint32_t v, unsigned peers, bool leader) { v = __reduce_max_sync(peers, v); if (leader) atomicMax(addr, v); }

__global__ void k_bin(const GItem* items, const uint32_t* node_of, uint32_t n, const Work* work, uint32_t* bins, const Ctl* ctl) {
    if (ctl->count <= (uint32_t)kTopNodesDecl || ctl->count == 0) return;      // the top of the tree is k_bin_top's
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t slot = kInvalid;
    GItem it; const Work* W = nullptr;
    if (i < n) {
        const uint32_t w = node_of[i];
        if (w != kInvalid) { W = &work[w]; slot = W->bins; if (slot != kInvalid) it = items[i]; }
    }
    for (int a = 0; a < 3; a++) {
        uint32_t target = kInvalid;
        if (slot != kInvalid) target = slot * (3 * NB) + a * NB + bin_of(centroid(it, a), W->cmin[a], W->scale[a]);
        const unsigned act = __ballot_sync(0xFFFFFFFFu, target != kInvalid);
        if (target == kInvalid) continue;
        const unsigned peers = __match_any_sync(act, target);
        const bool leader = lane == (unsigned)(__ffs(peers) - 1);
        uint32_t* b = bins + (size_t)target * 7;
        agg_min(b + 0, fkey(it.lo[0]), peers, leader); agg_min(b + 1, fkey(it.lo[1]), peers, leader); agg_min(b + 2, fkey(it.lo[2]), peers, leader);
        agg_max(b + 3, fkey(it.hi[0]), peers, leader); agg_max(b + 4, fkey(it.hi[1]), peers, leader); agg_max(b + 5, fkey(it.hi[2]), peers, leader);
        if (leader) atomicAdd(b + 6, (uint32_t)__popc(peers));
    }
}

// Top of the tree: a handful of nodes own all the items, so k_bin's global atomics all land on the same few hundred words.
// Here every block bins into a private copy in shared memory (slots 0 .. kTopNodes-1 of the bin pool) and merges it once.
constexpr int kTopNodes = 32;            // 32 x 336 words = 42 KB of shared memory
static_assert(kTopNodes == kTopNodesDecl, "kTopNodes");
__global__ void __launch_bounds__(256) k_bin_top(const GItem* items, const uint32_t* node_of, uint32_t n, const Work* work, uint32_t* bins, const Ctl* ctl) {
    if (ctl->count > (uint32_t)kTopNodes || ctl->count == 0) return;           // every bin slot < count <= kTopNodes
    __shared__ uint32_t sb[kTopNodes * kBinWords];
    for (int i = threadIdx.x; i < kTopNodes * kBinWords; i += blockDim.x) sb[i] = (i % 7) < 3 ? kKeyMax : 0u;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t w = node_of[i];
        if (w == kInvalid) continue;
        const Work& W = work[w];
        if (W.bins == kInvalid) continue;
        const GItem it = items[i];
        // shared-memory atomics take the lanes' conflicts as they come: cheaper than sorting the lanes into groups first (match_any + six
        // reductions per group, the form the global atomics of k_bin need)
        for (int a = 0; a < 3; a++) {
            uint32_t* b = sb + (size_t)(W.bins * (3 * NB) + a * NB + bin_of(centroid(it, a), W.cmin[a], W.scale[a])) * 7;
            atomicMin(b + 0, fkey(it.lo[0])); atomicMin(b + 1, fkey(it.lo[1])); atomicMin(b + 2, fkey(it.lo[2]));
            atomicMax(b + 3, fkey(it.hi[0])); atomicMax(b + 4, fkey(it.hi[1])); atomicMax(b + 5, fkey(it.hi[2]));
            atomicAdd(b + 6, 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTopNodes * kBinWords; i += blockDim.x) {
        const uint32_t v = sb[i];
        const int f = i % 7;
        if (f < 3) { if (v != kKeyMax) atomicMin(bins + i, v); }
        else if (f < 6) { if (v != 0u) atomicMax(bins + i, v); }
        else if (v) atomicAdd(bins + i, v);
    }
}

struct FBox {
    float lo[3], hi[3];
    __device__ void reset() { for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    __device__ void grow(const float* l, const float* h) { for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]); } }
    __device__ float half_area() const { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + x * z + y * z; }
};
__device__ __forceinline__ void bin_box(const uint32_t* b, float* lo, float* hi) { for (int k = 0; k < 3; k++) { lo[k] = kfloat(b[k]); hi[k] = kfloat(b[3 + k]); } }

// SAH sweep over the bins of one axis: the best split so far is replaced only by a strictly cheaper one
__device__ __forceinline__ void sah_axis(const uint32_t* B, int a, float& best, int& best_axis, int& best_bin) {
    float la[NB]; uint32_t lc[NB];
    FBox acc; acc.reset(); uint32_t c = 0;
    for (int k = 0; k < NB; k++) {
        const uint32_t* b = B + (a * NB + k) * 7;
        if (b[6]) { float lo[3], hi[3]; bin_box(b, lo, hi); acc.grow(lo, hi); }
        c += b[6]; la[k] = c ? acc.half_area() : 0.0f; lc[k] = c;
    }
    acc.reset(); c = 0;
    for (int k = NB - 1; k >= 1; k--) {
        const uint32_t* b = B + (a * NB + k) * 7;
        if (b[6]) { float lo[3], hi[3]; bin_box(b, lo, hi); acc.grow(lo, hi); }
        c += b[6];
        if (!c || !lc[k - 1]) continue;
        const float cost = la[k - 1] * (float)lc[k - 1] + acc.half_area() * (float)c;
        if (cost < best) { best = cost; best_axis = a; best_bin = k - 1; }
    }
}

// The same sweep by a warp: lane k (and its mirror k + 16) owns bin k of axis a, prefix and suffix boxes come from shuffle scans.  It finds
// what sah_axis finds: min / max are exact in any order, the cost is the same expression of the same operands, and among equal costs the
// larger bin wins, as it does in the descending loop above.
__device__ __forceinline__ void sah_axis_warp(const uint32_t* B, int a, unsigned lane, float& best, int& best_axis, int& best_bin) {
    const unsigned k = lane & 15u;
    const uint32_t* b = B + (a * NB + k) * 7;
    const uint32_t cnt = b[6];
    FBox pre, suf; pre.reset();
    if (cnt) bin_box(b, pre.lo, pre.hi);
    suf = pre;
    uint32_t pc = cnt, sc = cnt;
    for (unsigned d = 1; d < 16; d <<= 1) {
        for (int q = 0; q < 3; q++) {
            const float pl = __shfl_up_sync(0xFFFFFFFFu, pre.lo[q], d, 16), ph = __shfl_up_sync(0xFFFFFFFFu, pre.hi[q], d, 16);
            const float sl = __shfl_down_sync(0xFFFFFFFFu, suf.lo[q], d, 16), sh = __shfl_down_sync(0xFFFFFFFFu, suf.hi[q], d, 16);
            if (k >= d) { pre.lo[q] = fminf(pre.lo[q], pl); pre.hi[q] = fmaxf(pre.hi[q], ph); }
            if (k + d < 16u) { suf.lo[q] = fminf(suf.lo[q], sl); suf.hi[q] = fmaxf(suf.hi[q], sh); }
        }
        const uint32_t c0 = __shfl_up_sync(0xFFFFFFFFu, pc, d, 16), c1 = __shfl_down_sync(0xFFFFFFFFu, sc, d, 16);
        if (k >= d) pc += c0;
        if (k + d < 16u) sc += c1;
    }
    const float pa = pc ? pre.half_area() : 0.0f;
    const float la = __shfl_up_sync(0xFFFFFFFFu, pa, 1, 16);                 // prefix over bins 0 .. k-1
    const uint32_t lc = __shfl_up_sync(0xFFFFFFFFu, pc, 1, 16);
    float c = INFINITY; int kk = -1;
    if (k >= 1u && sc && lc) {
        const float cost = la * (float)lc + suf.half_area() * (float)sc;
        if (cost < INFINITY) { c = cost; kk = (int)k; }
    }
    for (unsigned d = 8; d > 0; d >>= 1) {
        const float oc = __shfl_xor_sync(0xFFFFFFFFu, c, d); const int ok = __shfl_xor_sync(0xFFFFFFFFu, kk, d);
        if (oc < c || (oc == c && ok > kk)) { c = oc; kk = ok; }
    }
    if (kk >= 0 && c < best) { best = c; best_axis = a; best_bin = kk - 1; }
}

__device__ __forceinline__ void set_child_word(HostNode* nodes, uint32_t parent, uint32_t which, uint32_t word) {
    if (parent == kInvalid) return;
    if (which) nodes[parent].c1 = word; else nodes[parent].c0 = word;
}

__global__ void k_split(Work* work, const uint32_t* bins, const GItem* items, HostNode* nodes, Work* next, Ctl* ctl, uint32_t next_cap, SubRoot* subs) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ctl->count) return;
    Work& W = work[w];
    const uint32_t n = W.end - W.begin;
    W.child[0] = W.child[1] = kInvalid;
    if (n <= kSub && W.parent != kInvalid) {                   // its items stay where they are (k_scatter), a warp of k_subtrees builds the rest
        W.mode = 4;
        subs[atomicAdd(&ctl->sub_count, 1u)] = SubRoot{W.begin, W.end, W.parent, W.which, W.depth};
        return;
    }
    FBox lbox, rbox; lbox.reset(); rbox.reset();
    uint32_t nl = 0;
    if (n <= (uint32_t)kMaxLeaf) {
        bool mixed = false; uint32_t tmin = items[W.begin].type;
        for (uint32_t i = 1; i < n; i++) { const uint32_t t = items[W.begin + i].type; mixed |= t != items[W.begin].type; tmin = min(tmin, t); }
        if (!mixed && W.parent != kInvalid) {                   // a leaf: its word goes into the parent's slot
            W.mode = 3;
            set_child_word(nodes, W.parent, W.which, kLeafBit | (tmin << 29) | ((n - 1) << 24) | W.begin);
            atomicMax(&ctl->max_depth, W.depth);
            return;
        }
        if (mixed) {                                            // small but mixed types: the lowest type goes left
            W.mode = 1; W.tmin = tmin;
            for (uint32_t i = 0; i < n; i++) {
                const GItem& it = items[W.begin + i];
                if (it.type == tmin) { nl++; lbox.grow(it.lo, it.hi); } else rbox.grow(it.lo, it.hi);
            }
        } else {                                                // the whole (tiny, homogeneous) scene: the caller special-cases it
            W.mode = 3;
            return;
        }
    } else {
        const uint32_t* B = bins + (size_t)W.bins * kBinWords;
        float best = INFINITY; int best_axis = -1, best_bin = -1;
        if (W.depth < 40) {
            for (int a = 0; a < 3; a++) {
                if (W.scale[a] == 0.0f) continue;
                sah_axis(B, a, best, best_axis, best_bin);
            }
        }
        if (best_axis >= 0) {
            W.mode = 0; W.axis = best_axis; W.bin = best_bin;
            for (int k = 0; k < NB; k++) {
                const uint32_t* b = B + (best_axis * NB + k) * 7;
                if (!b[6]) continue;
                float lo[3], hi[3]; bin_box(b, lo, hi);
                if (k <= best_bin) { lbox.grow(lo, hi); nl += b[6]; } else rbox.grow(lo, hi);
            }
        } else {                                                // coincident centroids or depth guard: halve by position
            W.mode = 2; nl = n / 2;
            for (int k = 0; k < NB; k++) {                       // any axis' bins cover all items; both children get the whole box
                const uint32_t* b = B + k * 7;
                if (!b[6]) continue;
                float lo[3], hi[3]; bin_box(b, lo, hi); lbox.grow(lo, hi);
            }
            bool any = false; for (int k = 0; k < NB; k++) any |= B[k * 7 + 6] != 0;
            if (!any) for (uint32_t i = 0; i < n; i++) lbox.grow(items[W.begin + i].lo, items[W.begin + i].hi);
            rbox = lbox;
        }
    }
    W.nl = nl;
    // this node becomes an interior node of the output
    const uint32_t me = W.parent == kInvalid ? 0u : atomicAdd(&ctl->node_count, 1u);
    set_child_word(nodes, W.parent, W.which, me);
    HostNode& nd = nodes[me];
    for (int k = 0; k < 3; k++) { nd.v[k] = lbox.lo[k]; nd.v[3 + k] = lbox.hi[k]; nd.v[6 + k] = rbox.lo[k]; nd.v[9 + k] = rbox.hi[k]; }
    nd.pad0 = nd.pad1 = 0;
    const uint32_t base = atomicAdd(&ctl->next_count, 2u);
    if (base + 2 > next_cap) return;                            // cannot happen: next_cap = item count (every node holds >= 1 item)
    for (uint32_t c = 0; c < 2; c++) {
        Work& C = next[base + c];
        C.begin = c ? W.begin + nl : W.begin; C.end = c ? W.end : W.begin + nl;
        C.parent = me; C.which = c; C.depth = W.depth + 1;
        for (int a = 0; a < 3; a++) { C.cb_lo[a] = kKeyMax; C.cb_hi[a] = 0; }
        W.child[c] = base + c;
    }
}

// The top of the tree once more: with a handful of nodes owning all the items, the cursors and child bounds of k_scatter are a few dozen
// words that every warp of the grid hits with global atomics (level 0: 31 k warps x 14 atomics on 14 addresses).  Here a block of 1024
// threads settles its items among themselves in shared memory and goes to global memory once per (node, side) it holds.
constexpr int kScatterTopThreads = 1024;
__global__ void __launch_bounds__(kScatterTopThreads) k_scatter_top(const GItem* items, uint32_t* node_of, uint32_t n, Work* work, Work* next, GItem* items_out,
                                                                     uint32_t* node_of_out, GItem* final_items, const Ctl* ctl) {
    if (ctl->count > (uint32_t)kTopNodes || ctl->count == 0) return;
    __shared__ uint32_t s_cnt[2 * kTopNodes], s_base[2 * kTopNodes], s_lo[2 * kTopNodes][3], s_hi[2 * kTopNodes][3];
    for (int k = threadIdx.x; k < 2 * kTopNodes; k += blockDim.x) { s_cnt[k] = 0; for (int a = 0; a < 3; a++) { s_lo[k][a] = kKeyMax; s_hi[k][a] = 0; } }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t w = kInvalid;
    if (i < n) w = node_of[i];
    GItem it; bool live = false, left = false; uint32_t rank = 0, key = 0;
    if (w != kInvalid) {
        const Work& W = work[w];
        it = items[i];
        if (W.mode >= 3) { final_items[i] = it; node_of_out[i] = kInvalid; node_of[i] = kInvalid; }
        else {
            live = true;
            if (W.mode == 0) left = bin_of(centroid(it, W.axis), W.cmin[W.axis], W.scale[W.axis]) <= (int)W.bin;
            else if (W.mode == 1) left = it.type == W.tmin;
            else left = i - W.begin < W.nl;
            key = w * 2u + (left ? 0u : 1u);
        }
    }
    const unsigned act = __ballot_sync(0xFFFFFFFFu, live);
    if (live) {
        const unsigned peers = __match_any_sync(act, key);
        const int leader = __ffs(peers) - 1;
        const bool lead = (int)lane == leader;
        uint32_t b0 = 0;
        if (lead) b0 = atomicAdd(&s_cnt[key], (uint32_t)__popc(peers));
        rank = __shfl_sync(peers, b0, leader) + __popc(peers & ((1u << lane) - 1u));
        for (int a = 0; a < 3; a++) {
            const uint32_t k = fkey(centroid(it, a));
            agg_min(&s_lo[key][a], k, peers, lead);
            agg_max(&s_hi[key][a], k, peers, lead);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kTopNodes && s_cnt[threadIdx.x]) {
        const uint32_t k = threadIdx.x;
        Work& W = work[k >> 1];
        s_base[k] = atomicAdd((k & 1u) ? &W.cur_r : &W.cur_l, s_cnt[k]);
        Work& Cn = next[W.child[k & 1u]];
        for (int a = 0; a < 3; a++) { atomicMin(&Cn.cb_lo[a], s_lo[k][a]); atomicMax(&Cn.cb_hi[a], s_hi[k][a]); }
    }
    __syncthreads();
    if (live) {
        const Work& W = work[w];
        const uint32_t dest = (left ? W.begin : W.begin + W.nl) + s_base[key] + rank;
        items_out[dest] = it;
        node_of_out[dest] = W.child[left ? 0 : 1];
    }
}

__global__ void k_scatter(const GItem* items, uint32_t* node_of, uint32_t n, Work* work, Work* next, GItem* items_out, uint32_t* node_of_out, GItem* final_items, const Ctl* ctl) {
    if (ctl->count <= (uint32_t)kTopNodes) return;             // the top of the tree is k_scatter_top's
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t w = kInvalid;
    if (i < n) w = node_of[i];
    const unsigned act = __ballot_sync(0xFFFFFFFFu, w != kInvalid);
    if (w == kInvalid) return;
    Work& W = work[w];
    const GItem it = items[i];
    const bool is_leaf = W.mode >= 3;                           // a leaf, or the root of a sub-tree k_subtrees finishes in place
    const unsigned live = __ballot_sync(act, !is_leaf);
    if (is_leaf) { final_items[i] = it; node_of_out[i] = kInvalid; node_of[i] = kInvalid; return; }   // leaf: the item is where it stays; retired in BOTH buffers
    bool left;
    if (W.mode == 0) left = bin_of(centroid(it, W.axis), W.cmin[W.axis], W.scale[W.axis]) <= (int)W.bin;
    else if (W.mode == 1) left = it.type == W.tmin;
    else left = i - W.begin < W.nl;
    // warp-aggregated cursor: lanes of the same node and side take consecutive slots with one atomic
    const uint32_t key = w * 2u + (left ? 0u : 1u);
    const unsigned peers = __match_any_sync(live, key);
    const int leader = __ffs(peers) - 1;
    uint32_t basepos = 0;
    if ((int)lane == leader) basepos = atomicAdd(left ? &W.cur_l : &W.cur_r, (uint32_t)__popc(peers));
    basepos = __shfl_sync(peers, basepos, leader);
    const uint32_t dest = (left ? W.begin : W.begin + W.nl) + basepos + __popc(peers & ((1u << lane) - 1u));
    const uint32_t child = W.child[left ? 0 : 1];
    items_out[dest] = it;
    node_of_out[dest] = child;
    Work& C = next[child];
    const bool lead = (int)lane == leader;
    for (int a = 0; a < 3; a++) {
        const uint32_t k = fkey(centroid(it, a));
        agg_min(&C.cb_lo[a], k, peers, lead);
        agg_max(&C.cb_hi[a], k, peers, lead);
    }
}

// One warp per sub-tree root of at most kSub items: the items sit in shared memory, a permutation of 16-bit indices is partitioned
// node by node (depth first, explicit stack), and every decision is k_split's: same centroid bounds (k_scatter grows them from the same
// items), same bin map, same bins (min / max / count do not depend on the order), same sweep.
constexpr int kSubWarps = 4;
// Node indices: a sub-tree writes its nodes to its own run of a scratch array under local indices (at most n - 1 interior nodes for n
// items, so the run [begin, end) of the item range does), takes its place in the output with ONE atomic when it is complete, and copies
// the nodes over with their child words re-based.  One atomic per node on the common counter was the kernel's bound: 570 k same-address
// atomics for a million triangles, 1 ms when issued late, 6 ms when issued early.
constexpr uint32_t kSubRootParent = 0x7FFFFFFFu;       // stack entry of the sub-tree's root: its parent is a node of the level loop
__global__ void __launch_bounds__(32 * kSubWarps) k_subtrees(const SubRoot* subs, GItem* final_items, HostNode* nodes, HostNode* scratch, Ctl* ctl) {
    __shared__ uint4 s_items[kSubWarps][kSub * 2];
    __shared__ uint16_t s_idx[kSubWarps][2][kSub];
    __shared__ uint32_t s_bins[kSubWarps][kBinWords];
    __shared__ uint4 s_stack[kSubWarps][64];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t s = blockIdx.x * kSubWarps + warp;
    if (s >= ctl->sub_count) return;
    const SubRoot R = subs[s];
    const uint32_t n0 = R.end - R.begin;
    const GItem* items = reinterpret_cast<const GItem*>(s_items[warp]);
    uint16_t* idx = s_idx[warp][0]; uint16_t* alt = s_idx[warp][1];
    uint32_t* bins = s_bins[warp];
    uint4* stack = s_stack[warp];
    {
        const uint4* src = reinterpret_cast<const uint4*>(final_items + R.begin);
        for (uint32_t i = lane; i < 2 * n0; i += 32) s_items[warp][i] = src[i];
        for (uint32_t i = lane; i < n0; i += 32) idx[i] = (uint16_t)i;
    }
    if (lane == 0) stack[0] = make_uint4(0u, n0, kSubRootParent, R.depth);
    __syncwarp();
    HostNode* mine = scratch + R.begin;
    uint32_t sp = 1, maxd = 0, n_mine = 0, root_word = 0;
    while (sp) {
        const uint4 top = stack[--sp];
        __syncwarp();
        const uint32_t b = top.x, e = top.y, n = e - b, parent = top.z & 0x7FFFFFFFu, which = top.z >> 31, depth = top.w;
        uint32_t mode = 0, nl = 0, tmin = 0; int axis = -1, bin = -1;
        float cmin[3] = {0.f, 0.f, 0.f}, scale[3] = {0.f, 0.f, 0.f};
        FBox lbox, rbox; lbox.reset(); rbox.reset();
        if (n <= (uint32_t)kMaxLeaf) {
            const uint32_t t = lane < n ? items[idx[b + lane]].type : 0xFFFFFFFFu;
            const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, t, 0);
            const bool mixed = __any_sync(0xFFFFFFFFu, lane < n && t != t0);
            tmin = __reduce_min_sync(0xFFFFFFFFu, t);
            if (!mixed) {
                const uint32_t word = kLeafBit | (tmin << 29) | ((n - 1) << 24) | (R.begin + b);
                if (parent == kSubRootParent) root_word = word;
                else if (lane == 0) set_child_word(mine, parent, which, word);
                maxd = max(maxd, depth);
                continue;
            }
            mode = 1;
            for (uint32_t i = b; i < e; i++) {                   // at most kMaxLeaf items
                const GItem& it = items[idx[i]];
                if (it.type == tmin) { nl++; lbox.grow(it.lo, it.hi); } else rbox.grow(it.lo, it.hi);
            }
        } else {
            uint32_t clo[3] = {kKeyMax, kKeyMax, kKeyMax}, chi[3] = {0, 0, 0};
            for (uint32_t i = b + lane; i < e; i += 32) {
                const GItem& it = items[idx[i]];
                for (int a = 0; a < 3; a++) { const uint32_t k = fkey(centroid(it, a)); clo[a] = min(clo[a], k); chi[a] = max(chi[a], k); }
            }
            for (int a = 0; a < 3; a++) {
                const float lo = kfloat(__reduce_min_sync(0xFFFFFFFFu, clo[a])), hi = kfloat(__reduce_max_sync(0xFFFFFFFFu, chi[a]));
                const float ext = hi - lo;
                cmin[a] = lo;
                scale[a] = ext > 0.0f ? (float)NB * (1.0f - 1e-6f) / ext : 0.0f;
            }
            for (int i = lane; i < kBinWords; i += 32) bins[i] = (i % 7) < 3 ? kKeyMax : 0u;
            __syncwarp();
            for (uint32_t i = b + lane; i < e; i += 32) {
                const GItem& it = items[idx[i]];
                for (int a = 0; a < 3; a++) {
                    uint32_t* q = bins + (a * NB + bin_of(centroid(it, a), cmin[a], scale[a])) * 7;
                    atomicMin(q + 0, fkey(it.lo[0])); atomicMin(q + 1, fkey(it.lo[1])); atomicMin(q + 2, fkey(it.lo[2]));
                    atomicMax(q + 3, fkey(it.hi[0])); atomicMax(q + 4, fkey(it.hi[1])); atomicMax(q + 5, fkey(it.hi[2]));
                    atomicAdd(q + 6, 1u);
                }
            }
            __syncwarp();
            // the axes compete in k_split's order (a later axis wins only if strictly cheaper)
            float best = INFINITY;
            if (depth < 40)
                for (int a = 0; a < 3; a++) if (scale[a] != 0.0f) sah_axis_warp(bins, a, lane, best, axis, bin);
            if (axis >= 0) {
                mode = 0;
                for (int k = 0; k < NB; k++) {
                    const uint32_t* q = bins + (axis * NB + k) * 7;
                    if (!q[6]) continue;
                    float lo[3], hi[3]; bin_box(q, lo, hi);
                    if (k <= bin) { lbox.grow(lo, hi); nl += q[6]; } else rbox.grow(lo, hi);
                }
            } else {                                             // coincident centroids or depth guard: halve by position
                mode = 2; nl = n / 2;
                for (int k = 0; k < NB; k++) {
                    const uint32_t* q = bins + k * 7;
                    if (!q[6]) continue;
                    float lo[3], hi[3]; bin_box(q, lo, hi); lbox.grow(lo, hi);
                }
                rbox = lbox;
            }
        }
        const uint32_t me = n_mine++;                            // local index: the sub-tree's root is 0
        if (lane == 0) {
            if (parent != kSubRootParent) set_child_word(mine, parent, which, me);
            HostNode& nd = mine[me];
            for (int k = 0; k < 3; k++) { nd.v[k] = lbox.lo[k]; nd.v[3 + k] = lbox.hi[k]; nd.v[6 + k] = rbox.lo[k]; nd.v[9 + k] = rbox.hi[k]; }
            nd.pad0 = nd.pad1 = 0;
        }
        // stable partition of the node's index range
        uint32_t lc = 0, rc = 0;
        for (uint32_t base = b; base < e; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < e;
            const uint16_t id = valid ? idx[i] : (uint16_t)0;
            bool left = false;
            if (valid) {
                const GItem& it = items[id];
                if (mode == 0) left = bin_of(centroid(it, axis), cmin[axis], scale[axis]) <= bin;
                else if (mode == 1) left = it.type == tmin;
                else left = i - b < nl;
            }
            const unsigned lm = __ballot_sync(0xFFFFFFFFu, left), rm = __ballot_sync(0xFFFFFFFFu, valid && !left);
            if (left) alt[b + lc + __popc(lm & lt)] = id;
            else if (valid) alt[b + nl + rc + __popc(rm & lt)] = id;
            lc += __popc(lm); rc += __popc(rm);
        }
        __syncwarp();
        for (uint32_t i = b + lane; i < e; i += 32) idx[i] = alt[i];
        if (sp + 2 > 64u) { maxd = 0xFFFFu; break; }            // cannot happen below the depth guard; the caller rejects the tree
        if (lane == 0) {
            stack[sp] = make_uint4(b + nl, e, me | (1u << 31), depth + 1);
            stack[sp + 1] = make_uint4(b, b + nl, me, depth + 1);
        }
        sp += 2;
        __syncwarp();
    }
    __syncwarp();
    {
        uint4* dst = reinterpret_cast<uint4*>(final_items + R.begin);
        for (uint32_t i = lane; i < 2 * n0; i += 32) dst[i] = s_items[warp][2 * idx[i >> 1] + (i & 1)];
    }
    // the sub-tree's place in the output, its nodes with re-based child words, and its own word in the parent
    uint32_t base = 0;
    if (lane == 0 && n_mine) base = atomicAdd(&ctl->node_count, n_mine);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    {
        const uint4* src = reinterpret_cast<const uint4*>(mine);
        uint4* dst = reinterpret_cast<uint4*>(nodes + base);
        for (uint32_t i = lane; i < 4 * n_mine; i += 32) {
            uint4 v = src[i];
            if ((i & 3u) == 3u) { if (!(v.x & kLeafBit)) v.x += base; if (!(v.y & kLeafBit)) v.y += base; }
            dst[i] = v;
        }
    }
    if (lane == 0) {
        set_child_word(nodes, R.parent, R.which, n_mine ? base : root_word);
        atomicMax(&ctl->max_depth, maxd);
    }
}

__global__ void k_fill(uint32_t* p, uint32_t n, uint32_t v) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

struct TypeIs {
    uint32_t t;
    __host__ __device__ uint32_t operator()(const GItem& it) const { return it.type == t ? 1u : 0u; }
};

// leaf words carry the position of the leaf's first item; the primitive arrays are per type, so it becomes the
// rank of that item among the items of its type (tree order)
__global__ void k_patch_leaves(HostNode* nodes, uint32_t n_nodes, const uint32_t* const* typepos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    for (int c = 0; c < 2; c++) {
        uint32_t w = c ? nodes[i].c1 : nodes[i].c0;
        if (!(w & kLeafBit)) continue;
        const uint32_t type = (w >> 29) & 3u, pos = w & kLeafFirstMask;
        w = (w & ~kLeafFirstMask) | typepos[type][pos];
        if (c) nodes[i].c1 = w; else nodes[i].c0 = w;
    }
}

__device__ __forceinline__ float pad_down(float v, float pad) { return nextafterf(v - pad, -INFINITY); }
__device__ __forceinline__ float pad_up(float v, float pad) { return nextafterf(v + pad, INFINITY); }

__global__ void k_make_items(RawScene R, float pad, GItem* items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = R.n_spheres + R.n_cuboids + R.n_triangles;
    if (i >= n) return;
    GItem it;
    if (i < R.n_spheres) {
        const lgb_sphere sp = R.spheres[i];
        it.type = LGB_PRIM_SPHERE; it.index = i;
        for (int k = 0; k < 3; k++) {
            const double lo = sp.center[k] - sp.radius, hi = sp.center[k] + sp.radius;
            it.lo[k] = pad_down(__double2float_rd(fmin(lo, hi)), pad); it.hi[k] = pad_up(__double2float_ru(fmax(lo, hi)), pad);
        }
    } else if (i < R.n_spheres + R.n_cuboids) {
        const uint32_t j = i - R.n_spheres;
        const lgb_cuboid c = R.cuboids[j];
        it.type = LGB_PRIM_CUBOID; it.index = j;
        for (int k = 0; k < 3; k++) {
            it.lo[k] = pad_down(__double2float_rd(fmin(c.min[k], c.max[k])), pad); it.hi[k] = pad_up(__double2float_ru(fmax(c.min[k], c.max[k])), pad);
        }
    } else {
        const uint32_t j = i - R.n_spheres - R.n_cuboids;
        const lgb_triangle t = R.triangles[j];
        it.type = LGB_PRIM_TRIANGLE; it.index = j;
        for (int k = 0; k < 3; k++) {
            it.lo[k] = pad_down(fminf(t.p0[k], fminf(t.p1[k], t.p2[k])), pad); it.hi[k] = pad_up(fmaxf(t.p0[k], fmaxf(t.p1[k], t.p2[k])), pad);
        }
    }
    items[i] = it;
}

struct TypePos { const uint32_t* p[3]; };
__global__ void k_convert(RawScene R, LeafArrays O, const GItem* final_items, uint32_t n, TypePos tp, double cpad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t type = final_items[i].type, src = final_items[i].index;
    const uint32_t j = tp.p[type][i];
    if (type == LGB_PRIM_SPHERE) {
        const lgb_sphere sp = R.spheres[src];
        O.sph32[j] = make_float4((float)sp.center[0], (float)sp.center[1], (float)sp.center[2], __double2float_ru(fabs(sp.radius)));
        O.sph64[4 * (size_t)j] = sp.center[0]; O.sph64[4 * (size_t)j + 1] = sp.center[1]; O.sph64[4 * (size_t)j + 2] = sp.center[2]; O.sph64[4 * (size_t)j + 3] = sp.radius;
        O.sph_mat[j] = R.sphere_material[src]; O.sph_id[j] = R.sphere_id[src];
    } else if (type == LGB_PRIM_CUBOID) {
        const lgb_cuboid c = R.cuboids[src];
        O.cub32[2 * (size_t)j] = make_float4(__double2float_rd(c.min[0] - cpad), __double2float_rd(c.min[1] - cpad), __double2float_rd(c.min[2] - cpad), 0.f);
        O.cub32[2 * (size_t)j + 1] = make_float4(__double2float_ru(c.max[0] + cpad), __double2float_ru(c.max[1] + cpad), __double2float_ru(c.max[2] + cpad), 0.f);
        for (int k = 0; k < 3; k++) { O.cub64[6 * (size_t)j + k] = c.min[k]; O.cub64[6 * (size_t)j + 3 + k] = c.max[k]; }
        O.cub_mat[j] = R.cuboid_material[src]; O.cub_id[j] = R.cuboid_id[src];
    } else {
        const lgb_triangle t = R.triangles[src];
        uint32_t ni = 0xFFFFFFFFu;                                   // kNoNormals
        if (R.tri_normals && (!R.tri_has_normals || R.tri_has_normals[src])) {
            ni = j;
            const lgb_tri_normals q = R.tri_normals[src];
            float* o = O.tri_nrm + 9 * (size_t)j;
            for (int k = 0; k < 3; k++) { o[k] = q.n0[k]; o[3 + k] = q.n1[k]; o[6 + k] = q.n2[k]; }
        }
        O.tri[3 * (size_t)j] = make_float4(t.p0[0], t.p0[1], t.p0[2], __uint_as_float(R.triangle_id[src]));
        O.tri[3 * (size_t)j + 1] = make_float4(t.p1[0], t.p1[1], t.p1[2], __uint_as_float(R.triangle_material[src]));
        O.tri[3 * (size_t)j + 2] = make_float4(t.p2[0], t.p2[1], t.p2[2], __uint_as_float(ni));
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t gpu_build_temp_bytes(uint32_t n) {
    size_t b = 0;
    b += 2 * align256((size_t)n * sizeof(GItem));            // ping / pong
    b += 2 * align256((size_t)n * 4);                        // node_of ping / pong
    b += 2 * align256((size_t)(n + 2) * sizeof(Work));       // work lists of two levels
    b += align256((size_t)(n + 2) * sizeof(SubRoot));        // sub-tree roots (k_subtrees)
    b += align256((size_t)n * sizeof(HostNode));             // ... and their nodes under local indices
    b += align256(((size_t)n / (kMaxLeaf + 1) + 2) * kBinWords * 4);   // bin pool: nodes of more than kMaxLeaf items
    b += align256(sizeof(Ctl)) + align256(4 * sizeof(void*));
    b += 3 * align256((size_t)n * 4);                        // per-type positions
    size_t scan = 0;
    auto in = thrust::make_transform_iterator((const GItem*)nullptr, TypeIs{0});
    cub::DeviceScan::ExclusiveSum(nullptr, scan, in, (uint32_t*)nullptr, (int)n);
    b += align256(scan);
    return b;
}

// The per-type positions of the items in the order they are given (no tree: the deferred-BVH scenes of lgb_api.cu store their
// primitives in the caller's order until a tree is asked for).  typepos_out as from gpu_build_sah; `temp` of gpu_build_temp_bytes(n).
cudaError_t gpu_identity_positions(const GItem* items, uint32_t n, uint32_t** typepos_out, void* temp, size_t temp_bytes, cudaStream_t st) {
    if (temp_bytes < gpu_build_temp_bytes(n)) return cudaErrorInvalidValue;
    char* p = (char*)temp;
    auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
    uint32_t* typepos[3] = {(uint32_t*)take((size_t)n * 4), (uint32_t*)take((size_t)n * 4), (uint32_t*)take((size_t)n * 4)};
    size_t scan_bytes = 0;
    {
        auto in = thrust::make_transform_iterator((const GItem*)nullptr, TypeIs{0});
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, in, (uint32_t*)nullptr, (int)n);
    }
    void* scan_tmp = take(scan_bytes);
    cudaError_t e;
    for (uint32_t t = 0; t < 3; t++) {
        auto in = thrust::make_transform_iterator(items, TypeIs{t});
        if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, in, typepos[t], (int)n, st)) != cudaSuccess) return e;
    }
    for (int t = 0; t < 3; t++) typepos_out[t] = typepos[t];
    return cudaSuccess;
}

cudaError_t gpu_build_sah(const GItem* items_in, uint32_t n, HostNode* nodes_out, GItem* final_items, uint32_t** typepos_out, void* temp, size_t temp_bytes,
                          cudaStream_t st, GpuBuildInfo* info) {
    if (n <= (uint32_t)kMaxLeaf || temp_bytes < gpu_build_temp_bytes(n)) return cudaErrorInvalidValue;
    char* p = (char*)temp;
    auto take = [&](size_t bytes) { char* r = p; p += align256(bytes); return r; };
    GItem* buf[2] = {(GItem*)take((size_t)n * sizeof(GItem)), (GItem*)take((size_t)n * sizeof(GItem))};
    uint32_t* nof[2] = {(uint32_t*)take((size_t)n * 4), (uint32_t*)take((size_t)n * 4)};
    Work* work[2] = {(Work*)take((size_t)(n + 2) * sizeof(Work)), (Work*)take((size_t)(n + 2) * sizeof(Work))};
    SubRoot* subs = (SubRoot*)take((size_t)(n + 2) * sizeof(SubRoot));
    HostNode* sub_nodes = (HostNode*)take((size_t)n * sizeof(HostNode));
    uint32_t* bins = (uint32_t*)take(((size_t)n / (kMaxLeaf + 1) + 2) * kBinWords * 4);
    Ctl* ctl = (Ctl*)take(sizeof(Ctl));
    uint32_t** d_typepos = (uint32_t**)take(4 * sizeof(void*));
    uint32_t* typepos[3] = {(uint32_t*)take((size_t)n * 4), (uint32_t*)take((size_t)n * 4), (uint32_t*)take((size_t)n * 4)};
    size_t scan_bytes = 0;
    {
        auto in = thrust::make_transform_iterator((const GItem*)nullptr, TypeIs{0});
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, in, (uint32_t*)nullptr, (int)n);
    }
    void* scan_tmp = take(scan_bytes);

    cudaError_t e;
    const bool dbg = std::getenv("LGB_GPUBUILD_DEBUG") != nullptr;       // synchronise after every kernel: finds the one that faults, and says
    double k_ms[2][4] = {};                                              // where the levels' time goes (top 8 levels | the rest) x (prep, bin, split, scatter)
    auto t_k = std::chrono::steady_clock::now();
    auto check = [&](const char* what, uint32_t level) {
        if (!dbg) return cudaSuccess;
        cudaError_t ce = cudaStreamSynchronize(st);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) std::fprintf(stderr, "[gpu_build_sah] %s failed at level %u: %s\n", what, level, cudaGetErrorString(ce));
        const auto t = std::chrono::steady_clock::now();
        const int which = what[2] == 'p' ? 0 : what[2] == 'b' ? 1 : what[3] == 'p' ? 2 : 3;       // k_prep, k_bin, k_split, k_scatter
        k_ms[level < 8 ? 0 : 1][which] += std::chrono::duration<double, std::milli>(t - t_k).count();
        t_k = t;
        return ce;
    };
    const unsigned ib = (n + 255) / 256;
    // level 0: everything belongs to the root
    Work root{};
    root.begin = 0; root.end = n; root.parent = kInvalid; root.which = 0; root.depth = 0;
    for (int a = 0; a < 3; a++) { root.cb_lo[a] = kKeyMax; root.cb_hi[a] = 0; }
    if ((e = cudaMemcpyAsync(work[0], &root, sizeof root, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    Ctl c0{1, 1, 0, 0, 0, 0, 0};                           // node 0 is the root; one node (the root) enters the first level
    if ((e = cudaMemcpyAsync(ctl, &c0, sizeof c0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(buf[0], items_in, (size_t)n * sizeof(GItem), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    k_fill<<<ib, 256, 0, st>>>(nof[0], n, 0u);
    k_root<<<256, 256, 0, st>>>(buf[0], n, work[0], ctl);
    if ((e = check("k_root", 0)) != cudaSuccess) return e;
    uint32_t levels = 0, node_count = 1, max_depth = 0, launched = 0, sub_count = 0;
    int cur = 0;
    const bool timing = std::getenv("LGB_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    if (timing) cudaStreamSynchronize(st);                 // (the caller's upload is still in flight on this stream: keep it out of the first chunk)
    auto t_start = now();
    std::vector<double> chunk_ms;
    // a level never holds more nodes than there are items above leaf size ... bounded by the item count
    const unsigned wb = (unsigned)(((size_t)n + 2 + 127) / 128);
    constexpr uint32_t kChunk = 8;                         // levels queued between two looks at the control block
    for (;;) {
        auto t_chunk = now();
        for (uint32_t k = 0; k < kChunk; k++) {
            k_advance<<<1, 1, 0, st>>>(ctl);
            k_prep<<<wb, 128, 0, st>>>(work[cur], bins, ctl);
            if ((e = check("k_prep", launched)) != cudaSuccess) return e;
            k_bin_top<<<296, 256, 0, st>>>(buf[cur], nof[cur], n, work[cur], bins, ctl);
            k_bin<<<ib, 256, 0, st>>>(buf[cur], nof[cur], n, work[cur], bins, ctl);
            if ((e = check("k_bin", launched)) != cudaSuccess) return e;
            k_split<<<wb, 128, 0, st>>>(work[cur], bins, buf[cur], nodes_out, work[cur ^ 1], ctl, n + 2, subs);
            if ((e = check("k_split", launched)) != cudaSuccess) return e;
            k_scatter_top<<<(n + kScatterTopThreads - 1) / kScatterTopThreads, kScatterTopThreads, 0, st>>>(buf[cur], nof[cur], n, work[cur], work[cur ^ 1], buf[cur ^ 1], nof[cur ^ 1], final_items, ctl);
            k_scatter<<<ib, 256, 0, st>>>(buf[cur], nof[cur], n, work[cur], work[cur ^ 1], buf[cur ^ 1], nof[cur ^ 1], final_items, ctl);
            if ((e = check("k_scatter", launched)) != cudaSuccess) return e;
            cur ^= 1; launched++;
        }
        Ctl h;
        if ((e = cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        if (timing) chunk_ms.push_back(std::chrono::duration<double, std::milli>(now() - t_chunk).count());
        node_count = h.node_count; max_depth = h.max_depth; levels = h.levels;
        if (h.next_count == 0) { sub_count = h.sub_count; break; }
        if (launched > 96) return cudaErrorUnknown;         // depth guard (splits by position halve every node from depth 40 on)
    }
    double sub_ms = 0.0;
    if (sub_count) {                                        // every node the loop left at <= kSub items: one warp each, to the leaves
        auto t_sub = now();
        k_subtrees<<<(sub_count + kSubWarps - 1) / kSubWarps, 32 * kSubWarps, 0, st>>>(subs, final_items, nodes_out, sub_nodes, ctl);
        Ctl h;
        if ((e = cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        node_count = h.node_count; max_depth = h.max_depth;
        sub_ms = std::chrono::duration<double, std::milli>(now() - t_sub).count();
    }
    // per-type positions of the items in tree order; leaf words -> first index in the type's array
    for (uint32_t t = 0; t < 3; t++) {
        auto in = thrust::make_transform_iterator((const GItem*)final_items, TypeIs{t});
        if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, in, typepos[t], (int)n, st)) != cudaSuccess) return e;
    }
    uint32_t* h_tp[4] = {typepos[0], typepos[1], typepos[2], nullptr};
    if ((e = cudaMemcpyAsync(d_typepos, h_tp, sizeof h_tp, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    k_patch_leaves<<<(node_count + 255) / 256, 256, 0, st>>>(nodes_out, node_count, d_typepos);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;      // h_tp is on this stack frame
    if (timing) {
        std::fprintf(stderr, "[gpu_build_sah] %u items, %u levels (%u launched), %.2f ms total; per chunk of %u levels:", n, levels, launched, std::chrono::duration<double, std::milli>(now() - t_start).count(), kChunk);
        for (size_t i = 0; i < chunk_ms.size(); i++) std::fprintf(stderr, " %.2f", chunk_ms[i]);
        std::fprintf(stderr, "; %u sub-trees of <= %u items %.2f ms\n", sub_count, kSub, sub_ms);
        if (dbg) for (int h = 0; h < 2; h++) std::fprintf(stderr, "[gpu_build_sah]   %s: prep %.2f bin %.2f split %.2f scatter %.2f ms (synchronised after every kernel)\n", h ? "levels 8.." : "levels 0..7",
                                                          k_ms[h][0], k_ms[h][1], k_ms[h][2], k_ms[h][3]);
    }
    for (int t = 0; t < 3; t++) typepos_out[t] = typepos[t];
    info->n_nodes = node_count; info->max_depth = max_depth; info->levels = levels;
    return cudaSuccess;
}

cudaError_t launch_make_items(const RawScene& raw, float pad, GItem* items, cudaStream_t stream) {
    const uint32_t n = raw.n_spheres + raw.n_cuboids + raw.n_triangles;
    if (!n) return cudaSuccess;
    k_make_items<<<(n + 255) / 256, 256, 0, stream>>>(raw, pad, items);
    return cudaGetLastError();
}
cudaError_t launch_convert(const RawScene& raw, const LeafArrays& out, const GItem* final_items, uint32_t n, uint32_t* const typepos[3], double cuboid_pad,
                           cudaStream_t stream) {
    if (!n) return cudaSuccess;
    TypePos tp; for (int t = 0; t < 3; t++) tp.p[t] = typepos[t];
    k_convert<<<(n + 255) / 256, 256, 0, stream>>>(raw, out, final_items, n, tp, cuboid_pad);
    return cudaGetLastError();
}

}  // namespace lgb
