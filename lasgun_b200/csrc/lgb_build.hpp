// Device acceleration structure built inside lgb_scene_create (host code, part of the product).
//
// The caller hands over the reference's own BVH (lgb_scene_desc).  That tree partitions only in
// (y, z) and keeps up to 254 primitives per leaf, so the device traverses its own
// binned-SAH BVH over the same primitives instead.  Closest-hit results do not depend on the BVH
// except for exact-t ties, which the reference resolves by "first primitive tested wins"
// (sphere.rs:86, cuboid.rs:95, triangle.rs:251).  The reference's test order is a fixed function
// of the ray-direction octant (bvh.rs:463, :496), so build_rank_tables() records, per octant, the
// position of every primitive in the reference's traversal order; the kernel consults it only when
// two candidates have bit-identical t.  The device result is then the reference's for ANY device BVH.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/lasgun_b200.h"

namespace lgb {

struct HostNode { float v[12]; uint32_t c0, c1, pad0, pad1; };   // 64 B: child0 {lo, hi}, child1 {lo, hi}, child words

// child word: interior = node index; leaf = kLeafBit | type << 29 | (count - 1) << 24 | first
#ifndef LGB_LEAF_CONSTANTS
#define LGB_LEAF_CONSTANTS
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kLeafFirstMask = 0x00FFFFFFu;
#endif
constexpr int kMaxLeaf = 4;

struct PrimBox { float lo[3], hi[3]; uint32_t type, index; };

struct BuiltBVH {
    std::vector<HostNode> nodes;              // node 0 is the root (always an interior node)
    std::vector<uint32_t> order[3];           // per type: leaf-ordered list of original primitive indices
    double build_ms = 0.0;
    uint32_t max_depth = 0;
};

// Binned-SAH build over all primitives (multi-threaded).  `pad` widens every box (see lgb_api.cu).
int build_sah(const PrimBox* prims, size_t n, float pad, int threads, BuiltBVH& out);

// rank[o * prim_count + id]: position of canonical primitive `id` in the reference's traversal order
// for direction octant o (bit a set <=> dir_is_neg[a], bvh.rs:463).  Returns false on a malformed tree.
bool build_rank_tables(const lgb_scene_desc* d, uint32_t prim_count, int threads, uint32_t* rank);   // rank: 8 * prim_count words

}  // namespace lgb
