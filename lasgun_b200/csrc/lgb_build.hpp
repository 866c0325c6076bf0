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
#include <string>
#include <vector>

#include "../../include/lasgun_b200.h"
#include "lgb_parallel.hpp"

namespace lgb {

struct HostNode { float v[12]; uint32_t c0, c1, pad0, pad1; };   // 64 B: child0 {lo, hi}, child1 {lo, hi}, child words

// child word: interior = node index; leaf = kLeafBit | type << 29 | (count - 1) << 24 | first
#ifndef LGB_LEAF_CONSTANTS
#define LGB_LEAF_CONSTANTS
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kLeafFirstMask = 0x00FFFFFFu;
#endif
#ifndef LGB_MAX_LEAF
#define LGB_MAX_LEAF 4
#endif
constexpr int kMaxLeaf = LGB_MAX_LEAF;

struct PrimBox { float lo[3], hi[3]; uint32_t type, index; };    // type 3 (LGB_PRIM_INSTANCE): index = child space

struct BuiltBVH {
    std::vector<HostNode> nodes;              // node 0 is the root (always an interior node)
    std::vector<uint32_t> order[4];           // per type: leaf-ordered list of original primitive indices / child spaces
    double build_ms = 0.0;
    uint32_t max_depth = 0;
};

// Binned-SAH build over all primitives (multi-threaded).  `pad` widens every box (see lgb_api.cu).
int build_sah(const PrimBox* prims, size_t n, float pad, int threads, BuiltBVH& out);

// A SPACE is a coordinate system rays are traversed in: space 0 is the root aggregate's, and every nested BVH
// level whose aggregate carries a transform or swap_backface opens a new one (bvh.rs:462-464, :508-518).  Levels
// with the identity transform (meshes, plain groups) are traversed inline in their parent's space.
constexpr uint32_t kNoSpace = 0xFFFFFFFFu;
struct HostSpace {
    uint32_t parent = kNoSpace, depth = 0;
    uint32_t ref_root = 0;                    // node of the caller's reference tree where the defining level starts
    uint32_t root_node = 0;                   // root of the space's device BVH in BuiltScene::nodes
    uint32_t identity = 1, swap_backface = 0;
    double m[16], minv[16];
    double max_abs = 0.0;                     // bound on |coordinate| of geometry and ray origins in this space
    float err_abs = 0.0f;                     // absolute error bound of an f32 ray against this space (max_abs * 2^-20)
    uint32_t stack_need = 0;
};
struct BuiltScene {
    std::vector<HostNode> nodes;              // all spaces, child words absolute
    raw_vector<uint32_t> order[4];            // leaf order per type over all spaces; [3] = child space ids
    raw_vector<uint32_t> prim_space[3];       // by ORIGINAL primitive index: the space it lives in (empty when there is one space)
    std::vector<uint32_t> inst_space;         // by instance index: the space it opens, or kNoSpace when traversed inline
    std::vector<HostSpace> spaces;
    double build_ms = 0.0;
    bool instanced() const { return spaces.size() > 1 || !spaces[0].identity || spaces[0].swap_backface; }
};
// world_lo / world_hi: box containing the root level and every ray origin (camera), in world coordinates.
// Returns 0, or a negative code with `err` set.
int build_scene(const lgb_scene_desc* d, const double world_lo[3], const double world_hi[3], BuiltScene& out, std::string& err);

// rank[o * (prim_count + n_spaces) + item]: position of the item in the reference's test order inside ITS space for
// direction octant o of the ray in that space (bit a set <=> dir_is_neg[a], bvh.rs:463); item = canonical primitive id,
// or prim_count + c for the nested level that opens space c.  Returns false on a malformed tree.
bool build_rank_tables(const lgb_scene_desc* d, uint32_t prim_count, const BuiltScene& bs, uint32_t* rank);

}  // namespace lgb
