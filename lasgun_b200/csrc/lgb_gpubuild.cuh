// Device-side binned-SAH BVH build (lgb_gpubuild.cu) and the kernels around it (lgb_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lgb_build.hpp"

namespace lgb {

// One primitive during the build: padded f32 box and what it is (same 32 bytes as the host builder's item).
struct GItem { float lo[3]; uint32_t type; float hi[3]; uint32_t index; };

struct GpuBuildInfo { uint32_t n_nodes, max_depth, levels; };

size_t gpu_build_temp_bytes(uint32_t n);

// items_in: n padded item boxes (device, left untouched).  nodes_out: room for n HostNodes (node 0 = root).
// final_items: n items in tree order; typepos_out[t][i] = number of items of type t before position i in that order
// (valid until `temp` is released).  Leaf words of nodes_out index the per-type arrays in that order.
cudaError_t gpu_build_sah(const GItem* items_in, uint32_t n, HostNode* nodes_out, GItem* final_items, uint32_t** typepos_out, void* temp, size_t temp_bytes,
                          cudaStream_t stream, GpuBuildInfo* info);

cudaError_t gpu_identity_positions(const GItem* items, uint32_t n, uint32_t** typepos_out, void* temp, size_t temp_bytes, cudaStream_t st);

// The caller's arrays as uploaded (device pointers, caller layout) and the leaf-ordered arrays the kernels read.
struct RawScene {
    const lgb_sphere* spheres; const uint32_t* sphere_material; const uint32_t* sphere_id; uint32_t n_spheres;
    const lgb_cuboid* cuboids; const uint32_t* cuboid_material; const uint32_t* cuboid_id; uint32_t n_cuboids;
    const lgb_triangle* triangles; const uint32_t* triangle_material; const uint32_t* triangle_id;
    const lgb_tri_normals* tri_normals; const uint8_t* tri_has_normals; uint32_t n_triangles;
};
struct LeafArrays {
    float4* sph32; double* sph64; uint32_t* sph_mat; uint32_t* sph_id;
    float4* cub32; double* cub64; uint32_t* cub_mat; uint32_t* cub_id;
    float4* tri; float* tri_nrm;
};
// items[i] for primitive i in the order spheres, cuboids, triangles: box rounded outward to f32 and widened by `pad`
cudaError_t launch_make_items(const RawScene& raw, float pad, GItem* items, cudaStream_t stream);
// final_items (tree order) + per-type positions -> leaf-ordered primitive arrays (same records as the host path writes)
cudaError_t launch_convert(const RawScene& raw, const LeafArrays& out, const GItem* final_items, uint32_t n, uint32_t* const typepos[3], double cuboid_pad,
                           cudaStream_t stream);

}  // namespace lgb
