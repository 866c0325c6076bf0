// Light grids (lgb_grid.cu): per point light, a cube map of primitive lists that replaces the BVH for its shadow rays.
#pragma once
#include <math_constants.h>

#include "lgb_types.cuh"

namespace lgb {

constexpr uint32_t kGridLargeCells = 4096;   // a footprint of more cells than this goes to the light's `large` list
constexpr uint32_t kGridLargeCap = 128;      // more large primitives than this: no grid for that light (BVH shadow rays instead)

// Camera grid (primary rays of a perspective camera): host-computed parameters of the binning pass.
struct CamGridParams {
    double origin[3];
    double minv[9];              // rows of [aux | up | view]^-1: (a w, b w, w) = minv (X - origin)
    double ipw, iph, w, h;       // image plane extents (camera.rs:116-117), film size
    double delta0, delta1;       // a pixel's samples lie delta0 .. delta1 pixels beyond its corner (camera.rs:135-143)
    double w_eps;                // corners with w <= w_eps count as behind the eye
    uint32_t shift, nx, ny;      // tiles of 2^shift pixels, nx x ny of them
    uint32_t large_cells, large_cap;
};
cudaError_t camgrid_count(const DevScene& S, const CamGridParams& P, uint32_t* counts, uint32_t* starts, void* scan_tmp, size_t scan_bytes,
                          uint2* large, uint32_t* n_large_dev, cudaStream_t st, uint32_t* total_out, uint32_t* n_large_out);
cudaError_t camgrid_fill(const DevScene& S, const CamGridParams& P, uint32_t* counts, const uint32_t* starts, uint2* entries_tmp, uint2* entries, uint32_t total,
                         void* sort_tmp, size_t sort_bytes, uint2* large, uint32_t n_large, cudaStream_t st);
size_t scan_bytes_for(size_t n_cells);

size_t grid_cells(uint32_t res);
size_t grid_scan_bytes(uint32_t res);
constexpr double kGridSmallSpan = 0.25;      // footprints up to this much of a face's [-1, 1] per axis shape the face's mapping
cudaError_t grid_count(const DevScene& S, uint32_t light, uint32_t res, DevGrid* grid_dev, unsigned long long* bounds, uint32_t* counts, uint32_t* starts,
                       void* scan_tmp, size_t scan_bytes, uint2* large, uint32_t* totals_dev, cudaStream_t st);
size_t grid_sort_bytes(uint32_t total, size_t n_cells);
cudaError_t grid_fill(const DevScene& S, uint32_t light, uint32_t res, const DevGrid* grid_dev, uint32_t* counts, const uint32_t* starts, uint2* entries_tmp, uint2* entries,
                      uint32_t total, void* sort_tmp, size_t sort_bytes, uint2* large, uint32_t n_large, cudaStream_t st);

}  // namespace lgb
