// Light grids: an exact shadow-ray acceleration structure for point lights (light/point.rs:42-54).
//
// Every shadow ray towards light L lies on a line through L, so the primitives that can block it are those whose
// DIRECTION FOOTPRINT seen from L contains the ray's direction.  A light grid is a cube map around L (6 faces x res^2 cells);
// each cell lists the primitives whose conservative footprint touches it, nearest first.  A shadow ray looks up ONE cell
// (k_gshadow, lgb_kernels.cu) and runs the usual f32 filter + exact f64 test on its few entries up to the ray's own length:
// no BVH traversal, no per-ray stack, no queue compaction.  The candidate set is complete (a primitive hit at a point X is
// listed in the cell of X's direction from L, which is the ray's), the tests are the reference's, so the occlusion bits are
// those of the traversal, bit for bit.  Primitives whose footprint covers more than kGridLargeCells cells (a ground slab)
// go to a short per-light list every ray tests.
//
// Built on the device at scene creation from the leaf-ordered primitive arrays: count -> exclusive scan (cub::DeviceScan,
// off the frame's path) -> fill -> per-cell sort by distance.  Single-space scenes only (no transformed aggregates).
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "lgb_grid.cuh"

namespace lgb {
namespace {

__device__ __forceinline__ bool prim_box(const DevScene& S, uint32_t i, double lo[3], double hi[3], uint32_t& ref) {
    const double pad = (double)S.err_abs;
    if (i < S.n_sph) {
        const float4 s = S.sph32[i];
        const double r = (double)s.w + pad;            // (s.w is the radius rounded up)
        lo[0] = s.x - r; lo[1] = s.y - r; lo[2] = s.z - r; hi[0] = s.x + r; hi[1] = s.y + r; hi[2] = s.z + r;
        // the f32 centre is the f64 centre rounded: one more ulp of the coordinate bound covers it
        ref = LGB_PRIM_REF(LGB_PRIM_SPHERE, i);
    } else if (i < S.n_sph + S.n_cub) {
        const uint32_t k = i - S.n_sph;
        const float4 a = S.cub32[2 * k], b = S.cub32[2 * k + 1];      // padded already
        lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
        ref = LGB_PRIM_REF(LGB_PRIM_CUBOID, k);
    } else {
        const uint32_t k = i - S.n_sph - S.n_cub;
        const float4 q0 = S.tri[3 * (size_t)k], q1 = S.tri[3 * (size_t)k + 1], q2 = S.tri[3 * (size_t)k + 2];
        lo[0] = fmin(q0.x, fmin(q1.x, q2.x)); hi[0] = fmax(q0.x, fmax(q1.x, q2.x));
        lo[1] = fmin(q0.y, fmin(q1.y, q2.y)); hi[1] = fmax(q0.y, fmax(q1.y, q2.y));
        lo[2] = fmin(q0.z, fmin(q1.z, q2.z)); hi[2] = fmax(q0.z, fmax(q1.z, q2.z));
        ref = LGB_PRIM_REF(LGB_PRIM_TRIANGLE, k);
    }
    for (int k = 0; k < 3; k++) { lo[k] -= pad; hi[k] += pad; }
    return true;
}

// Direction footprint of the box [lo, hi] - o on cube-map face `face` (major axis a = face / 2, negative side iff face & 1):
// u = x_b / w, v = x_c / w over the part of the box with w = +-x_a > 0, b = (a + 1) % 3, c = (a + 2) % 3, clipped to the face
// [-1, 1]^2 and widened by 1e-9.  false: not seen on this face.  Conservative.
__device__ __forceinline__ bool face_uv(const double lo[3], const double hi[3], const double o[3], int face, double uv[4]) {
    const int a = face >> 1, b = (a + 1) % 3, c = (a + 2) % 3;
    double wl, wh;
    if (face & 1) { wl = -(hi[a] - o[a]); wh = -(lo[a] - o[a]); } else { wl = lo[a] - o[a]; wh = hi[a] - o[a]; }
    if (!(wh > 0.0)) return false;
    const double inv_l = wl > 0.0 ? 1.0 / wl : CUDART_INF, inv_h = 1.0 / wh;
    const double lim[2][2] = {{lo[b] - o[b], hi[b] - o[b]}, {lo[c] - o[c], hi[c] - o[c]}};
    for (int k = 0; k < 2; k++) {
        const double xl = lim[k][0], xh = lim[k][1];
        double umin = xl < 0.0 ? xl * inv_l : xl * inv_h, umax = xh > 0.0 ? xh * inv_l : xh * inv_h;
        if (xl < 0.0 && !(wl > 0.0)) umin = -CUDART_INF;
        if (xh > 0.0 && !(wl > 0.0)) umax = CUDART_INF;
        if (umin > 1.0 || umax < -1.0) return false;          // seen on a neighbouring face only
        uv[2 * k] = fmax(umin - 1e-9, -1.0); uv[2 * k + 1] = fmin(umax + 1e-9, 1.0);
    }
    return true;
}
// Cells of that footprint under the face's mapping {u0, su, v0, sv}: floor((u - u0) su), clamped to [0, res) -- the expression
// grid_blocked (lgb_kernels.cu) evaluates for a ray's direction; clamping keeps the listing complete for directions outside the mapped part.
__device__ __forceinline__ void face_cells(const double uv[4], const double* map, uint32_t res, int rect[4]) {
    for (int k = 0; k < 2; k++) {
        int c0 = (int)fmin(fmax(floor((uv[2 * k] - map[2 * k]) * map[2 * k + 1] - 1e-6), -1.0), (double)res);
        int c1 = (int)fmin(fmax(floor((uv[2 * k + 1] - map[2 * k]) * map[2 * k + 1] + 1e-6), -1.0), (double)res);
        rect[2 * k] = min(max(c0, 0), (int)res - 1); rect[2 * k + 1] = min(max(c1, 0), (int)res - 1);
    }
}

__device__ __forceinline__ float box_dmin(const double lo[3], const double hi[3], const double o[3]) {
    double d2 = 0.0;
    for (int k = 0; k < 3; k++) { const double d = fmax(fmax(lo[k] - o[k], o[k] - hi[k]), 0.0); d2 += d * d; }
    return fmaxf(__double2float_rd(sqrt(d2) * (1.0 - 1e-6)), 0.0f);
}

constexpr int kBoundWords = 7;       // per face: min u, max u, min v, max v (dkeys), sum of u extents, sum of v extents, count (doubles)
// order-preserving integer image of a double, for atomicMin / atomicMax
__device__ __forceinline__ unsigned long long dkey(double v) { const unsigned long long b = (unsigned long long)__double_as_longlong(v); return (b >> 63) ? ~b : (b | 0x8000000000000000ull); }
__host__ __device__ inline double dkey_inv(unsigned long long k) { const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k; double d; memcpy(&d, &b, 8); return d; }

// pass A: the part of every face that the scene's SMALL primitives cover (footprints of at most `small_frac` of the face per axis);
// bounds[face] = {min u, max u, min v, max v} as dkeys.  The cells are then laid over that part only: a light far from a compact
// scene spends its res^2 cells on the few degrees the scene subtends, not on the whole face.
__global__ void __launch_bounds__(256) k_grid_bounds(DevScene S, uint32_t light, double small_span, unsigned long long* bounds) {
    // per block in shared memory first: 42 words that every primitive of the scene would otherwise hit with global atomics (measured:
    // 2.9 - 5.5 ms per light for 0.6 - 1 M primitives that way)
    __shared__ unsigned long long sb[6 * kBoundWords];
    for (int k = threadIdx.x; k < 6 * kBoundWords; k += blockDim.x) { const int w = k % kBoundWords; sb[k] = (w == 0 || w == 2) ? ~0ull : 0ull; }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.n_sph + S.n_cub + S.n_tri) {
        const double* L = S.lights + 9 * (size_t)light;
        const double o[3] = {L[0], L[1], L[2]};
        double lo[3], hi[3]; uint32_t ref;
        prim_box(S, i, lo, hi, ref);
        for (int f = 0; f < 6; f++) {
            double uv[4];
            if (!face_uv(lo, hi, o, f, uv)) continue;
            if (uv[1] - uv[0] > small_span || uv[3] - uv[2] > small_span) continue;
            unsigned long long* b = sb + kBoundWords * f;
            atomicMin(b + 0, dkey(uv[0])); atomicMax(b + 1, dkey(uv[1])); atomicMin(b + 2, dkey(uv[2])); atomicMax(b + 3, dkey(uv[3]));
            double* sum = reinterpret_cast<double*>(b + 4);              // {sum of u extents, sum of v extents, count}
            atomicAdd(sum + 0, uv[1] - uv[0]); atomicAdd(sum + 1, uv[3] - uv[2]); atomicAdd(sum + 2, 1.0);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 6 * kBoundWords; k += blockDim.x) {
        const int w = k % kBoundWords;
        if (w == 0 || w == 2) { if (sb[k] != ~0ull) atomicMin(bounds + k, sb[k]); }
        else if (w == 1 || w == 3) { if (sb[k] != 0ull) atomicMax(bounds + k, sb[k]); }
        else { const double v = __longlong_as_double((long long)sb[k]); if (v != 0.0) atomicAdd(reinterpret_cast<double*>(bounds + k), v); }
    }
}
// A face's cells: laid over the covered part, but never finer than the average small footprint -- cells smaller than the primitives
// only multiply the entries (every primitive is listed in each cell it touches) without shortening the lists a ray walks.
__global__ void k_grid_map(const unsigned long long* bounds, uint32_t res, double cell_factor, DevGrid* grid) {
    const int f = threadIdx.x;
    if (f >= 6) return;
    const unsigned long long* b = bounds + kBoundWords * f;
    double u0 = dkey_inv(b[0]), u1 = dkey_inv(b[1]), v0 = dkey_inv(b[2]), v1 = dkey_inv(b[3]);
    const double* sum = reinterpret_cast<const double*>(b + 4);
    if (!(u0 <= u1) || !(v0 <= v1) || !(sum[2] > 0.0) || cell_factor < 0.0) { u0 = v0 = -1.0; u1 = v1 = 1.0; }          // nothing small on this face (cell_factor < 0: the whole face, experiments)
    const double au = sum[2] > 0.0 ? sum[0] / sum[2] : 0.0, av = sum[2] > 0.0 ? sum[1] / sum[2] : 0.0;
    const double cf = fmax(cell_factor, 0.0);
    const double cu = fmax(fmax(u1 - u0, 1e-6) / (double)res, cf * au), cv = fmax(fmax(v1 - v0, 1e-6) / (double)res, cf * av);
    grid->map[f][0] = u0; grid->map[f][1] = 1.0 / cu; grid->map[f][2] = v0; grid->map[f][3] = 1.0 / cv;
}

// pass 0: counts per cell (and the large list); pass 1: the entries
template <int PASS>
__global__ void __launch_bounds__(256) k_grid_pass(DevScene S, uint32_t light, const DevGrid* grid, uint32_t res, uint32_t large_cells, uint32_t large_cap,
                                                   uint32_t* counts, const uint32_t* starts, uint2* entries, uint2* large, uint32_t* n_large) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n_sph + S.n_cub + S.n_tri) return;
    const double* L = S.lights + 9 * (size_t)light;
    const double o[3] = {L[0], L[1], L[2]};
    double lo[3], hi[3]; uint32_t ref;
    prim_box(S, i, lo, hi, ref);
    int rect[6][4]; bool on[6]; uint64_t cells = 0;
    for (int f = 0; f < 6; f++) {
        double uv[4];
        on[f] = face_uv(lo, hi, o, f, uv);
        if (on[f]) { face_cells(uv, grid->map[f], res, rect[f]); cells += (uint64_t)(rect[f][1] - rect[f][0] + 1) * (uint64_t)(rect[f][3] - rect[f][2] + 1); }
    }
    if (cells == 0) return;
    const uint2 rec = make_uint2(ref, __float_as_uint(box_dmin(lo, hi, o)));
    if (cells > large_cells) {
        if (PASS == 0) { const uint32_t k = atomicAdd(n_large, 1u); if (k < large_cap) large[k] = rec; }
        return;
    }
    for (int f = 0; f < 6; f++) {
        if (!on[f]) continue;
        for (int v = rect[f][2]; v <= rect[f][3]; v++)
            for (int u = rect[f][0]; u <= rect[f][1]; u++) {
                const size_t cell = ((size_t)f * res + (size_t)v) * res + (size_t)u;
                if (PASS == 0) atomicAdd(&counts[cell], 1u);
                else entries[starts[cell] + (atomicSub(&counts[cell], 1u) - 1u)] = rec;
            }
    }
}

__device__ __forceinline__ bool entry_less(uint2 x, uint2 y) {
    const float kx = __uint_as_float(x.y), ky = __uint_as_float(y.y);
    return kx < ky || (kx == ky && x.x < y.x);
}
__global__ void k_grid_sort_large(uint2* large, uint32_t n) {
    if (blockIdx.x || threadIdx.x) return;
    for (uint32_t i = 1; i < n; i++) {
        const uint2 x = large[i]; uint32_t j = i;
        while (j > 0 && entry_less(x, large[j - 1])) { large[j] = large[j - 1]; j--; }
        large[j] = x;
    }
}

// ---- camera grid: the same idea for the primary rays of a perspective camera (camera.rs:113-146: every ray leaves `origin`).
// Cells are tiles of 2^shift x 2^shift pixels of the film; a primitive is listed in the tiles that hold a pixel one of whose samples
// can see its box.  In the camera's own (not necessarily orthonormal) frame a point X = origin + w (view + a aux + b up) is seen at
// image-plane coordinates (a, b); sample (i, j) of pixel (x, y) looks along a = sox(x) + (i + 0.5) sep, b = soy(y) + (j + 0.5) sep.
template <int PASS>
__global__ void __launch_bounds__(256) k_cam_pass(DevScene S, CamGridParams P, uint32_t* counts, const uint32_t* starts, uint2* entries, uint2* large, uint32_t* n_large) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.n_sph + S.n_cub + S.n_tri) return;
    double lo[3], hi[3]; uint32_t ref;
    prim_box(S, i, lo, hi, ref);
    double amin = CUDART_INF, amax = -CUDART_INF, bmin = CUDART_INF, bmax = -CUDART_INF;
    bool behind = false, front = false;
    for (int k = 0; k < 8; k++) {
        const double x = ((k & 1) ? hi[0] : lo[0]) - P.origin[0], y = ((k & 2) ? hi[1] : lo[1]) - P.origin[1], z = ((k & 4) ? hi[2] : lo[2]) - P.origin[2];
        const double qa = P.minv[0] * x + P.minv[1] * y + P.minv[2] * z, qb = P.minv[3] * x + P.minv[4] * y + P.minv[5] * z, w = P.minv[6] * x + P.minv[7] * y + P.minv[8] * z;
        if (w > P.w_eps) { front = true; const double a = qa / w, b = qb / w; amin = fmin(amin, a); amax = fmax(amax, a); bmin = fmin(bmin, b); bmax = fmax(bmax, b); }
        else behind = true;
    }
    if (!front) return;                                   // entirely behind the eye: no camera ray reaches it
    const uint2 rec = make_uint2(ref, __float_as_uint(box_dmin(lo, hi, P.origin)));
    bool is_large = behind;                               // the box crosses the eye's plane: its projection is unbounded
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (!is_large) {
        // continuous pixel coordinates of the image-plane point: fa(a) = (a / ipw + 0.5) w  (sox(fa) = a), fb(b) = (0.5 - b / iph) h - 1
        const double xa = (amin / P.ipw + 0.5) * P.w, xb = (amax / P.ipw + 0.5) * P.w;
        const double ya = (0.5 - bmin / P.iph) * P.h - 1.0, yb = (0.5 - bmax / P.iph) * P.h - 1.0;
        // a pixel's samples reach delta0 .. delta1 pixels beyond its corner: x + delta in [min(xa, xb), max(xa, xb)], y - delta likewise
        const double xl = fmin(xa, xb) - P.delta1 - 1e-3, xh = fmax(xa, xb) - P.delta0 + 1e-3;
        const double yl = fmin(ya, yb) + P.delta0 - 1e-3, yh = fmax(ya, yb) + P.delta1 + 1e-3;
        if (xh < 0.0 || yh < 0.0 || xl > P.w - 1.0 || yl > P.h - 1.0) return;      // outside the film
        x0 = (int)fmax(ceil(xl), 0.0); x1 = (int)fmin(floor(xh), P.w - 1.0); y0 = (int)fmax(ceil(yl), 0.0); y1 = (int)fmin(floor(yh), P.h - 1.0);
        if (x0 > x1 || y0 > y1) return;
        x0 >>= P.shift; x1 >>= P.shift; y0 >>= P.shift; y1 >>= P.shift;
        is_large = (uint64_t)(x1 - x0 + 1) * (uint64_t)(y1 - y0 + 1) > P.large_cells;
    }
    if (is_large) {
        if (PASS == 0) { const uint32_t k = atomicAdd(n_large, 1u); if (k < P.large_cap) large[k] = rec; }
        return;
    }
    for (int y = y0; y <= y1; y++)
        for (int x = x0; x <= x1; x++) {
            const size_t cell = (size_t)y * P.nx + (size_t)x;
            if (PASS == 0) atomicAdd(&counts[cell], 1u);
            else entries[starts[cell] + (atomicSub(&counts[cell], 1u) - 1u)] = rec;
        }
}

}  // namespace

// Every cell's entries nearest first: the record (ref, dmin bits) read as one little-endian u64 is (dmin << 32) | ref, and the bits of
// a non-negative float order like the float, so a cell is sorted as plain u64 keys (all distinct: a primitive is listed once per cell).
// k_sort_cells: one warp per 32 consecutive cells; a list of up to 32 entries is a bitonic network over the lanes (shuffles only), empty
// cells and single entries cost a ballot; longer cells are listed.  k_sort_medium: one warp per listed cell of up to kSortWarp entries,
// bitonic sort in the warp's slice of shared memory (a warp of their own each: a view along a row of primitives puts hundreds of entries
// into each of a run of adjacent cells).  Cells above kSortWarp entries go to cub::DeviceSegmentedSort (off the frame's path, like the
// scans) as a compacted list of ranges.  Both lists are sized by their bounds -- at most total / 33 medium and total / (kSortWarp + 1)
// long cells -- and the launches are made for the bounds (unused ranges are empty), so the host need not read anything back.
// Measured before: one THREAD per short cell (insertion sort in local memory) + the segmented sort for everything above 32 entries
// took 1.8 ms per light and 4.7 ms for the camera grid on `spheres1m` (12 M entries, 45 per tile); now 0.3 / 0.5 ms.
constexpr uint32_t kSortWarp = 1024;
constexpr int kSortWarps = 4;            // warps per block of k_sort_medium: 4 x 8 KB of shared memory
static uint32_t long_bound(uint32_t total) { return total / (kSortWarp + 1) + 1; }
static uint32_t medium_bound(uint32_t total) { return total / 33 + 1; }
static size_t long_list_bytes(uint32_t total) { return ((size_t)(2 * long_bound(total) + 2 * medium_bound(total) + 2) * 4 + 255) & ~(size_t)255; }
size_t grid_sort_bytes(uint32_t total, size_t) {
    size_t b = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)total, (int)long_bound(total), (const uint32_t*)nullptr, (const uint32_t*)nullptr);
    return b + long_list_bytes(total);
}
__global__ void __launch_bounds__(128) k_sort_cells(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, const uint32_t* __restrict__ starts,
                                                    uint32_t n_cells, uint32_t* long_begin, uint32_t* long_end, uint32_t* n_long, uint2* medium, uint32_t* n_medium) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;         // a warp's lanes hold 32 consecutive cells
    uint32_t b = 0, n = 0;
    if (cell < n_cells) { b = starts[cell]; n = starts[cell + 1] - b; }
    if (n == 1) out[b] = in[b];
    else if (n > kSortWarp) { const uint32_t k = atomicAdd(n_long, 1u); long_begin[k] = b; long_end[k] = b + n; }
    else if (n > 32u) medium[atomicAdd(n_medium, 1u)] = make_uint2(b, n);
    unsigned todo = __ballot_sync(0xFFFFFFFFu, n >= 2u && n <= 32u);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t cb = __shfl_sync(0xFFFFFFFFu, b, src), cn = __shfl_sync(0xFFFFFFFFu, n, src);
        unsigned long long v = lane < cn ? in[cb + lane] : ~0ull;          // one entry per lane, padded with the largest key
        for (unsigned k = 2; k <= 32u; k <<= 1)
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, j);
                const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                v = keep_min ? (o < v ? o : v) : (o > v ? o : v);
            }
        if (lane < cn) out[cb + lane] = v;
    }
}
__global__ void __launch_bounds__(32 * kSortWarps) k_sort_medium(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, const uint2* __restrict__ medium,
                                                                  const uint32_t* __restrict__ n_medium) {
    __shared__ unsigned long long sm[kSortWarps][kSortWarp];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t m = blockIdx.x * kSortWarps + warp;
    if (m >= *n_medium) return;
    unsigned long long* buf = sm[warp];
    const uint32_t cb = medium[m].x, cn = medium[m].y;
    uint32_t P = 64; while (P < cn) P <<= 1;
    for (uint32_t i = lane; i < P; i += 32) buf[i] = i < cn ? in[cb + i] : ~0ull;
    __syncwarp();
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < P / 2; t += 32) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;      // the pair (i, i + j), bit j of i clear
                const unsigned long long x = buf[i], y = buf[l];
                if ((x > y) == ((i & k) == 0)) { buf[i] = y; buf[l] = x; }
            }
            __syncwarp();
        }
    for (uint32_t i = lane; i < cn; i += 32) out[cb + i] = buf[i];
}
static cudaError_t sort_cells(const uint2* in, uint2* out, uint32_t total, const uint32_t* starts, size_t n_cells, void* tmp, size_t tmp_bytes, cudaStream_t st) {
    if (total == 0) return cudaSuccess;
    const uint32_t bound = long_bound(total), mbound = medium_bound(total);
    const size_t list_bytes = long_list_bytes(total);
    uint32_t* long_begin = (uint32_t*)tmp; uint32_t* long_end = long_begin + bound; uint32_t* n_long = long_end + bound; uint32_t* n_medium = n_long + 1;
    uint2* medium = (uint2*)(n_medium + 1);                                   // (8-byte aligned: 2 * bound + 2 words precede it)
    if (cudaError_t e = cudaMemsetAsync(tmp, 0, list_bytes, st)) return e;
    auto keys_in = reinterpret_cast<const unsigned long long*>(in); auto keys_out = reinterpret_cast<unsigned long long*>(out);
    k_sort_cells<<<(unsigned)((n_cells + 127) / 128), 128, 0, st>>>(keys_in, keys_out, starts, (uint32_t)n_cells, long_begin, long_end, n_long, medium, n_medium);
    k_sort_medium<<<(mbound + kSortWarps - 1) / kSortWarps, 32 * kSortWarps, 0, st>>>(keys_in, keys_out, medium, n_medium);
    size_t cub_bytes = tmp_bytes - list_bytes;
    return cub::DeviceSegmentedSort::SortKeys((char*)tmp + list_bytes, cub_bytes, keys_in, keys_out, (int)total, (int)bound, long_begin, long_end, st);
}
cudaError_t camgrid_count(const DevScene& S, const CamGridParams& P, uint32_t* counts, uint32_t* starts, void* scan_tmp, size_t scan_bytes,
                          uint2* large, uint32_t* n_large_dev, cudaStream_t st, uint32_t* total_out, uint32_t* n_large_out) {
    const size_t nc = (size_t)P.nx * P.ny;
    const uint32_t n = S.n_sph + S.n_cub + S.n_tri;
    cudaError_t e;
    if ((e = cudaMemsetAsync(counts, 0, (nc + 1) * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(n_large_dev, 0, 4, st)) != cudaSuccess) return e;
    k_cam_pass<0><<<(n + 255) / 256, 256, 0, st>>>(S, P, counts, nullptr, nullptr, large, n_large_dev);
    if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, starts, (int)(nc + 1), st)) != cudaSuccess) return e;
    uint32_t h[2] = {0, 0};
    if ((e = cudaMemcpyAsync(&h[0], starts + nc, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(&h[1], n_large_dev, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    *total_out = h[0]; *n_large_out = h[1];
    return cudaGetLastError();
}
cudaError_t camgrid_fill(const DevScene& S, const CamGridParams& P, uint32_t* counts, const uint32_t* starts, uint2* entries_tmp, uint2* entries, uint32_t total,
                         void* sort_tmp, size_t sort_bytes, uint2* large, uint32_t n_large, cudaStream_t st) {
    const size_t nc = (size_t)P.nx * P.ny;
    const uint32_t n = S.n_sph + S.n_cub + S.n_tri;
    k_cam_pass<1><<<(n + 255) / 256, 256, 0, st>>>(S, P, counts, starts, entries_tmp, nullptr, nullptr);
    if (cudaError_t e = sort_cells(entries_tmp, entries, total, starts, nc, sort_tmp, sort_bytes, st)) return e;
    if (n_large > 1) k_grid_sort_large<<<1, 32, 0, st>>>(large, n_large);
    return cudaGetLastError();
}
size_t scan_bytes_for(size_t n_cells) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n_cells + 1));
    return b;
}

size_t grid_cells(uint32_t res) { return (size_t)6 * res * res; }

// Light grid of light `light`, first half: the face mappings (into *grid_dev), the counts, their scan into `starts`, the large list.
// Asynchronous: the totals land in totals_dev[0] (entries) and totals_dev[1] (large primitives).
cudaError_t grid_count(const DevScene& S, uint32_t light, uint32_t res, DevGrid* grid_dev, unsigned long long* bounds, uint32_t* counts, uint32_t* starts,
                       void* scan_tmp, size_t scan_bytes, uint2* large, uint32_t* totals_dev, cudaStream_t st) {
    const size_t nc = grid_cells(res);
    const uint32_t n = S.n_sph + S.n_cub + S.n_tri;
    cudaError_t e;
    unsigned long long init[6 * kBoundWords] = {};
    for (int f = 0; f < 6; f++) { init[kBoundWords * f] = init[kBoundWords * f + 2] = ~0ull; }      // min keys: all ones; max keys and the sums: zero
    if ((e = cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;      // (pageable source: copied before the call returns)
    if ((e = cudaMemsetAsync(counts, 0, (nc + 1) * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(totals_dev, 0, 8, st)) != cudaSuccess) return e;
    const bool timing = getenv("LGB_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!timing) return; cudaStreamSynchronize(st); fprintf(stderr, "[light grids]     light %u %-14s %.2f ms\n", light, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); t0 = std::chrono::steady_clock::now(); };
    lap("memsets");
    k_grid_bounds<<<(n + 255) / 256, 256, 0, st>>>(S, light, kGridSmallSpan, bounds);
    lap("k_grid_bounds");
    double cell_factor = 1.0;                        // cells no finer than this times the average small footprint (env LGB_GRID_CELL: experiments)
    if (const char* ev = getenv("LGB_GRID_CELL")) cell_factor = atof(ev);
    k_grid_map<<<1, 32, 0, st>>>(bounds, res, cell_factor, grid_dev);
    k_grid_pass<0><<<(n + 255) / 256, 256, 0, st>>>(S, light, grid_dev, res, kGridLargeCells, kGridLargeCap, counts, nullptr, nullptr, large, totals_dev + 1);
    lap("count pass");
    if ((e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, starts, (int)(nc + 1), st)) != cudaSuccess) return e;
    lap("scan");
    if ((e = cudaMemcpyAsync(totals_dev, starts + nc, 4, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    return cudaGetLastError();
}
cudaError_t grid_fill(const DevScene& S, uint32_t light, uint32_t res, const DevGrid* grid_dev, uint32_t* counts, const uint32_t* starts, uint2* entries_tmp, uint2* entries,
                      uint32_t total, void* sort_tmp, size_t sort_bytes, uint2* large, uint32_t n_large, cudaStream_t st) {
    const size_t nc = grid_cells(res);
    const uint32_t n = S.n_sph + S.n_cub + S.n_tri;
    cudaError_t e;
    // the counts of THIS light again (the buffer is shared by the lights): pass 1 consumes them as cursors
    if ((e = cudaMemsetAsync(counts, 0, (nc + 1) * 4, st)) != cudaSuccess) return e;
    const bool timing = getenv("LGB_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) { if (!timing) return; cudaStreamSynchronize(st); fprintf(stderr, "[light grids]     light %u %-14s %.2f ms\n", light, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); t0 = std::chrono::steady_clock::now(); };
    k_grid_pass<0><<<(n + 255) / 256, 256, 0, st>>>(S, light, grid_dev, res, kGridLargeCells, 0u, counts, nullptr, nullptr, nullptr, counts + nc);
    lap("recount");
    k_grid_pass<1><<<(n + 255) / 256, 256, 0, st>>>(S, light, grid_dev, res, kGridLargeCells, kGridLargeCap, counts, starts, entries_tmp, nullptr, nullptr);
    lap("fill pass");
    if ((e = sort_cells(entries_tmp, entries, total, starts, nc, sort_tmp, sort_bytes, st)) != cudaSuccess) return e;
    lap("sort");
    if (n_large > 1) k_grid_sort_large<<<1, 32, 0, st>>>(large, n_large);
    return cudaGetLastError();
}
size_t grid_scan_bytes(uint32_t res) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(grid_cells(res) + 1));
    return b;
}

}  // namespace lgb
