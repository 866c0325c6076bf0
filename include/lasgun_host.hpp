// lasgun_host.hpp — C++ mirror of lasgun's public Rust API above the lgb_* C ABI.
//
// The reference is compiled Rust and no Rust toolchain exists in this image, so the host side
// (scene builder, HLBVH build, flattening, capture) is C++ with the reference's names, argument
// meaning and error behaviour (a Rust `panic!`/`unwrap` becomes a lasgun::Error exception):
//   Scene, Aggregate, Material, ObjRef     src/scene.rs, src/scene/node.rs, src/material/mod.rs
//   Camera::look_at / set_supersampling     src/camera.rs:85-98
//   Film                                    src/film.rs
//   Accel::from, capture, capture_subset, render   src/lib.rs:42-162
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "lasgun_b200.h"

namespace lasgun {

// std::vector whose resize() leaves trivially-constructible elements uninitialised: the big scene arrays are
// filled by all threads right after they are sized, and a sequential zero-fill (plus its page faults) first
// would cost more than the fill itself.
// Their storage comes from the library's cache of freed blocks (lgh_block_alloc: requests of a megabyte and more reuse a block a
// previous scene released): `capture(scene, film)` sizes the same ~100 MB of arrays for every frame, and fresh mappings cost their page
// faults on first touch and an munmap on release -- measured 5-11 ms per flattened million triangles on 8 cores, as much as the fill.
extern "C" void* lgh_block_alloc(size_t bytes);
extern "C" void lgh_block_free(void* p, size_t bytes);
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    T* allocate(size_t n) { return static_cast<T*>(lgh_block_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) noexcept { lgh_block_free(p, n * sizeof(T)); }
    template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

struct Material {                      // material/mod.rs:4-46
    enum Kind : int { Matte = 0, Plastic = 1, Metal = 2, Glass = 3, Mirror = 4 };
    int kind = Matte;
    double kd[3] = {0.5, 0.5, 0.5}, ks[3] = {0, 0, 0};
    double roughness = 0.0;            // matte: sigma; metal: u_roughness; glass: eta
    double roughness_v = 0.0;          // metal: v_roughness
    static Material default_() { return Material{}; }
    static Material matte(const double kd[3], double sigma);
    static Material plastic(const double kd[3], const double ks[3], double roughness);
    static Material metal(const double eta[3], const double k[3], double u_roughness, double v_roughness);   // eta in kd, k in ks
    static Material glass(const double kr[3], const double kt[3], double eta);                               // kr in kd, kt in ks
    static Material mirror(const double kr[3]);                                                              // kr in kd
};

struct ObjData {                       // what the `obj` crate yields for one file
    std::vector<float> positions;      // 3 per vertex
    std::vector<float> normals;        // 3 per normal, empty => Triangle::has_n() is false
    std::vector<uint32_t> faces;       // 3 position indices per polygon (first three only, triangle.rs:40-55)
    std::vector<uint32_t> normal_faces;
};
struct ObjRef { size_t index; };       // scene.rs:42-44

struct Transform { double m[16], minv[16]; bool identity = true; Transform(); };   // column-major, cgmath

class Aggregate {                      // scene/node.rs:24-115
public:
    struct Node {
        enum Kind { Sphere, Cube, Cuboid, Mesh, Group } kind;
        double a[3], b[3], r;
        Material mat; bool has_mat;
        size_t ref;                    // mesh index, or index into `groups`
    };
    std::vector<Node> contents;
    std::vector<Aggregate> groups;
    Transform transform;
    bool swap_backface_flag = false;

    void add_group(Aggregate aggregate);
    void add_sphere(const double center[3], double radius, const Material& material);
    void add_cube(const double origin[3], double dim, const Material& material);
    void add_box(const double minbound[3], const double maxbound[3], const Material& material);
    void add_obj(ObjRef mesh);
    void add_obj_of(ObjRef mesh, const Material& material);
    void swap_backface() { swap_backface_flag = !swap_backface_flag; }
    Aggregate& translate(const double delta[3]);
    Aggregate& scale(double x, double y, double z);
    Aggregate& rotate_x(double theta);
    Aggregate& rotate_y(double theta);
    Aggregate& rotate_z(double theta);
    Aggregate& rotate(double theta, const double axis[3]);
};

class Camera {                         // camera.rs:6-98
public:
    double origin[3] = {0, 0, 0}, view[3] = {0, 0, 1}, up[3] = {0, 1, 0}, aux[3] = {1, 0, 0};
    static Camera perspective(double fov);
    static Camera orthographic(double height);
    void look_at(const double origin[3], const double look[3], const double up[3]);
    void set_supersampling(uint8_t base);
    void set_aperture_radius(double r) { aperture_radius = r; }
    size_t num_samples() const { return root * root; }
    bool is_perspective = true; double param = 45.0;
    size_t root = 1; double distance = 1.0;
    double aperture_radius = 0.0, image_plane_height = 0.0, pixel_separation = 0.0;
private:
    double plane_height(double focal) const;
};

class Scene {                          // scene.rs:11-143
public:
    Aggregate root;
    Camera camera = Camera::perspective(45.0);
    double bg_inner[3] = {0, 0, 0}, bg_outer[3] = {0, 0, 0}, bg_scale = 1.0;
    double ambient[3] = {0, 0, 0};
    bool smoothing = true;
    uint32_t recursion = 3;
    size_t threads = 0;                // kept for API compatibility; the device path ignores it
    struct Light { double position[3], intensity[3], falloff[3]; };
    std::vector<Light> lights;
    std::vector<ObjData> meshes;

    Camera& set_camera(const Camera& c) { camera = c; return camera; }
    Camera& set_perspective_camera(double fov) { camera = Camera::perspective(fov); return camera; }
    Camera& set_orthographic_camera(double scale) { camera = Camera::orthographic(scale); return camera; }
    void set_solid_background(const double c[3]);
    void set_radial_background(const double inner[3], const double outer[3], double scale);
    void set_ambient_light(const double c[3]);
    void set_mesh_smoothing(bool enabled) { smoothing = enabled; }
    void set_max_recursion_depth(uint32_t d) { recursion = d; }
    void set_threads(size_t t) { threads = t; }
    void add_point_light(const double position[3], const double intensity[3], const double falloff[3]);
    ObjRef add_obj(ObjData mesh);
    void set_root(Aggregate node) { root = std::move(node); }
};

class Film {                           // film.rs:7-67
public:
    uint32_t w, h;
    double winv, hinv, aspect;
    Film(uint32_t width, uint32_t height);
    Film(uint32_t width, uint32_t height, uint8_t* external);   // new_with_output
    uint8_t* data() { return ptr; }
    const uint8_t* operator[](size_t at) const { return ptr + 4 * at; }
private:
    std::vector<uint8_t> own;
    uint8_t* ptr;
};

// Flattened scene: the arrays handed to lgb_scene_create (all host memory).
struct FlatScene {
    std::vector<lgb_node> nodes;
    std::vector<uint32_t> prim_refs;
    raw_vector<lgb_sphere> spheres; raw_vector<uint32_t> sphere_material, sphere_id;
    raw_vector<lgb_cuboid> cuboids; raw_vector<uint32_t> cuboid_material, cuboid_id;
    raw_vector<lgb_triangle> triangles; raw_vector<uint32_t> triangle_material, triangle_id;
    raw_vector<lgb_tri_normals> tri_normals; raw_vector<uint8_t> tri_has_normals;
    std::vector<lgb_instance> instances;
    lgb_instance root{};               // transform / swap_backface of the root aggregate
    std::vector<lgb_material> materials;
    std::vector<lgb_light> lights;
    lgb_camera camera{};
    double ambient[3], bg_inner[3], bg_outer[3], bg_scale;
    uint32_t recursion = 3;
    uint32_t flags = 0;
    uint32_t prim_count = 0;           // canonical primitive ids are 0 .. prim_count-1
    // Per BVH level (pre-order of construction): the reference's own arrays, for builder parity tests.
    struct Level { std::vector<double> bounds; std::vector<uint32_t> meta; std::vector<uint64_t> order; uint32_t node_offset; };
    std::vector<Level> levels;
    double build_ms = 0.0;
    double bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};   // the reference root box (root aggregate's coordinates)
    // Lazy mode (BuildOptions::lazy_tree): nodes / prim_refs / instances[].root_node are filled by build_reference_tree(),
    // which lgb_scene_desc::reference_tree calls if the device ever meets an exact-t tie.  The FlatScene must not move
    // (its address is the callback's user pointer) while a device scene created from it is alive.
    bool tree_built = false, keep_levels = false;
    std::shared_ptr<void> pending;     // the assembled levels, until their trees are built
    void build_reference_tree();
    void describe(lgb_scene_desc* out) const;
};

struct BuildOptions {
    bool keep_levels = false;          // keep per-level reference arrays (tests)
    bool lazy_tree = false;            // defer the reference BVH build until the device needs it (exact-t ties only)
    uint32_t expected_film_pixels = 0; // lgb_scene_desc.expected_film_pixels (capture fills it in)
};

FlatScene flatten(const Scene& scene, const BuildOptions& opt);   // Accel::from + flatten (bvh.rs:135-453)

// Accel::from(scene): builds the BVH, flattens it and uploads it to the device of `ctx`.
class Accel {
public:
    static std::unique_ptr<Accel> from(const Scene& scene, lgb_ctx* ctx, const BuildOptions& opt = BuildOptions());
    ~Accel();
    lgb_ctx* ctx = nullptr;
    lgb_scene* dev = nullptr;
    FlatScene flat;
};

lgb_ctx* default_context();            // lazily created context on device 0 (throws lasgun::Error without a GPU)
void capture(const Scene& scene, Film& film);                                   // lib.rs:55
void capture_subset(size_t k, size_t n, const Accel& root, Film& film);         // lib.rs:110
Film render(const Scene& scene, uint32_t width, uint32_t height);               // lib.rs:46

}  // namespace lasgun
