"""Deterministic synthetic scenes for the five BASELINE.json configs.

Each recipe takes the API module's classes and returns (Scene, (width, height)).
Vertex and sphere data are generated in f64; mesh vertices are then rounded to
f32 exactly as OBJ storage does (src/shape/triangle.rs:41-42).
"""
from __future__ import annotations

import math

import numpy as np

from .api import Aggregate, Material, ObjData, Scene

_MASK = (1 << 64) - 1


def splitmix64_uniform(seed: int, count: int) -> np.ndarray:
    """`count` draws u = (next_u64 >> 11) * 2^-53 of SplitMix64 seeded with `seed`."""
    gamma = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _MASK) + gamma * np.arange(1, count + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def mesh_grid(n: int, s: float, smooth: bool = False) -> ObjData:
    """MESH(N, S) of SURVEY §8d: bumpy sphere, (N+1)^2 vertices (seam duplicated), 2*N^2 triangles.
    smooth: also one vertex normal per vertex (the radial direction, f32), as an OBJ with `vn` records would carry."""
    i = np.arange(n + 1, dtype=np.float64)
    u = 2.0 * math.pi * i / n
    v = math.pi * i / n
    U, V = np.meshgrid(u, v, indexing="xy")            # V varies with row j, U with column i
    rho = s * (1.0 + 0.05 * np.sin(7.0 * U) * np.cos(5.0 * V) + 0.03 * np.sin(23.0 * U + 1.0) * np.sin(19.0 * V))
    pos = np.stack([rho * np.sin(V) * np.cos(U), rho * np.cos(V), rho * np.sin(V) * np.sin(U)], axis=-1)
    pos32 = pos.reshape(-1, 3).astype(np.float32)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")   # j outer, i inner
    a = (jj * (n + 1) + ii).reshape(-1)
    b = a + 1
    c = a + (n + 1) + 1
    d = a + (n + 1)
    faces = np.empty((2 * n * n, 3), np.uint32)
    faces[0::2] = np.stack([a, b, c], axis=-1)
    faces[1::2] = np.stack([a, c, d], axis=-1)
    if smooth:
        nrm = pos.reshape(-1, 3) / np.linalg.norm(pos.reshape(-1, 3), axis=1, keepdims=True)
        return ObjData(pos32, faces, nrm.astype(np.float32), faces.copy())
    return ObjData(pos32, faces)


def star_mesh(radius_inner=70.0, radius_tip=100.0) -> ObjData:
    """60-triangle star: icosahedron with raised face centres (stand-in for the
    Git-LFS-only smstdodeca.obj of src/examples/simple.rs:25)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    verts = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                      [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    verts *= radius_inner / np.linalg.norm(verts[0])
    tris = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
            (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
            (8, 6, 7), (9, 8, 1)]
    pos = [v for v in verts]
    faces = []
    for (a, b, c) in tris:
        ctr = (verts[a] + verts[b] + verts[c]) / 3.0
        ctr *= radius_tip / np.linalg.norm(ctr)
        k = len(pos)
        pos.append(ctr)
        faces += [(a, b, k), (b, c, k), (c, a, k)]
    return ObjData(np.array(pos).astype(np.float32), np.array(faces, np.uint32))


def simple(variant="b", supersampling=2, res=512):
    """C1: src/examples/simple.rs:11-36.  variant 'a' omits the mesh, 'b' substitutes star_mesh()."""
    scene = Scene()
    scene.set_ambient_light([0.2, 0.2, 0.2])
    scene.set_radial_background([0.26, 0.78, 0.67], [0.1, 0.09, 0.33], 0.5)
    camera = scene.set_perspective_camera(45.0)
    camera.look_at([25.0, 0.0, 800.0], [25.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    mat0 = Material.plastic([0.7, 1.0, 0.7], [0.5, 0.7, 0.5], 0.25)
    mat1 = Material.plastic([0.5, 0.5, 0.5], [0.5, 0.7, 0.5], 0.25)
    mat2 = Material.plastic([1.0, 0.6, 0.1], [0.5, 0.7, 0.5], 0.25)
    mat3 = Material.plastic([0.7, 0.6, 1.0], [0.5, 0.4, 0.8], 0.25)
    mesh = scene.add_obj(star_mesh()) if variant == "b" else None
    scene.add_point_light([-100.0, 150.0, 400.0], [0.9, 0.9, 0.9], [1.0, 0.0, 0.0])
    scene.add_point_light([400.0, 100.0, 150.0], [0.7, 0.0, 0.7], [1.0, 0.0, 0.0])
    scene.root.add_sphere([0.0, 0.0, -400.0], 100.0, mat0)
    scene.root.add_sphere([200.0, 50.0, -100.0], 150.0, mat0)
    scene.root.add_sphere([0.0, -1200.0, -500.0], 1000.0, mat1)
    scene.root.add_sphere([-100.0, 25.0, -300.0], 50.0, mat2)
    scene.root.add_sphere([0.0, 100.0, -250.0], 25.0, mat0)
    scene.root.add_cube([-200.0, -125.0, 0.0], 100.0, mat3)
    if mesh is not None:
        scene.root.add_obj_of(mesh, mat2)
    return scene, (res, res)


def mesh1m(n=708, res=1024):
    """C2: one MESH(N, 1) (2*708^2 = 1,002,528 triangles), 2 point lights, 1 spp."""
    scene = Scene()
    scene.set_ambient_light([0.1, 0.1, 0.1])
    scene.set_solid_background([0.05, 0.05, 0.08])
    camera = scene.set_perspective_camera(45.0)
    camera.look_at([0.0, 0.0, 4.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    mat = Material.plastic([0.7, 0.6, 0.5], [0.4, 0.4, 0.4], 0.25)
    mesh = scene.add_obj(mesh_grid(n, 1.0))
    scene.add_point_light([-3.0, 4.0, 5.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    scene.add_point_light([4.0, 2.0, 3.0], [0.5, 0.4, 0.3], [1.0, 0.0, 0.0])
    scene.root.add_obj_of(mesh, mat)
    return scene, (res, res)


def cornell(res=(1920, 1080), supersampling=1):
    """C3: cuboid walls + spheres + cube, plastic only, 4 spp."""
    scene = Scene()
    scene.set_ambient_light([0.2, 0.2, 0.2])
    camera = scene.set_perspective_camera(60.0)
    camera.look_at([0.0, 0.0, 5.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    ks = [0.5, 0.7, 0.5]
    white = Material.plastic([0.9, 0.9, 0.9], ks, 0.25)
    red = Material.plastic([1.0, 0.0, 0.0], ks, 0.25)
    green = Material.plastic([0.0, 1.0, 0.0], ks, 0.25)
    scene.add_point_light([0.0, 1.75, 0.0], [0.9, 0.9, 0.9], [1.0, 0.0, 0.0])
    r = scene.root
    r.add_box([-2.1, -2.1, -2.1], [2.1, -2.0, 2.0], white)      # floor
    r.add_box([-2.1, 2.0, -2.1], [2.1, 2.1, 2.0], white)        # ceiling
    r.add_box([-2.1, -2.1, -2.1], [2.1, 2.1, -2.0], white)      # back
    r.add_box([-2.1, -2.0, -2.0], [-2.0, 2.0, 2.0], red)        # left
    r.add_box([2.0, -2.0, -2.0], [2.1, 2.0, 2.0], green)        # right
    r.add_sphere([1.0, -0.999, 0.0], 1.0, white)
    r.add_cube([-1.999, -1.999, 0.0], 1.0, white)
    r.add_sphere([-0.5, -1.499, 1.0], 0.5, red)
    r.add_sphere([0.3, -1.499, 1.4], 0.5, green)
    return scene, tuple(res)


def _cube_palette():
    mats = []
    for j in range(8):
        kd = [0.8 * (0.25 + 0.75 * ((j >> b) & 1)) for b in range(3)]
        mats.append(Material.plastic(kd, [0.3, 0.3, 0.3], 0.25))
    return mats


def spheres1m(count=1_000_000, res=2048):
    """C4: random sphere field, SplitMix64 seed 0x5EED0004, 3 point lights, 1 spp."""
    scene = Scene()
    scene.set_ambient_light([0.1, 0.1, 0.1])
    scene.set_solid_background([0.0, 0.0, 0.0])
    camera = scene.set_perspective_camera(45.0)
    camera.look_at([0.0, 0.0, 1400.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    u = splitmix64_uniform(0x5EED0004, 4 * count).reshape(count, 4)
    centers = -500.0 + 1000.0 * u[:, 0:3]
    radii = 1.0 + 3.0 * u[:, 3]
    for pos in ([-800.0, 900.0, 1200.0], [900.0, 400.0, 1000.0], [0.0, 1200.0, -200.0]):
        scene.add_point_light(pos, [0.6, 0.6, 0.6], [1.0, 0.0, 0.0])
    scene.root.add_spheres(centers, radii, _cube_palette(), np.arange(count) % 8)
    return scene, (res, res)


def mixed4k(mesh_n=500, nspheres=100_000, res=(3840, 2160), supersampling=3, whitted=False):
    """C5: MESH(500, 100) + 100k spheres on a shell + ground box, 2 lights, 16 spp.
    whitted: the same geometry with a quarter of the spheres in glass, an eighth mirrors, a metal mesh and a mirror ground
    (a large scene for the recursion of integrate.rs:69-132; not one of the five configs)."""
    scene = Scene()
    scene.set_ambient_light([0.1, 0.1, 0.1])
    camera = scene.set_perspective_camera(50.0)
    camera.look_at([0.0, 60.0, 520.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    mesh = scene.add_obj(mesh_grid(mesh_n, 100.0))
    scene.add_point_light([-400.0, 500.0, 600.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    scene.add_point_light([500.0, 300.0, 200.0], [0.5, 0.45, 0.4], [1.0, 0.0, 0.0])
    on = (lambda k: True) if whitted is True else (lambda k: bool(whitted) and k in whitted)      # (a subset, by name, for experiments)
    scene.root.add_obj_of(mesh, Material.metal([0.2, 0.9, 1.1], [3.9, 2.4, 2.2], 0.3, 0.3) if on("metal") else Material.plastic([0.7, 0.6, 0.5], [0.4, 0.4, 0.4], 0.25))
    u = splitmix64_uniform(0x5EED0005, 4 * nspheres).reshape(nspheres, 4)
    theta = np.arccos(1.0 - 2.0 * u[:, 0])
    phi = 2.0 * math.pi * u[:, 1]
    R = 150.0 + 250.0 * u[:, 2]
    rad = 0.5 + 2.5 * u[:, 3]
    centers = np.stack([R * np.sin(theta) * np.cos(phi), R * np.cos(theta), R * np.sin(theta) * np.sin(phi)], axis=-1)
    pal = _cube_palette()
    if on("glass"):
        pal[0] = Material.glass([0.9, 0.9, 0.9], [0.9, 0.95, 0.9], 1.5); pal[4] = Material.glass([1.0, 0.7, 1.0], [0.7, 1.0, 0.7], 1.25)
    if on("mirror"):
        pal[2] = Material.mirror([0.8, 0.8, 0.85])
    if on("matte"):
        pal[6] = Material.matte([0.8, 0.5, 0.3], 30.0)
    scene.root.add_spheres(centers, rad, pal, np.arange(nspheres) % 8)
    scene.root.add_box([-600.0, -130.0, -600.0], [600.0, -120.0, 600.0],
                       Material.mirror([0.6, 0.6, 0.6]) if on("ground") else Material.plastic([0.6, 0.6, 0.6], [0.0, 0.0, 0.0], 0.25))
    return scene, tuple(res)


def plane_mesh() -> ObjData:
    """Unit plane y = 0 over [-1, 1]^2 (stand-in for the Git-LFS-only meshes/plane.obj of src/examples/cornell.rs:25)."""
    pos = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]], np.float32)
    return ObjData(pos, np.array([[0, 2, 1], [0, 3, 2]], np.uint32))


def cornell_groups(res=(512, 512), supersampling=2, eye=(0.0, 0.0, 5.0)):
    """The reference's own Cornell box (src/examples/cornell.rs): five plane meshes inside scaled / rotated /
    translated groups (nested BVH levels with transforms, bvh.rs:462-518); the two glass objects are plastic here.
    With the reference's on-axis eye, the rays of the image diagonals run exactly along the edges where two walls
    meet; pass a slightly off-axis `eye` for a scene without those measure-zero grazing rays."""
    scene = Scene()
    scene.set_ambient_light([0.2, 0.2, 0.2])
    camera = scene.set_perspective_camera(60.0)
    camera.look_at(list(eye), [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    ks = [0.5, 0.7, 0.5]
    white = Material.plastic([0.9, 0.9, 0.9], ks, 0.25)
    red = Material.plastic([1.0, 0.0, 0.0], ks, 0.25)
    green = Material.plastic([0.0, 1.0, 0.0], ks, 0.25)
    plane = scene.add_obj(plane_mesh())
    scene.add_point_light([0.0, 1.75, 0.0], [0.9, 0.9, 0.9], [1.0, 0.0, 0.0])
    for ops, mat in ((("translate", [0.0, -2.0, 0.0]),), white), ((("translate", [0.0, 2.0, 0.0]),), white), \
                    ((("rotate_z", 90.0), ("translate", [-2.0, 0.0, 0.0])), red), ((("rotate_z", 90.0), ("translate", [2.0, 0.0, 0.0])), green), \
                    ((("rotate_x", 90.0), ("translate", [0.0, 0.0, -2.0])), white):
        g = Aggregate()
        g.scale(2.0, 1.0, 2.0)                                     # cornell.rs:32-64
        for name, arg in ops:
            getattr(g, name)(arg)
        g.add_obj_of(plane, mat)
        scene.root.add_group(g)
    scene.root.add_sphere([1.0, -1.25, 0.0], 1.0, Material.plastic([1.0, 0.7, 1.0], [0.7, 1.0, 0.7], 0.25))
    scene.root.add_cube([-1.999, -1.999, 0.0], 1.0, Material.plastic([0.7, 0.7, 1.0], [0.4, 0.4, 0.4], 0.25))
    return scene, tuple(res)


def coincident_planes(nspheres=34000, res=(160, 120), supersampling=0):
    """Every ray that reaches the floor meets TWO coincident triangles at bit-identical t (the mesh lists each of its two
    triangles twice), and the walls are cuboids sharing edges: the reference's "first primitive tested wins"
    (triangle.rs:251) decides most pixels.  The sphere field only makes the scene large enough for the device-side builder."""
    scene = Scene()
    scene.set_ambient_light([0.2, 0.2, 0.2])
    camera = scene.set_perspective_camera(55.0)
    camera.look_at([0.3, 2.5, 6.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    scene.add_point_light([2.0, 6.0, 3.0], [0.9, 0.9, 0.9], [1.0, 0.0, 0.0])
    pos = np.array([[-4, 0, -4], [4, 0, -4], [4, 0, 4], [-4, 0, 4]], np.float32)
    faces = np.array([[0, 2, 1], [0, 3, 2], [0, 2, 1], [0, 3, 2]], np.uint32)
    floor = scene.add_obj(ObjData(pos, faces))
    pal = _cube_palette()
    scene.root.add_obj_of(floor, pal[3])
    scene.root.add_box([-1.0, 0.0, -1.0], [0.0, 1.0, 0.0], pal[5]); scene.root.add_box([0.0, 0.0, -1.0], [1.0, 1.0, 0.0], pal[6])   # share a face
    u = splitmix64_uniform(0x5EED00C0, 4 * nspheres).reshape(nspheres, 4)
    centers = np.stack([-4.0 + 8.0 * u[:, 0], 2.0 + 3.0 * u[:, 1], -6.0 + 4.0 * u[:, 2]], axis=-1)
    scene.root.add_spheres(centers, 0.01 + 0.02 * u[:, 3], pal, np.arange(nspheres) % 8)
    return scene, tuple(res)


def nested_groups(res=(320, 240), supersampling=1, transformed_root=False, mesh_n=24):
    """Transforms two levels deep, a swap_backface level, spheres / cubes / a mesh inside transformed groups, and
    (optionally) a transformed ROOT aggregate: every branch of the nested-level code (bvh.rs:462-518)."""
    scene = Scene()
    scene.set_ambient_light([0.15, 0.15, 0.15])
    scene.set_radial_background([0.2, 0.3, 0.5], [0.05, 0.05, 0.1], 0.6)
    camera = scene.set_perspective_camera(50.0)
    camera.look_at([3.0, 4.0, 14.0], [0.0, 0.5, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    scene.add_point_light([6.0, 9.0, 8.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    scene.add_point_light([-7.0, 5.0, 3.0], [0.4, 0.5, 0.6], [1.0, 0.0, 0.0])
    pal = _cube_palette()
    mesh = scene.add_obj(mesh_grid(mesh_n, 1.0))
    r = scene.root
    r.add_box([-8.0, -1.2, -8.0], [8.0, -1.0, 8.0], Material.plastic([0.6, 0.6, 0.6], [0.1, 0.1, 0.1], 0.3))
    r.add_sphere([0.0, 0.0, 0.0], 1.0, pal[1])
    a = Aggregate()                                   # level 1: scale + rotate + translate
    a.scale(1.5, 0.7, 1.2); a.rotate_y(30.0); a.translate([3.0, 0.4, -1.0])
    a.add_sphere([0.0, 0.0, 0.0], 1.0, pal[2]); a.add_sphere([1.6, 0.3, 0.2], 0.6, pal[3]); a.add_cube([-2.5, -1.0, -0.5], 1.0, pal[4])
    b = Aggregate()                                   # level 2, inside level 1, with flipped normals
    b.rotate_z(45.0); b.translate([0.0, 2.2, 0.0]); b.swap_backface()
    b.add_obj_of(mesh, pal[5]); b.add_sphere([1.4, 0.0, 0.0], 0.4, pal[6])
    a.add_group(b)
    r.add_group(a)
    c = Aggregate()                                   # an identity group (traversed inline) holding a transformed one
    d = Aggregate(); d.rotate(70.0, [0.0, 0.6, 0.8]); d.translate([-3.5, 0.8, 1.5])
    d.add_cube([-0.5, -0.5, -0.5], 1.0, pal[7]); d.add_obj_of(mesh, pal[0])
    c.add_group(d); c.add_sphere([-1.5, -0.4, 3.0], 0.6, pal[3])
    r.add_group(c)
    if transformed_root:
        r.rotate_y(-20.0); r.translate([0.5, 0.0, -1.0])
    return scene, tuple(res)


def simplereflect(supersampling=2, res=512, recursion=4):
    """The reference's src/examples/simplereflect.rs: the `simple` layout in glass and mirror, Whitted depth 4
    (star_mesh() stands in for the Git-LFS-only smstdodeca.obj, as in `simple`)."""
    scene = Scene()
    scene.set_ambient_light([0.2, 0.2, 0.2])
    scene.set_radial_background([0.93, 0.87, 0.36], [0.94, 0.6, 0.1], 0.5)
    scene.set_max_recursion_depth(recursion)
    camera = scene.set_perspective_camera(45.0)
    camera.look_at([25.0, 0.0, 800.0], [25.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    mat0 = Material.glass([0.7, 1.0, 0.7], [0.5, 0.7, 0.5], 1.333)
    mat1 = Material.mirror([0.5, 0.5, 0.5])
    mat2 = Material.glass([1.0, 0.6, 0.1], [0.7, 0.7, 1.0], 1.75)
    mat3 = Material.glass([0.7, 0.6, 1.0], [0.5, 0.4, 0.8], 1.5)
    mesh = scene.add_obj(star_mesh())
    scene.add_point_light([-100.0, 150.0, 400.0], [0.9, 0.9, 0.9], [1.0, 0.0, 0.0])
    scene.add_point_light([400.0, 100.0, 150.0], [0.7, 0.0, 0.7], [1.0, 0.0, 0.0])
    scene.root.add_sphere([0.0, 0.0, -400.0], 100.0, mat0)
    scene.root.add_sphere([200.0, 50.0, -100.0], 150.0, mat0)
    scene.root.add_sphere([0.0, -1200.0, -500.0], 1000.0, mat1)
    scene.root.add_sphere([-100.0, 25.0, -300.0], 50.0, mat2)
    scene.root.add_sphere([0.0, 100.0, -250.0], 25.0, mat0)
    scene.root.add_cube([-200.0, -125.0, 0.0], 100.0, mat3)
    scene.root.add_obj_of(mesh, mat2)
    return scene, (res, res)


def materials(res=(384, 288), supersampling=1, recursion=3, grouped=False):
    """Every `Material` variant (material/mod.rs:19-46) on every shape: Oren-Nayar matte, isotropic and anisotropic metal (the
    parameters of src/examples/playground.rs:15 and simplecows.rs:18 among them), glass, mirror, plastic; a smooth mesh (vertex
    normals: the shading frame of bsdf.rs:33-36 is not orthonormal there) in metal and in glass.  grouped: the specular objects sit
    in a rotated, scaled group, so the rays below them cross instance transforms."""
    scene = Scene()
    scene.set_ambient_light([0.15, 0.15, 0.15])
    scene.set_radial_background([0.35, 0.5, 0.8], [0.9, 0.85, 0.7], 0.7)
    scene.set_max_recursion_depth(recursion)
    camera = scene.set_perspective_camera(48.0)
    camera.look_at([0.5, 3.2, 11.0], [0.0, 0.6, 0.0], [0.0, 1.0, 0.0])
    camera.set_supersampling(supersampling)
    scene.add_point_light([6.0, 9.0, 8.0], [0.8, 0.8, 0.8], [1.0, 0.0, 0.0])
    scene.add_point_light([-7.0, 6.0, 4.0], [0.5, 0.45, 0.4], [1.0, 0.02, 0.0])
    smooth = scene.add_obj(mesh_grid(20, 1.0, smooth=True))
    flat = scene.add_obj(star_mesh(0.7, 1.0))
    r = scene.root
    r.add_box([-9.0, -1.2, -9.0], [9.0, -1.0, 9.0], Material.matte([0.7, 0.65, 0.6], 25.0))                 # Oren-Nayar floor
    r.add_box([-9.0, -1.0, -9.2], [9.0, 6.0, -9.0], Material.mirror([0.85, 0.85, 0.9]))                      # mirror back wall
    r.add_sphere([-4.0, 0.0, 0.5], 1.0, Material.metal([0.9, 0.1, 0.9], [0.7, 1.0, 0.7], 0.25, 0.25))        # playground.rs:15
    r.add_sphere([-1.6, 0.0, 1.5], 1.0, Material.metal([0.2, 0.9, 1.1], [3.9, 2.4, 2.2], 0.08, 0.45))        # anisotropic
    r.add_cube([3.2, -1.0, -1.5], 1.6, Material.metal([0.0, 0.0, 0.0], [0.7, 0.7, 0.7], 0.5, 0.5))           # simplecows.rs:18
    r.add_sphere([4.2, 0.1, 2.6], 0.9, Material.matte([0.9, 0.3, 0.2], 60.0))
    r.add_sphere([-3.0, -0.5, 3.8], 0.5, Material.plastic([0.2, 0.7, 0.3], [0.5, 0.5, 0.5], 0.2))
    g = Aggregate() if grouped else r
    if grouped:
        g.scale(1.0, 1.15, 0.9); g.rotate_y(25.0); g.translate([0.2, 0.15, 0.3])
    g.add_sphere([1.0, 0.2, 2.2], 1.2, Material.glass([1.0, 0.7, 1.0], [0.7, 1.0, 0.7], 1.25))               # cornell.rs:22
    g.add_cube([-0.6, -1.0, 3.6], 0.9, Material.glass([0.7, 0.6, 1.0], [0.8, 0.8, 0.8], 1.333))              # spooky.rs:22
    g.add_sphere([2.6, -0.4, 4.4], 0.6, Material.mirror([0.5, 0.5, 0.5]))
    g.add_sphere([-2.2, 2.2, -2.0], 0.9, Material.glass([0.0, 0.0, 0.0], [0.9, 0.9, 0.9], 1.5))              # kr = 0: transmission lobe only
    g.add_sphere([0.2, 2.6, -3.0], 0.8, Material.glass([0.9, 0.9, 0.9], [0.0, 0.0, 0.0], 1.5))               # kt = 0: reflection lobe only
    m1 = Aggregate(); m1.translate([2.3, 2.3, -1.0]); m1.add_obj_of(smooth, Material.metal([0.3, 0.5, 0.9], [2.5, 2.0, 1.6], 0.3, 0.12))
    m2 = Aggregate(); m2.scale(0.9, 0.9, 0.9); m2.translate([-0.4, 1.9, 0.6]); m2.add_obj_of(smooth, Material.glass([0.8, 0.8, 0.8], [0.9, 0.9, 0.9], 1.4))
    m3 = Aggregate(); m3.translate([-4.6, 1.9, -1.5]); m3.add_obj_of(flat, Material.matte([0.4, 0.5, 0.9], 35.0))
    g.add_group(m1); g.add_group(m2); g.add_group(m3)
    if grouped:
        r.add_group(g)
    return scene, tuple(res)


CONFIGS = {
    "simple": lambda: simple("b", 2),
    "simple_nomesh": lambda: simple("a", 2),
    "simple_1spp": lambda: simple("b", 0),
    "mesh1m": mesh1m,
    "cornell": cornell,
    "spheres1m": spheres1m,
    "mixed4k": mixed4k,
    "simplereflect": simplereflect,
    "materials": materials,
    "simplereflect2k": lambda: simplereflect(2, 2048),                     # 4.2 Mpixel, 9 spp, depth 4: 37.7 M ray trees
    "materials2k": lambda: materials((2560, 1920), 1),                     # 4.9 Mpixel, 4 spp
    "mixed4k_whitted": lambda: mixed4k(whitted=True),                      # the C5 geometry in glass / mirror / metal, 16 spp, depth 3
}
