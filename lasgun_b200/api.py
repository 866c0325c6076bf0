"""Host-side mirror of lasgun's public scene-builder surface.

Same names and argument meaning as the reference's Rust API
(src/scene.rs:49-143, src/scene/node.rs:35-115, src/material/mod.rs:12-46,
src/camera.rs:85-98, src/film.rs:22-45).  The objects here only *describe* a
scene; `lasgun_b200.capture` replays the description into the C++ host
library (BVH build + flatten) which calls the CUDA path through the C ABI.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Material", "ObjData", "ObjRef", "Aggregate", "Camera", "Scene", "Film", "parse_obj_text"]

MAT_MATTE, MAT_PLASTIC, MAT_METAL, MAT_GLASS, MAT_MIRROR = 0, 1, 2, 3, 4


def _v3(v):
    v = [float(c) for c in v]
    if len(v) != 3:
        raise ValueError("expected 3 components")
    return v


class Material:
    """material/mod.rs:4-46.  Value type, copied into each primitive."""

    __slots__ = ("kind", "kd", "ks", "roughness", "roughness2")

    def __init__(self, kind, kd, ks, roughness, roughness2=0.0):
        self.kind, self.kd, self.ks, self.roughness, self.roughness2 = kind, _v3(kd), _v3(ks), float(roughness), float(roughness2)

    @staticmethod
    def default():                       # mod.rs:15-17
        return Material.matte([0.5, 0.5, 0.5], 0.0)

    @staticmethod
    def matte(kd, sigma):                # mod.rs:19-22, matte.rs:14-16 (sigma clamped to [0, 90])
        return Material(MAT_MATTE, kd, [0, 0, 0], min(max(float(sigma), 0.0), 90.0))

    @staticmethod
    def plastic(kd, ks, roughness):      # mod.rs:24-28
        return Material(MAT_PLASTIC, kd, ks, roughness)

    # One record for the five variants: metal keeps eta in kd, k in ks and the two roughnesses (metal.rs:13-15); glass keeps
    # kr in kd, kt in ks and eta in roughness (mod.rs:36-41: its microfacet roughness is always 0); mirror keeps kr in kd.
    @staticmethod
    def metal(eta, k, u_roughness, v_roughness):      # mod.rs:30-34
        return Material(MAT_METAL, eta, k, u_roughness, v_roughness)

    @staticmethod
    def glass(kr, kt, eta):                           # mod.rs:36-41
        return Material(MAT_GLASS, kr, kt, eta)

    @staticmethod
    def mirror(kr):                                   # mod.rs:43-46
        return Material(MAT_MIRROR, kr, [0, 0, 0], 0.0)


class ObjData:
    """What the `obj` crate hands to lasgun for one file: f32 positions / normals and
    polygons, of which only the first three index tuples are used (triangle.rs:40-55)."""

    def __init__(self, positions, faces, normals=None, normal_faces=None):
        self.positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        if normals is not None and len(normals):
            self.normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
            self.normal_faces = np.ascontiguousarray(normal_faces, dtype=np.uint32).reshape(-1, 3)
            if self.normal_faces.shape != self.faces.shape:
                raise ValueError("normal_faces must match faces")
        else:
            self.normals = np.zeros((0, 3), np.float32)
            self.normal_faces = np.zeros((0, 3), np.uint32)
        if self.faces.size and int(self.faces.max()) >= len(self.positions):
            raise ValueError("face index out of range")


def parse_obj_text(text):
    """Minimal `v` / `vn` / `f` reader standing in for the `obj` crate (triangle.rs:385-395).
    Polygons keep their first three vertices, as Triangle does."""
    pos, nrm, faces, nfaces = [], [], [], []
    for line in text.splitlines():
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "v":
            pos.append([float(t) for t in tok[1:4]])
        elif tok[0] == "vn":
            nrm.append([float(t) for t in tok[1:4]])
        elif tok[0] == "f":
            vi, ni = [], []
            for t in tok[1:4]:
                parts = t.split("/")
                i = int(parts[0])
                vi.append(i - 1 if i > 0 else len(pos) + i)
                if len(parts) > 2 and parts[2]:
                    j = int(parts[2])
                    ni.append(j - 1 if j > 0 else len(nrm) + j)
            if len(vi) < 3:
                raise ValueError("ObjError: polygon with fewer than 3 vertices")
            faces.append(vi)
            if len(ni) == 3:
                nfaces.append(ni)
    if nrm and len(nfaces) != len(faces):
        raise ValueError("ObjError: mesh has normals but a face lacks normal indices")
    return ObjData(pos, faces, nrm if nrm else None, nfaces if nrm else None)


class ObjRef:
    """scene.rs:42-44: opaque reference to a mesh registered on a Scene."""
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index


class Aggregate:
    """scene/node.rs:24-115."""

    def __init__(self):
        self.contents = []        # ("sphere", c, r, mat) | ("cube", o, dim, mat) | ("box", a, b, mat) | ("mesh", ref, mat|None) | ("group", Aggregate)
        self.transforms = []      # recorded in call order; the backends replay concat_self
        self._swap_backface = False

    def add_group(self, aggregate):
        self.contents.append(("group", aggregate))

    def add_sphere(self, center, radius, material):
        self.contents.append(("sphere", _v3(center), float(radius), material))

    def add_spheres(self, centers, radii, materials, material_index):
        """Bulk form of add_sphere (host convenience, not in the reference): sphere i gets
        materials[material_index[i]].  Equivalent to calling add_sphere in index order."""
        centers = np.ascontiguousarray(centers, np.float64).reshape(-1, 3)
        radii = np.ascontiguousarray(radii, np.float64).reshape(-1)
        material_index = np.ascontiguousarray(material_index, np.int32).reshape(-1)
        if not (len(centers) == len(radii) == len(material_index)):
            raise ValueError("centers, radii and material_index must have equal length")
        self.contents.append(("spheres", centers, radii, list(materials), material_index))

    def add_cube(self, origin, dim, material):
        self.contents.append(("cube", _v3(origin), float(dim), material))

    def add_box(self, minbound, maxbound, material):
        self.contents.append(("box", _v3(minbound), _v3(maxbound), material))

    def add_obj(self, mesh):
        self.contents.append(("mesh", mesh, None))

    def add_obj_of(self, mesh, material):
        self.contents.append(("mesh", mesh, material))

    def swap_backface(self):
        self._swap_backface = not self._swap_backface

    def translate(self, delta):
        self.transforms.append(("translate", _v3(delta))); return self

    def scale(self, x, y, z):
        self.transforms.append(("scale", [float(x), float(y), float(z)])); return self

    def rotate_x(self, theta):
        self.transforms.append(("rotate_axis", 0, float(theta))); return self

    def rotate_y(self, theta):
        self.transforms.append(("rotate_axis", 1, float(theta))); return self

    def rotate_z(self, theta):
        self.transforms.append(("rotate_axis", 2, float(theta))); return self

    def rotate(self, theta, axis):
        self.transforms.append(("rotate", float(theta), _v3(axis))); return self


class Camera:
    """camera.rs:6-98 (state only; ray generation happens in the backends)."""

    def __init__(self, perspective=True, param=45.0):
        self.perspective, self.param = perspective, float(param)
        self.look = None
        self.supersampling = 0
        self.aperture_radius = 0.0

    @staticmethod
    def perspective_camera(fov):
        return Camera(True, fov)

    @staticmethod
    def orthographic_camera(height):
        return Camera(False, height)

    def look_at(self, origin, look, up):
        self.look = (_v3(origin), _v3(look), _v3(up))

    def set_supersampling(self, base):
        if not 0 <= int(base) < 255:
            raise ValueError("supersampling base must be in [0, 255)")
        self.supersampling = int(base)

    def set_aperture_radius(self, radius):   # stored, unused by the reference (camera.rs:142)
        self.aperture_radius = float(radius)

    def num_samples(self):
        return (self.supersampling + 1) ** 2


class Scene:
    """scene.rs:11-143."""

    def __init__(self):
        self.root = Aggregate()
        self.camera = Camera()
        self.background = ([0.0, 0.0, 0.0], [0.0, 0.0, 0.0], 1.0)
        self.ambient = [0.0, 0.0, 0.0]
        self.smoothing = True
        self.recursion = 3
        self.threads = 0
        self.lights = []
        self.meshes = []

    def set_camera(self, camera):
        self.camera = camera; return self.camera

    def set_perspective_camera(self, fov):
        self.camera = Camera(True, fov); return self.camera

    def set_orthographic_camera(self, scale):
        self.camera = Camera(False, scale); return self.camera

    def set_solid_background(self, color):
        c = _v3(color); self.background = (c, list(c), 1.0)

    def set_radial_background(self, inner, outer, scale):
        self.background = (_v3(inner), _v3(outer), float(scale))

    def set_ambient_light(self, color):
        self.ambient = _v3(color)

    def set_mesh_smoothing(self, enabled):
        self.smoothing = bool(enabled)

    def set_max_recursion_depth(self, max_depth):
        self.recursion = int(max_depth)

    def set_threads(self, threads):
        self.threads = int(threads)

    def add_point_light(self, position, intensity, falloff):
        self.lights.append((_v3(position), _v3(intensity), _v3(falloff)))

    def add_obj(self, mesh: ObjData):
        if not self.smoothing:                 # scene.rs:111
            mesh = ObjData(mesh.positions, mesh.faces)
        self.meshes.append(mesh)
        return ObjRef(len(self.meshes) - 1)

    def parse_obj(self, text):
        return self.add_obj(parse_obj_text(text))

    def load_obj(self, path):
        with open(path) as f:
            return self.parse_obj(f.read())

    def set_root(self, node):
        self.root = node


class Film:
    """film.rs:7-67: owns a row-major RGBA8 buffer, zero-initialised."""

    def __init__(self, width, height, output=None):
        self.w, self.h = int(width), int(height)
        if output is None:
            output = np.zeros((self.h, self.w, 4), np.uint8)
        if output.dtype != np.uint8 or output.size != self.w * self.h * 4 or not output.flags["C_CONTIGUOUS"]:
            raise ValueError("output must be a contiguous uint8 buffer of w*h*4 bytes")
        self.output = output.reshape(self.h, self.w, 4)
        self.winv, self.hinv = 1.0 / self.w, 1.0 / self.h
        self.aspect = self.w / self.h

    @staticmethod
    def new_with_output(width, height, output):
        return Film(width, height, output)

    def pixels(self):
        return self.output

    def __getitem__(self, at):
        return self.output.reshape(-1, 4)[at]

    def __len__(self):
        return self.w * self.h
