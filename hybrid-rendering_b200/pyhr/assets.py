"""ctypes binding of libhr_assets.so (include/hr_assets.h): scene / asset ingestion in the reference's formats.
Test / bench driver only, like the rest of pyhr."""
import ctypes as C
import os

import numpy as np

from . import BUILD_DIR, HrError, PKG_ROOT, hr_instance, hr_material, hr_material_textures, hr_texture, hr_vertex

LIB_ASSETS = os.path.join(BUILD_DIR, "libhr_assets.so")
_lib = None

TEX_ALBEDO, TEX_NORMAL, TEX_ROUGHNESS, TEX_METALLIC, TEX_EMISSIVE = range(5)


class hra_submesh(C.Structure):
    _fields_ = [("mat_idx", C.c_uint32), ("index_count", C.c_uint32), ("base_vertex", C.c_uint32), ("base_index", C.c_uint32), ("vertex_count", C.c_uint32),
                ("max_extents", C.c_float * 3), ("min_extents", C.c_float * 3)]


def load_assets():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_ASSETS):
            raise RuntimeError(f"{LIB_ASSETS} not found: build it with `make -C {PKG_ROOT}`")
        L = C.CDLL(LIB_ASSETS)
        L.hra_last_error.restype = C.c_char_p
        ip = C.POINTER(C.c_int)
        L.hra_image_load.argtypes = [C.c_char_p, C.c_int, ip, ip, ip, C.POINTER(C.c_void_p)]
        L.hra_image_load_memory.argtypes = [C.c_void_p, C.c_size_t, C.c_int, ip, ip, ip, C.POINTER(C.c_void_p)]
        L.hra_image_loadf.argtypes = [C.c_char_p, C.c_int, ip, ip, C.POINTER(C.c_void_p)]
        L.hra_image_free.argtypes = [C.c_void_p]
        L.hra_image_save_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hra_bluenoise_load.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.hra_brdf_lut_load.argtypes = [C.c_char_p, C.c_void_p]
        L.hra_environment_constant.argtypes = [C.c_char_p, C.POINTER(C.c_float)]
        L.hra_mesh_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.hra_mesh_destroy.argtypes = [C.c_void_p]
        L.hra_mesh_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
        for n in ("hra_mesh_vertices", "hra_mesh_indices", "hra_mesh_materials", "hra_scene_vertices", "hra_scene_indices", "hra_scene_instances", "hra_scene_materials"):
            getattr(L, n).restype = C.c_void_p
            getattr(L, n).argtypes = [C.c_void_p]
        L.hra_mesh_submesh.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(hra_submesh)]
        L.hra_mesh_extents.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.hra_mesh_material_texture.restype = C.c_char_p
        L.hra_mesh_material_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.hra_scene_create.restype = C.c_void_p
        L.hra_scene_destroy.argtypes = [C.c_void_p]
        L.hra_scene_add_instance.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.hra_scene_finalize.argtypes = [C.c_void_p]
        L.hra_scene_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
        L.hra_scene_texture_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 2
        for n in ("hra_scene_textures", "hra_scene_material_textures"):
            getattr(L, n).restype = C.c_void_p
            getattr(L, n).argtypes = [C.c_void_p]
        L.hra_scene_texture_warnings.restype = C.c_char_p
        L.hra_scene_texture_warnings.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise HrError(f"{what} failed ({rc}): {load_assets().hra_last_error().decode(errors='replace')}")


def image_load(path=None, data=None, flip_vertical=False):
    """stbi_load as Image::create_from_file uses it (vk.cpp:136-190): uint8 array (H, W, channels); RGB files come back as RGBA."""
    L = load_assets()
    w, h, c, p = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
    if data is not None:
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _check(L.hra_image_load_memory(buf, len(data), int(flip_vertical), C.byref(w), C.byref(h), C.byref(c), C.byref(p)), "hra_image_load_memory")
    else:
        _check(L.hra_image_load(os.fsencode(path), int(flip_vertical), C.byref(w), C.byref(h), C.byref(c), C.byref(p)), "hra_image_load")
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (h.value, w.value, c.value)).copy()
    finally:
        L.hra_image_free(p)


def image_save_png(path, img):
    """uint8 array (H, W) or (H, W, C), C in 1..4 -> PNG file (e.g. TonemapPass.download(100))"""
    a = np.ascontiguousarray(img, np.uint8)
    h, w = a.shape[:2]
    c = 1 if a.ndim == 2 else a.shape[2]
    _check(load_assets().hra_image_save_png(os.fsencode(path), w, h, c, a.ctypes.data), "hra_image_save_png")


def image_loadf(path, flip_vertical=False):
    """stbi_loadf(path, 4): float32 array (H, W, 4)"""
    L = load_assets()
    w, h, p = C.c_int(), C.c_int(), C.c_void_p()
    _check(L.hra_image_loadf(os.fsencode(path), int(flip_vertical), C.byref(w), C.byref(h), C.byref(p)), "hra_image_loadf")
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (h.value, w.value, 4)).copy()
    finally:
        L.hra_image_free(p)


def bluenoise_load(directory):
    """BlueNoise::BlueNoise (blue_noise.cpp:21-33): (sobol (256, 4) u8, tables (9, 128, 128, 4) u8, bit mask of the slots present)"""
    sobol = np.zeros((256, 4), np.uint8)
    sr = np.zeros((9, 128, 128, 4), np.uint8)
    mask = C.c_uint32()
    _check(load_assets().hra_bluenoise_load(os.fsencode(directory), sobol.ctypes.data, sr.ctypes.data, C.byref(mask)), "hra_bluenoise_load")
    return sobol, sr, mask.value


def brdf_lut_load(path):
    lut = np.zeros((512, 512, 2), np.uint16)
    _check(load_assets().hra_brdf_lut_load(os.fsencode(path), lut.ctypes.data), "hra_brdf_lut_load")
    return lut


def environment_constant(path):
    rgb = (C.c_float * 3)()
    _check(load_assets().hra_environment_constant(os.fsencode(path), rgb), "hra_environment_constant")
    return tuple(rgb[:])


class Mesh:
    """dw::Mesh::load(path)"""

    def __init__(self, path):
        self.lib = load_assets()
        h = C.c_void_p()
        _check(self.lib.hra_mesh_load(os.fsencode(path), C.byref(h)), "hra_mesh_load")
        self.h = h
        c = [C.c_uint64() for _ in range(4)]
        self.lib.hra_mesh_counts(self.h, *[C.byref(x) for x in c])
        self.n_vertices, self.n_indices, self.n_submeshes, self.n_materials = [int(x.value) for x in c]

    def __del__(self):
        try:
            if self.h:
                self.lib.hra_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def vertices(self):
        """(n, 20) float32: position.xyzw | tex_coord | normal | tangent | bitangent (dw::Vertex, mesh.h:16-23)"""
        a = (C.c_float * (20 * self.n_vertices)).from_address(self.lib.hra_mesh_vertices(self.h))
        return np.frombuffer(a, np.float32).reshape(-1, 20).copy()

    def indices(self):
        a = (C.c_uint32 * self.n_indices).from_address(self.lib.hra_mesh_indices(self.h))
        return np.frombuffer(a, np.uint32).copy()

    def materials(self):
        a = (hr_material * self.n_materials).from_address(self.lib.hra_mesh_materials(self.h))
        return [dict(albedo=tuple(m.albedo[:]), emissive=tuple(m.emissive[:]), roughness=m.roughness, metallic=m.metallic) for m in a]

    def submeshes(self):
        out = []
        for i in range(self.n_submeshes):
            s = hra_submesh()
            _check(self.lib.hra_mesh_submesh(self.h, i, C.byref(s)), "hra_mesh_submesh")
            out.append(dict(mat_idx=s.mat_idx, index_count=s.index_count, base_vertex=s.base_vertex, base_index=s.base_index, vertex_count=s.vertex_count,
                            min_extents=tuple(s.min_extents[:]), max_extents=tuple(s.max_extents[:])))
        return out

    def extents(self):
        mn, mx = (C.c_float * 3)(), (C.c_float * 3)()
        self.lib.hra_mesh_extents(self.h, mn, mx)
        return np.array(mn[:], np.float32), np.array(mx[:], np.float32)

    def texture(self, material, kind):
        return self.lib.hra_mesh_material_texture(self.h, material, kind).decode()


class AssetScene:
    """dw::RayTracedScene::create(backend, instances): instances = [(Mesh, model 4x4 column-major, 16 floats)].  Has the
    attributes pyhr.Context.build_scene reads (n_vertices, n_indices, n_instances, n_materials, raw())."""

    def __init__(self, instances):
        self.lib = load_assets()
        self.h = C.c_void_p(self.lib.hra_scene_create())
        self._meshes = [m for m, _ in instances]  # keep alive
        for mesh, model in instances:
            m16 = (C.c_float * 16)(*[float(x) for x in np.asarray(model, np.float32).reshape(-1)])
            _check(self.lib.hra_scene_add_instance(self.h, mesh.h, m16), "hra_scene_add_instance")
        _check(self.lib.hra_scene_finalize(self.h), "hra_scene_finalize")
        c = [C.c_uint64() for _ in range(4)]
        self.lib.hra_scene_counts(self.h, *[C.byref(x) for x in c])
        self.n_vertices, self.n_indices, self.n_instances, self.n_materials = [int(x.value) for x in c]
        self.n_tris = self.n_indices // 3

    def __del__(self):
        try:
            if self.h:
                self.lib.hra_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def raw(self):
        L = self.lib
        return (L.hra_scene_vertices(self.h), L.hra_scene_indices(self.h), L.hra_scene_instances(self.h), L.hra_scene_materials(self.h))

    def textures(self):
        """(textures, bindings, warnings) in the form pyhr.Context.set_textures / oracle.ShadingScene.set_textures take"""
        nt, nb = C.c_uint64(), C.c_uint64()
        self.lib.hra_scene_texture_counts(self.h, C.byref(nt), C.byref(nb))
        tx = (hr_texture * nt.value).from_address(self.lib.hra_scene_textures(self.h)) if nt.value else []
        bd = (hr_material_textures * nb.value).from_address(self.lib.hra_scene_material_textures(self.h)) if nb.value else []
        textures = []
        for t in tx:
            a = np.ctypeslib.as_array(C.cast(t.data, C.POINTER(C.c_uint8)), (t.height, t.width, t.channels)).copy()
            textures.append((a[..., 0] if t.channels == 1 else a, bool(t.srgb)))
        keys = ("albedo", "normal", "roughness", "roughness_channel", "metallic", "metallic_channel", "emissive")
        bindings = [{k: getattr(b, k) for k in keys} for b in bd]
        return textures, bindings, self.lib.hra_scene_texture_warnings(self.h).decode(errors="replace")

    def instances(self):
        a = (hr_instance * self.n_instances).from_address(self.lib.hra_scene_instances(self.h))
        return [dict(model=tuple(i.model[:]), first_index=i.first_index, index_count=i.index_count, base_vertex=i.base_vertex, material_idx=i.material_idx) for i in a]

    def vertices(self):
        a = (C.c_float * (20 * self.n_vertices)).from_address(self.lib.hra_scene_vertices(self.h))
        return np.frombuffer(a, np.float32).reshape(-1, 20).copy()

    def indices(self):
        a = (C.c_uint32 * self.n_indices).from_address(self.lib.hra_scene_indices(self.h))
        return np.frombuffer(a, np.uint32).copy()

    def world_triangles(self):
        """world-space triangle soup in hr_scene_build's primitive order, same arithmetic (hr_api.cu: ((m0*x + m1*y) + m2*z) + m3)"""
        V, I = self.vertices(), self.indices()
        tris, inst = [], []
        for k, it in enumerate(self.instances()):
            M = np.array(it["model"], np.float32).reshape(4, 4).T  # column-major -> M[r, c]
            idx = I[it["first_index"]: it["first_index"] + it["index_count"]].astype(np.int64) + it["base_vertex"]
            p = V[idx, 0:3]
            w = np.empty_like(p)
            for r in range(3):
                w[:, r] = ((M[r, 0] * p[:, 0] + M[r, 1] * p[:, 1]) + M[r, 2] * p[:, 2]) + M[r, 3]
            tris.append(w.reshape(-1, 9))
            inst.append(np.full(len(idx) // 3, k, np.uint32))
        return np.concatenate(tris).astype(np.float32), np.concatenate(inst)
