"""Host-side scene description for the harness (numpy -> yh_scene_desc).

Mirrors the reference's `yocto::pathtrace` scene structs
(libs/yocto_pathtrace/yocto_pathtrace.h:272-416) as plain Python records so
tests and bench.py can hand the SAME description to the CUDA library and to
the CPU checkers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _capi as capi

IDENTITY_FRAME = (1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)


def _f32(a, cols):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a.reshape(-1, cols) if cols > 1 else a.reshape(-1)


def _i32(a, cols):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32)).reshape(-1, cols)


def normalize(v):
    v = np.asarray(v, dtype=np.float32)
    l = np.float32(np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1]) + np.float32(v[2] * v[2])))
    return v / l if l != 0 else v


def lookat_frame(eye, center, up, inv_xz=False):
    """yocto::math::lookat_frame, libs/yocto/yocto_math.h:3229-3239 (fp32)."""
    eye, center, up = (np.asarray(x, dtype=np.float32) for x in (eye, center, up))
    w = normalize(eye - center)
    u = normalize(np.cross(up, w).astype(np.float32))
    v = normalize(np.cross(w, u).astype(np.float32))
    if inv_xz:
        w, u = -w, -u
    return tuple(float(x) for x in np.concatenate([u, v, w, eye]))


@dataclass
class Shape:
    positions: np.ndarray
    normals: Optional[np.ndarray] = None
    texcoords: Optional[np.ndarray] = None
    radius: Optional[np.ndarray] = None
    lines: Optional[np.ndarray] = None
    triangles: Optional[np.ndarray] = None
    points: Optional[np.ndarray] = None

    def __post_init__(self):
        self.positions = _f32(self.positions, 3)
        self.normals = _f32(self.normals, 3) if self.normals is not None else np.zeros((0, 3), np.float32)
        self.texcoords = _f32(self.texcoords, 2) if self.texcoords is not None else np.zeros((0, 2), np.float32)
        self.radius = _f32(self.radius, 1) if self.radius is not None else np.zeros((0,), np.float32)
        self.lines = _i32(self.lines, 2) if self.lines is not None else np.zeros((0, 2), np.int32)
        self.triangles = _i32(self.triangles, 3) if self.triangles is not None else np.zeros((0, 3), np.int32)
        self.points = np.ascontiguousarray(self.points, dtype=np.int32).reshape(-1) if self.points is not None else np.zeros((0,), np.int32)


@dataclass
class Texture:
    colorf: np.ndarray  # (H, W, 3) float32, row 0 on top

    def __post_init__(self):
        self.colorf = np.ascontiguousarray(self.colorf, dtype=np.float32)
        assert self.colorf.ndim == 3 and self.colorf.shape[2] == 3


@dataclass
class Material:
    """Defaults = yocto::pathtrace::material, yocto_pathtrace.h:293-329."""
    emission: tuple = (0, 0, 0)
    color: tuple = (0, 0, 0)
    specular: float = 0
    roughness: float = 0
    metallic: float = 0
    ior: float = 1.5
    spectint: tuple = (1, 1, 1)
    transmission: float = 0
    scattering: tuple = (0, 0, 0)
    scanisotropy: float = 0
    trdepth: float = 0.01
    opacity: float = 1
    thin: bool = True  # sceneio::material::thin defaults true (yocto_sceneio.h:148)
    eumelanin: float = 0
    pheomelanin: float = 0
    sigma_a: tuple = (0, 0, 0)
    beta_m: float = 0.3
    beta_n: float = 0.3
    alpha: float = 2
    eta: float = 1.55
    emission_tex: int = -1
    color_tex: int = -1
    specular_tex: int = -1
    metallic_tex: int = -1
    roughness_tex: int = -1
    transmission_tex: int = -1  # the reference reads its EMISSION texture for this factor (yocto_pathtrace.cpp:420-421)
    scattering_tex: int = -1
    opacity_tex: int = -1
    normal_tex: int = -1


@dataclass
class Object:
    shape: int
    material: int
    frame: tuple = IDENTITY_FRAME


@dataclass
class Environment:
    emission: tuple = (1, 1, 1)
    emission_tex: int = -1
    frame: tuple = IDENTITY_FRAME


@dataclass
class Camera:
    """Defaults = yocto::pathtrace::camera, yocto_pathtrace.h:272-278."""
    frame: tuple = IDENTITY_FRAME
    lens: float = 0.050
    film: tuple = (0.036, 0.024)
    focus: float = 10000
    aperture: float = 0
    aspect: float = 1.5      # the (aspect, film) pair `film` was derived from (scene JSON keys)
    film_size: float = 0.036

    @staticmethod
    def from_lookat(eye, center, up=(0, 1, 0), lens=0.05, aspect=1.0, film=0.036, aperture=0.0):
        """Camera set-up of the JSON loader: lookat overrides frame and sets
        focus = |eye - center| (libs/yocto/yocto_sceneio.cpp:1244-1249);
        film from aspect as in set_lens (yocto_pathtrace.cpp:2078-2083)."""
        eye32, c32 = np.asarray(eye, np.float32), np.asarray(center, np.float32)
        d = eye32 - c32
        focus = float(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1]) + np.float32(d[2] * d[2])))
        filmv = (film, np.float32(film) / np.float32(aspect)) if aspect >= 1 else (np.float32(film) * np.float32(aspect), film)
        return Camera(frame=lookat_frame(eye, center, up), lens=lens,
                      film=(float(filmv[0]), float(filmv[1])), focus=focus, aperture=aperture,
                      aspect=float(aspect), film_size=float(film))

    def to_desc(self) -> capi.CameraDesc:
        d = capi.CameraDesc()
        d.frame[:] = [float(x) for x in self.frame]
        d.lens = self.lens
        d.film[:] = [float(x) for x in self.film]
        d.focus = self.focus
        d.aperture = self.aperture
        return d


@dataclass
class Scene:
    shapes: List[Shape] = field(default_factory=list)
    textures: List[Texture] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    objects: List[Object] = field(default_factory=list)
    environments: List[Environment] = field(default_factory=list)

    def to_desc(self):
        """Returns (yh_scene_desc, keepalive) — keepalive owns the ctypes arrays."""
        keep = []

        def fptr(a):
            keep.append(a)
            return a.ctypes.data_as(capi.c_float_p) if a.size else None

        def iptr(a):
            keep.append(a)
            return a.ctypes.data_as(capi.c_int_p) if a.size else None

        shapes = (capi.ShapeDesc * max(1, len(self.shapes)))()
        for d, s in zip(shapes, self.shapes):
            d.positions, d.num_positions = fptr(s.positions), len(s.positions)
            d.normals, d.num_normals = fptr(s.normals), len(s.normals)
            d.texcoords, d.num_texcoords = fptr(s.texcoords), len(s.texcoords)
            d.radius, d.num_radius = fptr(s.radius), len(s.radius)
            d.lines, d.num_lines = iptr(s.lines), len(s.lines)
            d.triangles, d.num_triangles = iptr(s.triangles), len(s.triangles)
            d.points, d.num_points = iptr(s.points), len(s.points)
        textures = (capi.TextureDesc * max(1, len(self.textures)))()
        for d, t in zip(textures, self.textures):
            d.colorf, d.height, d.width = fptr(t.colorf), t.colorf.shape[0], t.colorf.shape[1]
        materials = (capi.MaterialDesc * max(1, len(self.materials)))()
        for d, m in zip(materials, self.materials):
            d.emission[:] = m.emission
            d.color[:] = m.color
            d.specular, d.roughness, d.metallic, d.ior = m.specular, m.roughness, m.metallic, m.ior
            d.spectint[:] = m.spectint
            d.transmission = m.transmission
            d.scattering[:] = m.scattering
            d.scanisotropy, d.trdepth, d.opacity, d.thin = m.scanisotropy, m.trdepth, m.opacity, int(m.thin)
            d.eumelanin, d.pheomelanin = m.eumelanin, m.pheomelanin
            d.sigma_a[:] = m.sigma_a
            d.beta_m, d.beta_n, d.alpha, d.eta = m.beta_m, m.beta_n, m.alpha, m.eta
            d.spectint_tex = -1
            for name in ("emission_tex", "color_tex", "specular_tex", "metallic_tex", "roughness_tex", "transmission_tex",
                         "scattering_tex", "opacity_tex", "normal_tex"):
                setattr(d, name, getattr(m, name))
        objects = (capi.ObjectDesc * max(1, len(self.objects)))()
        for d, o in zip(objects, self.objects):
            d.frame[:] = [float(x) for x in o.frame]
            d.shape, d.material = o.shape, o.material
        envs = (capi.EnvironmentDesc * max(1, len(self.environments)))()
        for d, e in zip(envs, self.environments):
            d.frame[:] = [float(x) for x in e.frame]
            d.emission[:] = e.emission
            d.emission_tex = e.emission_tex
        desc = capi.SceneDesc()
        desc.shapes, desc.num_shapes = shapes, len(self.shapes)
        desc.textures, desc.num_textures = textures, len(self.textures)
        desc.materials, desc.num_materials = materials, len(self.materials)
        desc.objects, desc.num_objects = objects, len(self.objects)
        desc.environments, desc.num_environments = envs, len(self.environments)
        keep += [shapes, textures, materials, objects, envs]
        return desc, keep


def make_params(resolution=720, samples=512, bounces=8, clamp=100.0, seed=961748941,
                shader=1) -> capi.TraceParams:
    """yocto::pathtrace::trace_params defaults, yocto_pathtrace.h:188-197."""
    p = capi.TraceParams()
    p.resolution, p.shader, p.samples, p.bounces = resolution, shader, samples, bounces
    p.clamp, p.seed, p.noparallel, p.pratio = clamp, seed, 0, 8
    return p
