"""Procedural hair scenes (the Bitterli / pbrt hair models are not available
offline; SURVEY.md section 8d fixes the synthetic stand-ins).

Geometry rule = the reference's pbrt `curve` importer
(libs/yocto/yocto_pbrt.h:1751-1796): every cubic Bezier becomes `sub` line
segments over `sub + 1` vertices that are NOT shared with the neighbouring
Bezier; positions = interpolate_bezier(i / sub), tangents (stored in the PLY
normal slot) = normalised Bezier derivative with end tangents p1 - p0 and
p3 - p2, radius = lerp(width0, width1, i / sub) (radius = pbrt width).

Scenes mirror the reference's test JSONs for cameras / lights:
  straight_hair_scene  tests/straight-hair/straight-hair.json  (config C2)
  curly_hair_scene     tests/curly-hair/curly-hair.json        (config C3, C5)
  hair_block_scene     tests/sphere-hairblock/sphere-hairblock.json (C1, C4)
"""
from __future__ import annotations

import numpy as np

from .scene import Camera, Environment, Material, Object, Scene, Shape, Texture, lookat_frame

F = np.float32


def _normalize(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return (v / np.maximum(n, 1e-20)).astype(F)


def beziers_to_lines(p0, p1, p2, p3, width0, width1, sub: int) -> Shape:
    """(n,3) control points + (n,) widths -> line Shape (yocto_pbrt.h:1764-1794)."""
    p0, p1, p2, p3 = (np.asarray(p, F) for p in (p0, p1, p2, p3))
    n = len(p0)
    u = (np.arange(sub + 1, dtype=F) / F(sub))[None, :, None]  # (1, sub+1, 1)
    P0, P1, P2, P3 = (p[:, None, :] for p in (p0, p1, p2, p3))
    pos = (P0 * (1 - u) ** 3 + P1 * 3 * u * (1 - u) ** 2 + P2 * 3 * u * u * (1 - u) + P3 * u ** 3).astype(F)
    pos[:, 0], pos[:, sub] = p0, p3
    der = ((P1 - P0) * 3 * (1 - u) ** 2 + (P2 - P1) * 6 * u * (1 - u) + (P3 - P2) * 3 * u * u).astype(F)
    der[:, 0], der[:, sub] = p1 - p0, p3 - p2
    tan = _normalize(der)
    w0, w1 = np.asarray(width0, F)[:, None], np.asarray(width1, F)[:, None]
    rad = (w0 * (1 - u[..., 0]) + w1 * u[..., 0]).astype(F)
    base = (np.arange(n, dtype=np.int32) * (sub + 1))[:, None] + np.arange(sub, dtype=np.int32)[None, :]
    lines = np.stack([base, base + 1], axis=-1).reshape(-1, 2)
    return Shape(positions=pos.reshape(-1, 3), normals=tan.reshape(-1, 3), radius=rad.reshape(-1), lines=lines)


def _fibres_to_shape(points, width_root, width_tip, sub, rng, jitter) -> Shape:
    """points: (fibres, knots, 3) guide polylines -> Catmull-Rom Beziers -> lines."""
    nf, nk, _ = points.shape
    if jitter > 0:
        points = points + rng.normal(0, jitter, points.shape).astype(F) * np.linspace(0, 1, nk, dtype=F)[None, :, None]
    tang = np.empty_like(points)
    tang[:, 1:-1] = (points[:, 2:] - points[:, :-2]) * F(0.5)
    tang[:, 0] = points[:, 1] - points[:, 0]
    tang[:, -1] = points[:, -1] - points[:, -2]
    p0, p3 = points[:, :-1], points[:, 1:]
    p1, p2 = p0 + tang[:, :-1] / F(3), p3 - tang[:, 1:] / F(3)
    t = np.linspace(0, 1, nk, dtype=F)
    w = (width_root * (1 - t) + width_tip * t).astype(F)
    w0 = np.broadcast_to(w[None, :-1], (nf, nk - 1))
    w1 = np.broadcast_to(w[None, 1:], (nf, nk - 1))
    r = lambda a: a.reshape(-1, 3)  # noqa: E731
    return beziers_to_lines(r(p0), r(p1), r(p2), r(p3), w0.reshape(-1), w1.reshape(-1), sub)


def _scalp_guides(num_fibres, knots, length, rng, curl=None):
    """Guide polylines growing out of a scalp hemisphere (radius 5 around
    (0, 9.5, 0)), bending under gravity and sliding over the head."""
    centre, radius = np.array([0, 9.5, 0], F), F(5.0)
    z = rng.uniform(-0.25, 1.0, num_fibres).astype(F)           # cap above y = centre - 0.25 R
    phi = rng.uniform(0, 2 * np.pi, num_fibres).astype(F)
    rxy = np.sqrt(np.maximum(0, 1 - z * z)).astype(F)
    normal = np.stack([rxy * np.cos(phi), z, rxy * np.sin(phi)], axis=1).astype(F)
    front = (normal[:, 2] > 0.55) & (normal[:, 1] < 0.35)       # keep the face clear
    normal[front, 2] *= F(-1)
    pos = centre + radius * normal
    step = F(length / (knots - 1))
    lengths = rng.uniform(0.75, 1.0, num_fibres).astype(F)[:, None]
    down = np.array([0, -1, 0], F)
    pts = [pos]
    d = normal.copy()
    for k in range(1, knots):
        blend = F(min(1.0, k / (0.35 * knots)))
        d = _normalize(d * (1 - blend * F(0.6)) + down * blend * F(0.6) + normal * F(0.15))
        pos = pos + d * step * lengths
        rel = pos - centre
        dist = np.linalg.norm(rel, axis=1, keepdims=True)
        inside = (dist < radius * F(1.02)) & (rel[:, 1:2] > -radius * F(0.6))
        pos = np.where(inside, centre + rel / np.maximum(dist, 1e-6) * radius * F(1.02), pos).astype(F)
        pts.append(pos)
    pts = np.stack(pts, axis=1).astype(F)
    if curl is not None:
        hr, pitch = curl
        rad = rng.uniform(hr[0], hr[1], (num_fibres, 1)).astype(F)
        pit = rng.uniform(pitch[0], pitch[1], (num_fibres, 1)).astype(F)
        phase = rng.uniform(0, 2 * np.pi, (num_fibres, 1)).astype(F)
        s = (np.arange(knots, dtype=F) * step)[None, :] * lengths
        ang = phase + 2 * np.pi * s / pit
        e1 = _normalize(np.cross(normal, down + F(1e-3)))
        e2 = _normalize(np.cross(normal, e1))
        grow = np.minimum(1.0, s / 1.5).astype(F)               # no curl at the root
        pts = pts + (rad * grow * np.cos(ang))[..., None] * e1[:, None, :] + (rad * grow * np.sin(ang))[..., None] * e2[:, None, :]
    return pts.astype(F)


def make_straight_hair(num_segments=250_000, seed=7, sub=2, beziers_per_fibre=10) -> Shape:
    rng = np.random.default_rng(seed)
    nf = max(1, num_segments // (sub * beziers_per_fibre))
    guides = _scalp_guides(nf, beziers_per_fibre + 1, 11.0, rng)
    return _fibres_to_shape(guides, F(0.02), F(0.005), sub, rng, 0.05)


def make_curly_hair(num_segments=1_000_000, seed=7, sub=4, beziers_per_fibre=20) -> Shape:
    rng = np.random.default_rng(seed)
    nf = max(1, num_segments // (sub * beziers_per_fibre))
    guides = _scalp_guides(nf, beziers_per_fibre + 1, 10.0, rng, curl=((0.3, 0.6), (1.0, 2.0)))
    return _fibres_to_shape(guides, F(0.02), F(0.005), sub, rng, 0.03)


def make_hair_block(num_segments=200_000, seed=7, sub=4, beziers_per_fibre=5) -> Shape:
    """Straight fibres filling the unit cube along +z (stand-in for pbrt's hair block)."""
    rng = np.random.default_rng(seed)
    nf = max(1, num_segments // (sub * beziers_per_fibre))
    xy = rng.uniform(0.02, 0.98, (nf, 2)).astype(F)
    knots = beziers_per_fibre + 1
    z = np.linspace(0, 1, knots, dtype=F)
    lean = rng.normal(0, 0.04, (nf, 2)).astype(F)
    pts = np.empty((nf, knots, 3), F)
    pts[:, :, 0] = xy[:, 0:1] + lean[:, 0:1] * z[None, :] ** 2
    pts[:, :, 1] = xy[:, 1:2] + lean[:, 1:2] * z[None, :] ** 2
    pts[:, :, 2] = z[None, :]
    return _fibres_to_shape(pts, F(0.004), F(0.001), sub, rng, 0.004)


def make_uv_sphere(steps=(64, 32), radius=1.0) -> Shape:
    nu, nv = steps
    u = np.linspace(0, 1, nu + 1, dtype=F)
    v = np.linspace(0, 1, nv + 1, dtype=F)
    uu, vv = np.meshgrid(u, v)
    phi, theta = 2 * np.pi * uu, np.pi * (1 - vv)
    n = np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1).astype(F).reshape(-1, 3)
    idx = lambda i, j: j * (nu + 1) + i  # noqa: E731
    tris = []
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            tris += [(a, b, d), (c, d, b)]  # quads_to_triangles order (yocto_shape.cpp)
    return Shape(positions=n * F(radius), normals=n, triangles=np.array(tris, np.int32))


def make_arealight_quad() -> Shape:
    """tests/*/shapes/arealight.ply: a 6 x 6 quad in the xy plane facing +z."""
    pos = np.array([[-3, -3, 0], [3, -3, 0], [-3, 3, 0], [3, 3, 0]], F)
    nrm = np.tile(np.array([0, 0, 1], F), (4, 1))
    return Shape(positions=pos, normals=nrm, triangles=np.array([[0, 1, 2], [3, 2, 1]], np.int32))


def make_sky_texture(width=512, height=256, sun=(0.35, 0.25), seed=3) -> Texture:
    """Procedural stand-in for tests/*/textures/sky.hdr (lat-long, float RGB):
    horizon-to-zenith gradient, dark ground and a small bright sun."""
    v = (np.arange(height, dtype=F) + F(0.5)) / F(height)
    u = (np.arange(width, dtype=F) + F(0.5)) / F(width)
    uu, vv = np.meshgrid(u, v)
    up = np.cos(np.pi * vv)
    sky = np.stack([0.35 + 0.35 * (1 - up), 0.55 + 0.30 * (1 - up), 1.0 + 0.0 * up], axis=-1)
    sky = np.where(up[..., None] > 0, sky * (0.6 + 0.8 * (1 - up[..., None]) ** 2), np.array([0.12, 0.11, 0.10]))
    du = np.minimum(np.abs(uu - sun[0]), 1 - np.abs(uu - sun[0])) * 2 * np.sin(np.pi * vv)
    d2 = du ** 2 + (vv - sun[1]) ** 2
    sky = sky + np.exp(-d2 / (2 * 0.012 ** 2))[..., None] * np.array([60.0, 52.0, 40.0])
    return Texture(sky.astype(F))


def _hair_camera():
    # tests/straight-hair/straight-hair.json, tests/curly-hair/curly-hair.json
    return Camera.from_lookat((0, 15, 23), (0, 9.5, 0), (0, 1, 0), lens=0.05, aspect=1.0)


HAIR_FRAME = (-1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0)  # object frame of both hair JSONs


def straight_hair_scene(num_segments=250_000, seed=7, sky=True, area_lights=True, eumelanin=1.3, sky_size=(512, 256)):
    """Config C2.  sky=True uses the procedural sky texture (stand-in for sky.hdr)."""
    sc = Scene()
    sc.shapes.append(make_straight_hair(num_segments, seed))
    sc.materials.append(Material(eumelanin=eumelanin))
    sc.objects.append(Object(shape=0, material=0, frame=HAIR_FRAME))
    if area_lights:
        sc.shapes.append(make_arealight_quad())
        sc.materials.append(Material(emission=(20, 20, 20)))
        for eye in ((-5, 20, 10), (5, 20, 10)):
            sc.objects.append(Object(shape=1, material=1, frame=lookat_frame(eye, (0, 9.5, 0), (0, 1, 0), inv_xz=True)))
    if sky:
        sc.textures.append(make_sky_texture(*sky_size))
        sc.environments.append(Environment(emission=(1, 1, 1), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(1, 1, 1)))
    return sc, _hair_camera()


def curly_hair_scene(num_segments=1_000_000, seed=7, sky=True, sigma_a=(0.06, 0.10, 0.20), sky_size=(512, 256)):
    """Config C3 / C5: blonde curly hair under the sky environment."""
    sc = Scene()
    sc.shapes.append(make_curly_hair(num_segments, seed))
    sc.materials.append(Material(sigma_a=sigma_a))
    sc.objects.append(Object(shape=0, material=0, frame=HAIR_FRAME))
    if sky:
        sc.textures.append(make_sky_texture(*sky_size))
        sc.environments.append(Environment(emission=(2.5, 2.5, 2.5), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(2.5, 2.5, 2.5)))
    return sc, _hair_camera()


def hair_block_scene(num_segments=200_000, seed=7, aspect=1.777777778, **hair):
    """Config C1 / C4: tests/sphere-hairblock/sphere-hairblock.json with a
    synthetic hair block; `hair` overrides material fields (beta_m, beta_n, alpha...)."""
    sc = Scene()
    sc.shapes.append(make_uv_sphere())
    sc.shapes.append(make_hair_block(num_segments, seed))
    sc.materials.append(Material(color=(0.8, 0.4, 0.05)))
    hair_mat = dict(color=(0.8, 0.4, 0.05))
    hair_mat.update(hair)
    sc.materials.append(Material(**hair_mat))
    sc.objects.append(Object(shape=1, material=1, frame=(1, 0, 0, 0, 0, 1, 0, -1, 0, 0.5, 1, -0.5)))  # "hairblock"
    sc.objects.append(Object(shape=0, material=0, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, -0.5, 0, 0)))     # "sphere"
    sc.environments.append(Environment(emission=(1, 1, 1)))
    cam = Camera.from_lookat((-0.5, 1.5, 5), (0.25, 0.5, 0), (0, 1, 0), lens=0.05, aspect=aspect)
    return sc, cam


SURFACE_MATERIALS = {
    "coated": dict(color=(0.8, 0.3, 0.1), specular=1, roughness=0.3),                # diffuse + rough dielectric coat
    "coated_smooth": dict(color=(0.2, 0.5, 0.8), specular=1, roughness=0),           # roughness clamps to 0.03^2 (diffuse != 0)
    "gold": dict(color=(1.0, 0.78, 0.34), metallic=1, roughness=0.25),               # rough conductor
    "mirror": dict(color=(0.9, 0.9, 0.9), metallic=1, roughness=0),                  # delta conductor
    "varnish": dict(color=(0, 0, 0), specular=1, roughness=0),                       # delta dielectric reflection only
    "frosted": dict(color=(0.9, 0.95, 1.0), transmission=1, roughness=0.4, thin=True),   # rough thin-sheet transmission
    "sheet": dict(color=(0.7, 0.9, 0.7), transmission=1, specular=1, roughness=0, thin=True),  # delta transmission + reflection
    "mixed": dict(color=(0.6, 0.6, 0.2), specular=0.7, metallic=0.3, transmission=0.4, roughness=0.5, ior=1.33, thin=True),
    "veil": dict(color=(0.5, 0.5, 0.9), specular=1, roughness=0.2, opacity=0.6),      # opacity pass-through in front of lobes
}
# refraction (thin = False with transmission) opens a homogeneous volume behind the surface (has_volume, eval_vsdf)
VOLUME_MATERIALS = {
    "glass": dict(color=(1, 1, 1), specular=1, transmission=1, roughness=0, thin=False),                     # delta refraction, empty volume
    "rough_glass": dict(color=(0.9, 0.95, 1.0), specular=1, transmission=1, roughness=0.15, thin=False, trdepth=0.5),
    "wax": dict(color=(0.8, 0.5, 0.25), specular=1, transmission=1, roughness=0.5, thin=False, scattering=(0.8, 0.5, 0.25),
                scanisotropy=-0.8, trdepth=0.05),                                                            # sloth "body"
    "milk": dict(color=(0.9, 0.9, 0.85), transmission=0.8, roughness=0.3, thin=False, scattering=(0.9, 0.9, 0.9),
                 scanisotropy=0.0, trdepth=0.2, ior=1.3),
    "air": dict(color=(0.5, 0.9, 0.5), transmission=1, roughness=0, thin=False, ior=1.0005, trdepth=0.3),    # |ior - 1| < 1e-3 branch
}


def surface_lobes_scene(num_segments=4_000, seed=7, sky=True, volumes=False):
    """a hair block among spheres and sheets that carry every surface lobe of the reference's
    material model (diffuse, specular, metal, thin transmission, delta forms, opacity; with volumes=True the
    refractive materials and their media instead), a diffuse floor, one area light and the sky."""
    sc = Scene()
    sc.shapes.append(make_hair_block(num_segments, seed))      # 0
    sc.shapes.append(make_uv_sphere((24, 12), radius=0.38))    # 1
    sc.shapes.append(make_arealight_quad())                    # 2 (6 x 6 quad: floor, sheets, light)
    sc.materials.append(Material(eumelanin=1.3))
    sc.objects.append(Object(shape=0, material=0, frame=(1, 0, 0, 0, 0, 1, 0, -1, 0, 0.5, 1, -0.5)))
    sc.materials.append(Material(color=(0.7, 0.7, 0.7)))       # floor: plain diffuse
    sc.objects.append(Object(shape=2, material=1, frame=(1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0)))
    table = VOLUME_MATERIALS if volumes else SURFACE_MATERIALS
    names = list(table)
    for k, name in enumerate(names):
        sc.materials.append(Material(**table[name]))
        m = len(sc.materials) - 1
        if name in ("frosted", "sheet", "veil"):               # upright sheets in front of the block
            s = 0.12
            x = -1.6 + 1.1 * (["frosted", "sheet", "veil"].index(name))
            sc.objects.append(Object(shape=2, material=m, frame=(s, 0, 0, 0, s, 0, 0, 0, s, x, 0.5, 1.2)))
        else:
            j = [n for n in names if n not in ("frosted", "sheet", "veil")].index(name)
            sc.objects.append(Object(shape=1, material=m, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, -2.0 + 0.8 * j, 0.38, 0.3 * (j % 2))))
    sc.materials.append(Material(emission=(20, 20, 20)))
    sc.objects.append(Object(shape=2, material=len(sc.materials) - 1,
                             frame=lookat_frame((3, 6, 5), (0, 0.5, 0), (0, 1, 0), inv_xz=True)))
    if sky:
        sc.textures.append(make_sky_texture())
        sc.environments.append(Environment(emission=(1, 1, 1), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(1, 1, 1)))
    cam = Camera.from_lookat((0.0, 2.2, 6.0), (0.0, 0.5, 0.0), (0, 1, 0), lens=0.05, aspect=1.5)
    return sc, cam


def make_hair_curl(num_segments=40_000, seed=7, sub=4, beziers_per_fibre=10) -> Shape:
    """A hanging lock of helical fibres around the y axis (stand-in for shapes/hair-curl.ply)."""
    rng = np.random.default_rng(seed)
    nf = max(1, num_segments // (sub * beziers_per_fibre))
    knots = beziers_per_fibre + 1
    t = np.linspace(0, 1, knots, dtype=F)[None, :]
    r0 = rng.uniform(0.05, 0.45, (nf, 1)).astype(F)
    ph = rng.uniform(0, 2 * np.pi, (nf, 1)).astype(F)
    turns = rng.uniform(2.0, 3.0, (nf, 1)).astype(F)
    ang = ph + 2 * np.pi * turns * t
    pts = np.stack([r0 * (1 + 0.6 * t) * np.cos(ang), 7.0 - 7.0 * t + 0 * r0, r0 * (1 + 0.6 * t) * np.sin(ang)], axis=-1).astype(F)
    return _fibres_to_shape(pts, F(0.012), F(0.004), sub, rng, 0.01)


def hair_curls_scene(num_segments=40_000, seed=7, sky=True):
    """tests/hair-curls/hair-curls.json: FOUR objects share ONE hair shape with different
    frames and materials (black / red / brown / blonde), two area lights, sky."""
    sc = Scene()
    sc.shapes.append(make_hair_curl(num_segments, seed))
    sc.shapes.append(make_arealight_quad())
    for mat in (dict(eumelanin=8.0), dict(pheomelanin=2.0), dict(eumelanin=1.3), dict(eumelanin=0.3)):
        sc.materials.append(Material(**mat))
    sc.materials.append(Material(emission=(20, 20, 20)))
    for k, x in enumerate((-3.75, -1.25, 1.25, 3.75)):
        sc.objects.append(Object(shape=0, material=k, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, x, 0, 0)))
    for eye, at in (((-8, 10, 5), (-6.5, 3.5, 0)), ((0, 10, 5), (-1.5, 3.5, 0))):
        sc.objects.append(Object(shape=1, material=4, frame=lookat_frame(eye, at, (0, 1, 0), inv_xz=True)))
    if sky:
        sc.textures.append(make_sky_texture())
        sc.environments.append(Environment(emission=(1, 1, 1), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(1, 1, 1)))
    cam = Camera.from_lookat((0, 4, 16), (0, 3.5, 0), (0, 1, 0), lens=0.05, aspect=1.777777778)
    return sc, cam


def random_rays(scene: Scene, n: int, seed=11):
    """Fixed-ray set for hit-id parity: origins inside / around the first line
    shape's bounds, uniformly random directions (incoherent rays)."""
    from .pathtrace import make_rays
    rng = np.random.default_rng(seed)
    obj = scene.objects[0]
    pos = scene.shapes[obj.shape].positions
    fr = np.asarray(obj.frame, F).reshape(4, 3)
    world = pos @ fr[:3] + fr[3]
    lo, hi = world.min(axis=0), world.max(axis=0)
    o = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), (n, 3)).astype(F)
    z = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    d = np.stack([r * np.cos(ph), r * np.sin(ph), z], axis=1).astype(F)
    return make_rays(o, d)


def camera_rays(camera: Camera, width: int, height: int, n: int, seed=5):
    """Primary rays through random film positions (coherent-ish set)."""
    from .pathtrace import make_rays
    rng = np.random.default_rng(seed)
    fr = np.asarray(camera.frame, F).reshape(4, 3)
    uv = rng.uniform(0, 1, (n, 2)).astype(F)
    q = np.stack([camera.film[0] * (0.5 - uv[:, 0]), camera.film[1] * (uv[:, 1] - 0.5), np.full(n, camera.lens)], axis=1).astype(F)
    d = _normalize(-_normalize(q) @ fr[:3])
    return make_rays(np.tile(fr[3], (n, 1)), d)


def make_point_cloud(num_points=20_000, seed=9, radius=(0.004, 0.02), normals=True, texcoords=False) -> Shape:
    """A ball of points with per-vertex radius (the reference's point primitive: yocto_math.h:3405)."""
    rng = np.random.default_rng(seed)
    d = _normalize(rng.normal(0, 1, (num_points, 3)).astype(F))
    pos = (d * rng.uniform(0, 1, (num_points, 1)).astype(F) ** F(1 / 3) * F(0.5)).astype(F) + np.array([0.5, 0.5, 0.5], F)
    rad = rng.uniform(radius[0], radius[1], num_points).astype(F)
    return Shape(positions=pos, normals=d if normals else None, radius=rad,
                 texcoords=rng.uniform(0, 1, (num_points, 2)).astype(F) if texcoords else None,
                 points=np.arange(num_points, dtype=np.int32))


def points_scene(num_points=20_000, num_segments=6_000, seed=9, sky=True, **point_material):
    """SURVEY 8f N3 leftover: a point cloud (surface lobes, normal = -outgoing) beside a hair block on a diffuse
    floor, one area light and the sky.  `point_material` overrides the cloud's material fields."""
    sc = Scene()
    sc.shapes.append(make_point_cloud(num_points, seed))                  # 0
    sc.shapes.append(make_hair_block(num_segments, seed))                 # 1
    sc.shapes.append(make_arealight_quad())                               # 2
    mat = dict(color=(0.7, 0.3, 0.2), specular=0.5, roughness=0.3)
    mat.update(point_material)
    sc.materials.append(Material(**mat))
    sc.materials.append(Material(eumelanin=1.3))
    sc.materials.append(Material(color=(0.7, 0.7, 0.7)))
    sc.materials.append(Material(emission=(20, 20, 20)))
    sc.objects.append(Object(shape=0, material=0, frame=(1.2, 0, 0, 0, 1.2, 0, 0, 0, 1.2, -1.6, 0.05, -0.3)))
    sc.objects.append(Object(shape=1, material=1, frame=(1, 0, 0, 0, 0, 1, 0, -1, 0, 0.4, 1, -0.5)))
    sc.objects.append(Object(shape=2, material=2, frame=(1, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0)))
    sc.objects.append(Object(shape=2, material=3, frame=lookat_frame((3, 6, 5), (0, 0.5, 0), (0, 1, 0), inv_xz=True)))
    if sky:
        sc.textures.append(make_sky_texture())
        sc.environments.append(Environment(emission=(1, 1, 1), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(1, 1, 1)))
    cam = Camera.from_lookat((0.0, 1.6, 5.0), (0.0, 0.6, 0.0), (0, 1, 0), lens=0.05, aspect=1.5)
    return sc, cam


def instanced_curls_scene(side=(10, 10, 10), num_segments=1_500, coincident=3, seed=7, sky=True):
    """Instance-heavy scene (sloth / bold-man scale in object count): side[0] x side[1] x side[2] objects share ONE hair
    lock on an exact lattice - centres coincide with the split planes of the reference's top-level tree, the case in
    which its visiting order decides - and `coincident` of them are doubled by objects with IDENTICAL frames but other
    materials, so every hit on those is an exact-t tie BETWEEN objects (the reference keeps the one its scene tree
    visits last, yocto_pathtrace.cpp:1003-1011)."""
    sc = Scene()
    sc.shapes.append(make_hair_curl(num_segments, seed))
    sc.shapes.append(make_arealight_quad())
    for mat in (dict(eumelanin=8.0), dict(pheomelanin=2.0), dict(eumelanin=1.3), dict(eumelanin=0.3)):
        sc.materials.append(Material(**mat))
    sc.materials.append(Material(emission=(20, 20, 20)))
    k = 0
    frames = []
    for ix in range(side[0]):
        for iy in range(side[1]):
            for iz in range(side[2]):
                s = 0.25
                frames.append((s, 0, 0, 0, s, 0, 0, 0, s, 1.0 * ix, 2.0 * iy, 1.0 * iz))
    for f in frames:
        sc.objects.append(Object(shape=0, material=k % 4, frame=f))
        k += 1
    rng = np.random.default_rng(seed)
    for j in rng.choice(len(frames), size=min(coincident, len(frames)), replace=False):
        sc.objects.append(Object(shape=0, material=(sc.objects[int(j)].material + 1) % 4, frame=frames[int(j)]))
    cx, cy, cz = (side[0] - 1) / 2, side[1] - 0.1, (side[2] - 1) / 2
    sc.objects.append(Object(shape=1, material=4, frame=lookat_frame((cx, 2.0 * side[1] + 6, cz + 4), (cx, cy, cz), (0, 1, 0), inv_xz=True)))
    if sky:
        sc.textures.append(make_sky_texture())
        sc.environments.append(Environment(emission=(1, 1, 1), emission_tex=0))
    else:
        sc.environments.append(Environment(emission=(1, 1, 1)))
    cam = Camera.from_lookat((cx + 1.3 * side[0], 1.4 * side[1] + 2, cz + 2.2 * side[2] + 4), (cx, side[1] * 0.9, cz), (0, 1, 0), lens=0.05, aspect=1.5)
    return sc, cam
