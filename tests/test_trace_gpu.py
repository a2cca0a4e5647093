"""CUDA traversal + wavefront integrator vs the CPU checker (SURVEY.md 8a rows
P5-P19, M1-M7; 8d "Image parity").

Integer / index outputs (hit flag, object, element) must be bit-exact; because
the primitive tests restate the reference's arithmetic and exact-t ties are
resolved in the reference's own visiting order, uv / distance and whole renders
(accumulation buffer, sample counts, per-pixel PCG32 state) are bit-identical
too, and the tests assert exactly that.  Converged-image rel-MSE (the north_star
criterion) is asserted on top with its stated tolerance.
"""
import numpy as np
import pytest

from yocto_hair_b200 import _capi as capi
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene, make_rays
from yocto_hair_b200.scene import (Camera, Environment, Material, Object, Scene, Shape, make_params)

pytestmark = pytest.mark.gpu

SCENES = {
    "block": lambda: synth.hair_block_scene(24_000),
    "straight": lambda: synth.straight_hair_scene(20_000),
    "curly": lambda: synth.curly_hair_scene(40_000),
    "straight_uniform_env": lambda: synth.straight_hair_scene(12_000, sky=False, area_lights=False),
    "curls_instanced": lambda: synth.hair_curls_scene(16_000),  # 4 objects share one shape
}


def ray_set(scene, cam, n_random=200_000, n_camera=100_000):
    return np.concatenate([synth.random_rays(scene, n_random), synth.camera_rays(cam, 256, 256, n_camera)])


def assert_hits_identical(hr, hd):
    assert np.array_equal(hr["hit"], hd["hit"])
    h = hr["hit"] == 1
    assert np.array_equal(hr["object"][h], hd["object"][h])
    assert np.array_equal(hr["element"][h], hd["element"][h])
    assert np.array_equal(hr["distance"][h].view(np.uint32), hd["distance"][h].view(np.uint32))
    assert np.array_equal(hr["uv"][h].view(np.uint32), hd["uv"][h].view(np.uint32))


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("max_split", [0, 4])
def test_hit_records_bit_exact(name, max_split, ref):
    scene, cam = SCENES[name]()
    rays = ray_set(scene, cam)
    rs = ref.scene(scene)
    hr = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=8)
    ds = DeviceScene(scene).init_bvh(max_split).init_lights()
    hd = ds.intersect(rays)
    assert 0.05 < hr["hit"].mean() < 0.95  # the set exercises both outcomes
    assert_hits_identical(hr, hd)
    # find_any: any hit is a valid answer, the flag must agree
    ha = ds.intersect(rays[:50_000], find_any=True)
    assert np.array_equal(ha["hit"], hr["hit"][:50_000])
    rs.close(); ds.close()


@pytest.mark.parametrize("name", sorted(SCENES))
def test_render_bit_identical(name, ref):
    scene, cam = SCENES[name]()
    p = make_params(resolution=160)
    rs = ref.scene(scene)
    st_r = rs.state(cam, p)
    st_r.trace_samples(6)
    r_r, a_r, s_r, rng_r = st_r.download()
    ds = DeviceScene(scene).init_bvh().init_lights()
    st_d = ds.init_state(cam, p)
    stats = st_d.trace_samples(6)
    r_d, a_d, s_d = st_d.download()
    assert (st_d.width, st_d.height) == (st_r.width, st_r.height)
    assert stats.samples == st_d.width * st_d.height * 6 and stats.rays >= stats.samples
    assert np.array_equal(s_r, s_d)
    assert np.array_equal(rng_r, st_d.rng())
    assert np.array_equal(a_r.view(np.uint32), a_d.view(np.uint32))
    assert np.array_equal(r_r.view(np.uint32), r_d.view(np.uint32))
    st_r.close(); rs.close(); st_d.close(); ds.close()


@pytest.mark.parametrize("alpha", [0.0, 2.0])
@pytest.mark.parametrize("beta_n", [0.3, 0.6, 0.9])
@pytest.mark.parametrize("beta_m", [0.1, 0.25, 0.6])
def test_bsdf_parameter_sweep_renders_bit_identical(beta_m, beta_n, alpha, ref):
    """BASELINE config C4 (the README's longitudinal / azimuthal roughness and scale-tilt sweep) on the hair
    block, at a size the CPU checker renders in a second: every sweep point is bit-identical."""
    scene, cam = synth.hair_block_scene(6_000, aspect=1.0, beta_m=beta_m, beta_n=beta_n, alpha=alpha)
    p = make_params(resolution=64)
    rs = ref.scene(scene)
    st_r = rs.state(cam, p)
    st_r.trace_samples(3)
    a_r = st_r.download()[1]
    ds = DeviceScene(scene).init_bvh().init_lights()
    st_d = ds.init_state(cam, p)
    st_d.trace_samples(3)
    a_d = st_d.download()[1]
    assert np.array_equal(a_r.view(np.uint32), a_d.view(np.uint32))
    st_r.close(); rs.close(); st_d.close(); ds.close()


def rel_mse(x, y):
    """mean((x - y)^2 / (y^2 + 1e-2)) on linear RGB."""
    x, y = x[..., :3].astype(np.float64), y[..., :3].astype(np.float64)
    return float(np.mean((x - y) ** 2 / (y ** 2 + 1e-2)))


def test_converged_image_rel_mse(ref):
    """GPU render with ANOTHER seed against two reference renders: its error must
    sit at the Monte Carlo noise floor.  Tolerance: relMSE(GPU,
    CPU_A) <= 1.5 * relMSE(CPU_B, CPU_A) and mean luminance within 0.5 %."""
    scene, cam = synth.curly_hair_scene(40_000)
    spp, res = 96, 128
    rs = ref.scene(scene)
    imgs = {}
    for tag, seed in (("A", 961748941), ("B", 20240229)):
        st = rs.state(cam, make_params(resolution=res, seed=seed))
        st.trace_samples(spp)
        imgs[tag] = st.download()[0]
        st.close()
    ds = DeviceScene(scene).init_bvh().init_lights()
    st = ds.init_state(cam, make_params(resolution=res, seed=777))
    st.trace_samples(spp)
    gpu = st.render()
    floor = rel_mse(imgs["B"], imgs["A"])
    err = rel_mse(gpu, imgs["A"])
    lum = lambda im: float((im[..., :3] @ np.array([0.2126, 0.7152, 0.0722])).mean())  # noqa: E731
    print("relMSE gpu-vs-A %.4g, B-vs-A %.4g, mean lum gpu %.5f A %.5f" % (err, floor, lum(gpu), lum(imgs["A"])))
    assert err <= 1.5 * floor
    assert abs(lum(gpu) / lum(imgs["A"]) - 1) <= 5e-3
    st.close(); ds.close(); rs.close()


def test_samples_are_additive_and_shards_tile_the_image():
    """n calls of 1 spp == 1 call of n spp; the union of world=3 shards rendered
    one after the other equals the single-device render bit for bit."""
    scene, cam = synth.curly_hair_scene(16_000)
    p = make_params(resolution=96)
    ds = DeviceScene(scene).init_bvh().init_lights()
    full = ds.init_state(cam, p)
    full.trace_samples(4)
    _, a_full, s_full = full.download()
    step = ds.init_state(cam, p)
    for _ in range(4):
        step.trace_samples(1)
    _, a_step, s_step = step.download()
    assert np.array_equal(a_full.view(np.uint32), a_step.view(np.uint32)) and np.array_equal(s_full, s_step)
    acc = np.zeros_like(a_full)
    cnt = np.zeros_like(s_full)
    for rank in range(3):
        st = ds.init_state(cam, p).set_shard(rank, 3)
        st.trace_samples(4)
        _, a, s = st.download()
        assert not np.any((cnt > 0) & (s > 0))  # disjoint ownership
        acc += a
        cnt += s
        st.close()
    assert np.array_equal(cnt, s_full)
    assert np.array_equal(acc.view(np.uint32), a_full.view(np.uint32))
    full.close(); step.close(); ds.close()


def tiny_scene(lines, positions, radius, normals=None, env=(1, 1, 1)):
    sc = Scene()
    sc.shapes.append(Shape(positions=positions, normals=normals, radius=radius, lines=lines))
    sc.materials.append(Material(sigma_a=(0.5, 0.8, 1.2)))
    sc.objects.append(Object(shape=0, material=0))
    sc.environments.append(Environment(emission=env))
    return sc


def test_edge_cases(ref):
    """<= 4 primitives (the tree is one leaf), a zero-length segment (det == 0
    rejects), segments without tangents (eval_normal falls back to the segment
    direction, :275), axis-parallel and zero-component directions."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 0], [0, 1, 0.5], [0.5, 0.5, 1]], np.float32)
    rad = np.full(6, 0.05, np.float32)
    for lines in ([[0, 1]], [[0, 1], [1, 2], [2, 3]], [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 0]]):
        sc = tiny_scene(np.array(lines, np.int32), pos, rad)
        rng = np.random.default_rng(3)
        o = rng.uniform(-0.5, 1.5, (20000, 3)).astype(np.float32)
        d = rng.normal(size=(20000, 3)).astype(np.float32)
        d[:2000] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 2000)] * rng.choice([-1, 1], (2000, 1))
        d[2000:4000, rng.integers(0, 3)] = 0
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = make_rays(o, d)
        rs = ref.scene(sc)
        ds = DeviceScene(sc).init_bvh().init_lights()
        assert_hits_identical(rs.intersect(rays.view(np.float32).reshape(-1, 8)), ds.intersect(rays))
        cam = Camera.from_lookat((0.5, 0.5, 3), (0.5, 0.5, 0))
        p = make_params(resolution=64)
        st_r, st_d = rs.state(cam, p), ds.init_state(cam, p)
        st_r.trace_samples(3)
        st_d.trace_samples(3)
        assert np.array_equal(st_r.download()[1].view(np.uint32), st_d.download()[1].view(np.uint32))
        st_r.close(); rs.close(); st_d.close(); ds.close()


@pytest.mark.parametrize("n,ratio", [(40, 2.0), (200, 1.25)], ids=["40x2", "200x1.25"])
@pytest.mark.parametrize("quality", [0, 1], ids=["lbvh", "sah"])
def test_deep_tree_uses_the_stack_overflow(quality, n, ratio, ref):
    """Sticks at geometrically spaced positions (ratio 2 over 40 octaves): every split peels off the largest
    one, so the tree is a chain and a ray that starts at the small end pushes one far child per level - the
    traversal stack grows past its 16 shared-memory entries into the local overflow (yh_traverse.cuh,
    yh_stack_overflow_store / _load).  With 200 sticks a pure SAH chain would be deeper than the stack (96):
    the builder halves ranges below depth 64 (kSahMaxDepth).  Hit records must still be the reference's."""
    x = (ratio ** (np.arange(n) - n // 2)).astype(np.float32)
    pos = np.zeros((2 * n, 3), np.float32)
    pos[0::2, 0], pos[1::2, 0] = x, x
    pos[0::2, 1], pos[1::2, 1] = -0.1 * x, 0.1 * x
    rad = np.repeat(0.02 * x, 2).astype(np.float32)
    lines = np.stack([np.arange(0, 2 * n, 2), np.arange(1, 2 * n, 2)], axis=1).astype(np.int32)
    sc = tiny_scene(lines, pos, rad)
    rng = np.random.default_rng(11)
    m = 30_000
    k = rng.integers(0, n, m)
    x0 = (x[k] * rng.uniform(0.55, 0.95, m)).astype(np.float32)
    slope = rng.uniform(-0.15, 0.15, (m, 2)).astype(np.float32)
    o = np.stack([x0, slope[:, 0] * x0, 0.01 * slope[:, 1] * x0], axis=1).astype(np.float32)
    d = np.stack([np.ones(m, np.float32), slope[:, 0], 0.01 * slope[:, 1]], axis=1).astype(np.float32)
    d[m // 2:] *= -1  # half of them walk towards the small end
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = make_rays(o, d)
    rs = ref.scene(sc)
    ds = DeviceScene(sc).init_bvh(0, 0, quality).init_lights()
    hr, hd = rs.intersect(rays.view(np.float32).reshape(-1, 8)), ds.intersect(rays)
    assert 0.2 < hr["hit"].mean() < 0.98
    assert_hits_identical(hr, hd)
    rs.close(); ds.close()


@pytest.mark.parametrize("offset", [(300.0, -200.0, 500.0), (0.0, 4000.0, 0.0)], ids=["far", "very_far_y"])
def test_shape_far_from_its_origin(offset, ref):
    """Object-space coordinates far from the origin compared with the shape's size (|p| / extent up to 4000):
    the 16-bit grid, the folded plane decode and its padding (yh_scene.cu fold_pad_steps) must stay
    conservative where fp32 has few bits left for the shape itself.  Hit records stay bit-exact."""
    sc, cam = synth.curly_hair_scene(6_000, sky=False)
    sh = sc.shapes[0]
    sh.positions = (sh.positions + np.array(offset, np.float32)).astype(np.float32)
    rays = synth.random_rays(sc, 150_000)
    rs = ref.scene(sc)
    ds = DeviceScene(sc).init_bvh().init_lights()
    hr, hd = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=8), ds.intersect(rays)
    assert 0.02 < hr["hit"].mean() < 0.98
    assert_hits_identical(hr, hd)
    rs.close(); ds.close()


def test_empty_shape_and_zero_bounces(ref):
    sc = tiny_scene(np.zeros((0, 2), np.int32), np.zeros((0, 3), np.float32), np.zeros((0,), np.float32))
    ds = DeviceScene(sc).init_bvh().init_lights()
    rays = make_rays(np.zeros((100, 3), np.float32), np.tile(np.array([0, 0, 1], np.float32), (100, 1)))
    assert not ds.intersect(rays)["hit"].any()
    cam = Camera.from_lookat((0, 0, 3), (0, 0, 0))
    st = ds.init_state(cam, make_params(resolution=32))
    st.trace_samples(2)
    r, a, s = st.download()
    assert np.all(s == 2) and np.allclose(r[..., :3], 1.0) and np.all(r[..., 3] == 0)  # sky only, alpha 0
    st.close()
    # bounces = 0: the path loop never runs -> black, samples still counted, rng advanced like the reference
    rs = ref.scene(sc)
    p0 = make_params(resolution=32, bounces=0)
    st_d, st_r = ds.init_state(cam, p0), rs.state(cam, p0)
    st_d.trace_samples(3)
    st_r.trace_samples(3)
    _, a_r, s_r, rng_r = st_r.download()
    _, a_d, s_d = st_d.download()
    assert np.array_equal(a_r, a_d) and np.array_equal(s_r, s_d) and np.array_equal(rng_r, st_d.rng())
    st_d.close(); st_r.close(); rs.close(); ds.close()


def test_errors_are_loud():
    lib = capi.load()
    sc, cam = synth.curly_hair_scene(2_000)
    ds = DeviceScene(sc)
    with pytest.raises(capi.YhError) as e:  # trace before init_bvh
        ds.init_state(cam, make_params(resolution=32))
    assert e.value.status == capi.YH_ERR_STATE
    ds.init_bvh().init_lights()
    with pytest.raises(capi.YhError) as e:  # get_trace_shader_func throws "sampler unknown" (:1669)
        ds.init_state(cam, make_params(resolution=32, shader=7))
    assert e.value.status == capi.YH_ERR_INVALID
    ds.close()
    sc.materials[0].specular = 0.5  # microfacet lobes are out of scope: refuse, do not approximate
    with pytest.raises(capi.YhError) as e:
        DeviceScene(sc)
    assert e.value.status == capi.YH_ERR_UNSUPPORTED
    sc.materials[0].specular = 0
    sc.environments.clear()  # the reference's sample_lights is undefined without lights
    ds = DeviceScene(sc).init_bvh().init_lights()
    with pytest.raises(capi.YhError) as e:
        ds.init_state(cam, make_params(resolution=32))
    assert e.value.status == capi.YH_ERR_INVALID
    assert lib.yh_last_error()


@pytest.mark.parametrize("shader", [0, 2, 3])
def test_scene_without_lights_renders_with_the_other_shaders(shader, ref):
    """Only trace_path reads the lights: naive / eyelight / normal render an emitter-less hair model as the reference
    does (eyelight is its usual preview)."""
    sc, cam = synth.curly_hair_scene(8_000, sky=False)
    sc.environments.clear()
    p = make_params(resolution=96, shader=shader)
    rs = ref.scene(sc)
    sr = rs.state(cam, p)
    sr.trace_samples(2)
    ds = DeviceScene(sc).init_bvh().init_lights()
    sd = ds.init_state(cam, p)
    sd.trace_samples(2)
    assert np.array_equal(sr.download()[1].view(np.uint32), sd.download()[1].view(np.uint32))
    if shader != 0:
        assert sd.download()[0][..., :3].max() > 0
    sd.close(); ds.close(); sr.close(); rs.close()


def test_stop_request_leaves_whole_samples(ref):
    """A raised stop flag: no pixel starts another sample, the samples in flight finish, the state stays on the
    reference's per-pixel streams and can be traced further."""
    import ctypes as C
    sc, cam = synth.curly_hair_scene(20_000)
    p = make_params(resolution=128)
    ds = DeviceScene(sc).init_bvh().init_lights()
    sd = ds.init_state(cam, p)
    flag = C.c_int(1)  # raised before the call: the first wavefront round is already queued when the flag is read, so
    capi.check(sd.lib.yh_trace_samples(sd.handle, 50, C.addressof(flag)))  # a pixel finishes one or two samples, not 50
    st = sd.stats()
    _, a1, s1 = sd.download()
    assert st.samples == int(s1.sum()) and s1.min() >= 1 and s1.max() <= 3
    rs = ref.scene(sc)
    sr = rs.state(cam, p)
    ref_accum = {}
    for k in range(1, 7):
        sr.trace_samples(1)
        ref_accum[k] = sr.download()[1]

    def equals_reference(accum, count):  # every pixel holds exactly the first count[pixel] samples of its stream
        for k in np.unique(count):
            m = count == k
            if not np.array_equal(ref_accum[int(k)][m].view(np.uint32), accum[m].view(np.uint32)):
                return False
        return True

    assert equals_reference(a1, s1)
    sd.trace_samples(2)  # and the state continues on the reference's streams
    _, a3, s3 = sd.download()
    assert np.array_equal(s3, s1 + 2) and equals_reference(a3, s3)
    sd.close(); ds.close(); sr.close(); rs.close()


@pytest.mark.parametrize("shader", [1, 2, 0])
def test_point_shapes_bit_identical(shader, ref):
    """Point primitives: bounds :715-721, intersect_point (yocto_math.h:3405) in the leaf loop :894-903,
    shaded with the surface lobes and the normal -outgoing (:364).  Hit records and renders equal the reference's."""
    scene, cam = synth.points_scene()
    rs = ref.scene(scene)
    ds = DeviceScene(scene).init_bvh().init_lights()
    if shader == 1:
        rays = ray_set(scene, cam, 150_000, 100_000)
        hr, hd = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=8), ds.intersect(rays)
        assert (hr["object"][hr["hit"] == 1] == 0).mean() > 0.05  # the cloud is hit
        assert_hits_identical(hr, hd)
    p = make_params(resolution=144, shader=shader)
    sr, sd = rs.state(cam, p), ds.init_state(cam, p)
    sr.trace_samples(3), sd.trace_samples(3)
    _, a_r, s_r, rng_r = sr.download()
    _, a_d, s_d = sd.download()
    assert np.array_equal(s_r, s_d) and np.array_equal(a_r.view(np.uint32), a_d.view(np.uint32))
    assert np.array_equal(rng_r, sd.rng())
    sd.close(); ds.close(); sr.close(); rs.close()


def test_point_shape_variants_bit_identical(ref):
    """One-point shape (root leaf), points without normals, with texcoords + a colour texture, a delta (mirror) cloud."""
    from yocto_hair_b200.scene import Texture
    for variant in ("single", "no_normals", "textured", "mirror"):
        kw = dict(color=(0.9, 0.9, 0.9), metallic=1, roughness=0) if variant == "mirror" else {}
        scene, cam = synth.points_scene(3_000, 1_000, **kw)
        if variant == "single":
            scene.shapes[0] = Shape(positions=[[0.5, 0.5, 0.5]], radius=[0.4], points=[0])
        elif variant == "no_normals":
            scene.shapes[0] = synth.make_point_cloud(3_000, normals=False)
        elif variant == "textured":
            scene.shapes[0] = synth.make_point_cloud(3_000, texcoords=True)
            rng = np.random.default_rng(1)
            scene.textures.append(Texture(rng.uniform(0.1, 1, (16, 16, 3)).astype(np.float32)))
            scene.materials[0].color_tex = len(scene.textures) - 1
        p = make_params(resolution=96)
        rs, ds = ref.scene(scene), DeviceScene(scene).init_bvh().init_lights()
        rays = ray_set(scene, cam, 20_000, 20_000)
        assert_hits_identical(rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=8), ds.intersect(rays))
        sr, sd = rs.state(cam, p), ds.init_state(cam, p)
        sr.trace_samples(2), sd.trace_samples(2)
        assert np.array_equal(sr.download()[1].view(np.uint32), sd.download()[1].view(np.uint32)), variant
        sd.close(); ds.close(); sr.close(); rs.close()


@pytest.mark.parametrize("side,coincident", [((2, 2, 2), 3), ((3, 2, 1), 6), ((10, 10, 10), 12)])
def test_many_instances_follow_the_reference_scene_tree(side, coincident, ref):
    """The top-level tree over the objects is the reference's own (build_bvh over the object bounds :793-814, walked as
    :960-1028): more than four objects, lattice positions that sit on its split planes, and coincident instances whose
    hits tie exactly BETWEEN objects - the hit object must be the one the reference's visiting order keeps.  The largest
    case has 1013 objects on one hair shape."""
    scene, cam = synth.instanced_curls_scene(side, coincident=coincident)
    assert len(scene.objects) > 4
    rs = ref.scene(scene)
    ds = DeviceScene(scene).init_bvh().init_lights()
    rays = ray_set(scene, cam, 60_000, 120_000)
    # rays from inside the lattice, towards everywhere
    rng = np.random.default_rng(5)
    o = rng.uniform((-0.5, 0, -0.5), (side[0] - 0.5, 2.0 * side[1], side[2] - 0.5), (80_000, 3)).astype(np.float32)
    d = rng.normal(0, 1, (80_000, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([rays, make_rays(o, d)])
    hr, hd = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=16), ds.intersect(rays)
    assert hr["hit"].mean() > 0.02
    doubled = set(range(side[0] * side[1] * side[2], len(scene.objects) - 1))
    assert np.isin(hr["object"][hr["hit"] == 1], list(doubled)).any() or coincident == 0  # ties between objects occur
    assert_hits_identical(hr, hd)
    ha = ds.intersect(rays[:40_000], find_any=True)
    assert np.array_equal(ha["hit"], hr["hit"][:40_000])
    p = make_params(resolution=120)
    sr, sd = rs.state(cam, p), ds.init_state(cam, p)
    sr.trace_samples(2), sd.trace_samples(2)
    assert np.array_equal(sr.download()[1].view(np.uint32), sd.download()[1].view(np.uint32))
    assert np.array_equal(sr.download()[3], sd.rng())
    sd.close(); ds.close(); sr.close(); rs.close()


@pytest.mark.parametrize("capacity", ["0", "41", ""])
def test_rays_carried_over_between_launches(capacity, ref, monkeypatch):
    """The traversal hands the long rays of a launch's tail to the next launch (suspend records, yh_traverse.cuh): with
    the record buffer off, too small (reservations that do not fit are skipped, the rays finish in place) or at its
    default size the render is the reference's, bit for bit, and the same number of rays is traced."""
    if capacity:
        monkeypatch.setenv("YH_SUSPEND_CAP", capacity)
    else:
        monkeypatch.delenv("YH_SUSPEND_CAP", raising=False)
    scene, cam = synth.curly_hair_scene(60_000)
    p = make_params(resolution=192)
    rs = ref.scene(scene)
    sr = rs.state(cam, p)
    sr.trace_samples(3)
    ds = DeviceScene(scene).init_bvh().init_lights()
    sd = ds.init_state(cam, p)
    stats = sd.trace_samples(3)
    _, a_r, s_r, rng_r = sr.download()
    _, a_d, s_d = sd.download()
    assert np.array_equal(s_r, s_d) and np.array_equal(a_r.view(np.uint32), a_d.view(np.uint32)) and np.array_equal(rng_r, sd.rng())
    test_rays_carried_over_between_launches.rays = getattr(test_rays_carried_over_between_launches, "rays", stats.rays)
    assert stats.rays == test_rays_carried_over_between_launches.rays
    sd.close(); ds.close(); sr.close(); rs.close()


def test_full_size_properties(ref):
    """BASELINE configs[2] geometry (1M curly segments): oracle agreement on a
    ray subset, and shard-union == whole image at the bench resolution's tile grid."""
    scene, cam = synth.curly_hair_scene(1_000_000)
    ds = DeviceScene(scene).init_bvh().init_lights()
    bs = ds.bvh_stats()
    assert bs.num_primitives == 1_000_000 and bs.num_references >= bs.num_primitives
    rays = ray_set(scene, cam, 60_000, 60_000)
    rs = ref.scene(scene)
    assert_hits_identical(rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=16), ds.intersect(rays))
    rs.close()
    p = make_params(resolution=320)
    full = ds.init_state(cam, p)
    full.trace_samples(2)
    _, a_full, s_full = full.download()
    acc, cnt = np.zeros_like(a_full), np.zeros_like(s_full)
    for rank in range(8):
        st = ds.init_state(cam, p).set_shard(rank, 8)
        st.trace_samples(2)
        _, a, s = st.download()
        acc += a
        cnt += s
        st.close()
    assert np.array_equal(cnt, s_full) and np.array_equal(acc.view(np.uint32), a_full.view(np.uint32))
    full.close(); ds.close()


def test_counter_mode_bit_identical_to_port_and_shards_add_up(port, ref):
    """yh_state_set_sample_shard: per-sample streams make_rng(seed, seq_pixel + (s << 31)).
    world = 1 must equal the port's restatement of the same rule bit for bit (and the
    reference itself at 1 spp); the sample shards of world = 4 must add up to it."""
    scene, cam = synth.curly_hair_scene(30_000)
    p = make_params(resolution=128)
    ds = DeviceScene(scene).init_bvh().init_lights()
    rp = port.scene(scene)
    sd, sp = ds.init_state(cam, p).set_sample_shard(0, 1), rp.state(cam, p).set_sample_shard(0, 1)
    sd.trace_samples(1), sp.trace_samples(1)
    if ref.kind == "reference":
        rr = ref.scene(scene)
        sr = rr.state(cam, p)
        sr.trace_samples(1)
        assert np.array_equal(sr.download()[1].view(np.uint32), sd.download()[1].view(np.uint32))
        sr.close(); rr.close()
    sd.trace_samples(7), sp.trace_samples(7)
    _, a_d, s_d = sd.download()
    _, a_p, s_p, rng_p = sp.download()
    assert np.array_equal(s_d, s_p) and (s_d == 8).all()
    assert np.array_equal(a_d.view(np.uint32), a_p.view(np.uint32))
    assert np.array_equal(sd.rng(), rng_p)
    acc, cnt = np.zeros_like(a_d, dtype=np.float64), np.zeros_like(s_d)
    for rank in range(4):
        st = ds.init_state(cam, p).set_sample_shard(rank, 4)
        st.trace_samples(2)
        _, a, s = st.download()
        acc += a
        cnt += s
        # shard `rank` alone == the port's shard `rank`
        spr = rp.state(cam, p).set_sample_shard(rank, 4)
        spr.trace_samples(2)
        assert np.array_equal(a.view(np.uint32), spr.download()[1].view(np.uint32))
        spr.close(); st.close()
    assert np.array_equal(cnt, s_d)
    assert np.allclose(acc, a_d, rtol=2e-6, atol=1e-6)
    sd.close(); sp.close(); rp.close(); ds.close()


def test_counter_mode_with_several_slots_per_pixel():
    """yh_state_set_sample_slots: S lanes per pixel trace the sample ordinals k, k + S, ... concurrently.  The image holds
    the same samples as the one-slot counter image (equal up to fp32 summation order), every pixel gets exactly the
    requested count also when it is not a multiple of S, and the result is deterministic."""
    scene, cam = synth.straight_hair_scene(20_000)
    p = make_params(resolution=128)
    ds = DeviceScene(scene).init_bvh().init_lights()
    one = ds.init_state(cam, p).set_sample_shard(0, 1)
    one.trace_samples(6)
    _, a_one, s_one = one.download()
    for slots in (2, 4, 0):
        imgs = []
        for rep in range(2):
            st = ds.init_state(cam, p).set_sample_shard(0, 1).set_sample_slots(slots)
            stats = st.trace_samples(6)
            assert stats.samples == st.width * st.height * 6
            _, a, s = st.download()
            assert (s == 6).all()
            imgs.append(a)
            st.close()
        assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32))  # deterministic
        assert np.allclose(imgs[0], a_one, rtol=2e-6, atol=1e-6)
    # two calls hand out the ordinals 0..2 and 3..5: the same six samples
    st = ds.init_state(cam, p).set_sample_shard(0, 1).set_sample_slots(4)
    st.trace_samples(3), st.trace_samples(3)
    _, a, s = st.download()
    assert (s == 6).all() and np.allclose(a, a_one, rtol=2e-6, atol=1e-6)
    st.close()
    # sample shards of a world of 2, four lanes each, add up to the same image
    acc = np.zeros_like(a_one, dtype=np.float64)
    for rank in range(2):
        st = ds.init_state(cam, p).set_sample_shard(rank, 2).set_sample_slots(4)
        st.trace_samples(3)
        acc += st.download()[1]
        st.close()
    assert np.allclose(acc, a_one, rtol=2e-6, atol=1e-6)
    with pytest.raises(capi.YhError) as e:  # the reference's one-stream-per-pixel mode has nothing to run side by side
        ds.init_state(cam, p).set_sample_slots(2)
    assert e.value.status == capi.YH_ERR_STATE
    one.close(); ds.close()


def test_counter_mode_converges_to_the_reference_image(ref):
    """Same rel-MSE criterion as test_converged_image_rel_mse, for counter-mode streams."""
    scene, cam = synth.curly_hair_scene(40_000)
    spp, res = 96, 128
    rs = ref.scene(scene)
    imgs = {}
    for tag, seed in (("A", 961748941), ("B", 20240229)):
        st = rs.state(cam, make_params(resolution=res, seed=seed))
        st.trace_samples(spp)
        imgs[tag] = st.download()[0]
        st.close()
    ds = DeviceScene(scene).init_bvh().init_lights()
    st = ds.init_state(cam, make_params(resolution=res, seed=4242)).set_sample_shard(0, 1)
    st.trace_samples(spp)
    gpu = st.render()
    floor, err = rel_mse(imgs["B"], imgs["A"]), rel_mse(gpu, imgs["A"])
    lum = lambda im: float((im[..., :3] @ np.array([0.2126, 0.7152, 0.0722])).mean())  # noqa: E731
    print("counter mode: relMSE gpu-vs-A %.4g, B-vs-A %.4g, mean lum gpu %.5f A %.5f" % (err, floor, lum(gpu), lum(imgs["A"])))
    assert err <= 1.5 * floor
    assert abs(lum(gpu) / lum(imgs["A"]) - 1) <= 5e-3
    st.close(); ds.close(); rs.close()


@pytest.mark.parametrize("shader", [0, 2, 3], ids=["naive", "eyelight", "normal"])
@pytest.mark.parametrize("name", ["block", "straight", "curls_instanced"])
def test_other_shaders_bit_identical(name, shader, ref):
    """trace_naive / trace_eyelight / trace_normal (yocto_pathtrace.cpp:1514-1657)."""
    scene, cam = SCENES[name]()
    for bounces in (8, 0):
        p = make_params(resolution=96, shader=shader, bounces=bounces)
        rs = ref.scene(scene)
        st_r = rs.state(cam, p)
        st_r.trace_samples(3)
        _, a_r, s_r, rng_r = st_r.download()
        ds = DeviceScene(scene).init_bvh().init_lights()
        st_d = ds.init_state(cam, p)
        st_d.trace_samples(3)
        _, a_d, s_d = st_d.download()
        assert np.array_equal(s_r, s_d) and np.array_equal(rng_r, st_d.rng())
        assert np.array_equal(a_r.view(np.uint32), a_d.view(np.uint32))
        st_r.close(); rs.close(); st_d.close(); ds.close()


@pytest.mark.parametrize("shader", [1, 0, 2], ids=["path", "naive", "eyelight"])
def test_surface_lobes_bit_identical(shader, ref):
    """specular / metal / thin-transmission lobes, their delta forms and opacity on the
    triangles around the hair (yh_surface.cuh) against the reference build: renders bit-identical, all
    three shaders that evaluate materials."""
    scene, cam = synth.surface_lobes_scene()
    p = make_params(resolution=144, shader=shader)
    rs = ref.scene(scene)
    st_r = rs.state(cam, p)
    st_r.trace_samples(4)
    _, a_r, s_r, rng_r = st_r.download()
    ds = DeviceScene(scene).init_bvh().init_lights()
    st_d = ds.init_state(cam, p)
    st_d.trace_samples(4)
    _, a_d, s_d = st_d.download()
    assert np.array_equal(s_r, s_d) and np.array_equal(rng_r, st_d.rng())
    diff = np.nonzero((a_r.view(np.uint32) != a_d.view(np.uint32)).any(axis=-1))
    assert diff[0].size == 0, "first differing pixels (y, x): %s" % list(zip(diff[0][:8], diff[1][:8]))
    assert np.isfinite(a_d).all() and a_d[..., :3].max() > 0
    st_r.close(); rs.close(); st_d.close(); ds.close()


@pytest.mark.parametrize("shader", [1, 0, 2], ids=["path", "naive", "eyelight"])
def test_refraction_and_volumes_bit_identical(shader, ref):
    """Refractive materials (thin = False with transmission; delta and rough, the |ior - 1| < 1e-3 branch) and the
    homogeneous volume trace_path opens behind them: transmittance sampling, Henyey-Greenstein scattering events
    with light MIS, the one-entry volume stack (yocto_pathtrace.cpp:1403-1414, :1458-1497).  The hair block stands
    among the spheres, so hair vertices are also reached from inside media."""
    scene, cam = synth.surface_lobes_scene(volumes=True)
    cam2 = Camera.from_lookat((0.4, 0.38, 0.05), (2.0, 0.6, 0.0), (0, 1, 0), lens=0.03, aspect=1.5)  # starts INSIDE the wax sphere
    for c, res in ((cam, 144), (cam2, 72)):
        p = make_params(resolution=res, shader=shader)
        rs = ref.scene(scene)
        st_r = rs.state(c, p)
        st_r.trace_samples(4)
        _, a_r, s_r, rng_r = st_r.download()
        ds = DeviceScene(scene).init_bvh().init_lights()
        st_d = ds.init_state(c, p)
        st_d.trace_samples(4)
        _, a_d, s_d = st_d.download()
        assert np.array_equal(s_r, s_d) and np.array_equal(rng_r, st_d.rng())
        diff = np.nonzero((a_r.view(np.uint32) != a_d.view(np.uint32)).any(axis=-1))
        assert diff[0].size == 0, "first differing pixels (y, x): %s" % list(zip(diff[0][:8], diff[1][:8]))
        assert np.isfinite(a_d).all() and a_d[..., :3].max() > 0
        st_r.close(); rs.close(); st_d.close(); ds.close()


@pytest.mark.parametrize("shader", [1, 0, 2], ids=["path", "naive", "eyelight"])
def test_material_textures_bit_identical(shader, ref):
    """Float RGB textures in every material slot the reference's tracer evaluates (emission, color, specular,
    metallic, roughness, opacity, scattering, and the transmission factor it reads from the EMISSION texture,
    yocto_pathtrace.cpp:397-424, :505-526), with vertex texcoords (tiled, negative) on the quads and the
    element-uv fallback on the spheres (eval_texcoord, :294-309)."""
    from yocto_hair_b200.scene import Texture
    scene, cam = synth.surface_lobes_scene(volumes=False)
    vol, _ = synth.surface_lobes_scene(volumes=True)
    rng = np.random.default_rng(5)

    def tex(lo, hi, size=(12, 20)):
        scene.textures.append(Texture(colorf=rng.uniform(lo, hi, (*size, 3)).astype(np.float32)))
        return len(scene.textures) - 1

    quad = scene.shapes[2]
    quad.texcoords = np.array([[-0.4, 0.1], [2.7, 0.1], [-0.4, 3.3], [2.7, 3.3]], np.float32)
    names = ["hair", "floor"] + list(synth.SURFACE_MATERIALS) + ["light"]
    m = dict(zip(names, scene.materials))
    m["floor"].color_tex = tex(0.2, 1.0)
    m["coated"].color_tex, m["coated"].roughness_tex, m["coated"].specular_tex = tex(0.1, 1.0), tex(0.3, 1.0), tex(0.0, 1.0)
    m["gold"].metallic_tex, m["gold"].roughness_tex = tex(0.0, 1.0), tex(0.0, 1.0)          # roughness 0 texels: delta vertices
    m["frosted"].emission = (0.3, 0.2, 0.1)
    m["frosted"].emission_tex = m["frosted"].transmission_tex = tex(0.0, 1.0)              # emission texture scales transmission
    m["veil"].opacity_tex, m["veil"].color_tex = tex(0.0, 1.0), tex(0.2, 0.9)
    m["light"].emission_tex = tex(0.5, 1.0)
    m["hair"].opacity, m["hair"].opacity_tex = 0.9, tex(0.6, 1.0)                           # textured opacity on the hair itself
    # a scattering sphere with textured albedo and density
    wax = vol.materials[2 + list(synth.VOLUME_MATERIALS).index("wax")]
    wax.scattering_tex, wax.color_tex = tex(0.3, 1.0), tex(0.4, 1.0)
    scene.materials.append(wax)
    scene.objects.append(Object(shape=1, material=len(scene.materials) - 1, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0.4, 0.38, 1.6)))
    # tangent-space normal maps (eval_normalmap, :326-347): on the floor quad (uv tangents) and on a sphere
    # without texcoords (no tangents: the frame degenerates exactly as the reference's does)
    m["floor"].normal_tex = tex(0.25, 0.75)
    m["mirror"].normal_tex = tex(0.4, 0.6)
    p = make_params(resolution=144, shader=shader)
    rs = ref.scene(scene)
    st_r = rs.state(cam, p)
    st_r.trace_samples(4)
    _, a_r, s_r, rng_r = st_r.download()
    ds = DeviceScene(scene).init_bvh().init_lights()
    st_d = ds.init_state(cam, p)
    st_d.trace_samples(4)
    _, a_d, s_d = st_d.download()
    assert np.array_equal(s_r, s_d) and np.array_equal(rng_r, st_d.rng())
    diff = np.nonzero((a_r.view(np.uint32) != a_d.view(np.uint32)).any(axis=-1))
    assert diff[0].size == 0, "first differing pixels (y, x): %s" % list(zip(diff[0][:8], diff[1][:8]))
    st_r.close(); rs.close(); st_d.close(); ds.close()
