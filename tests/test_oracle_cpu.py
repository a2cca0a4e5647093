"""Pins the CPU checkers (CPU only, no GPU):

  * the plain-C restatement oracle/yh_oracle.c ("port") against the committed
    golden vectors tests/golden/*.npz, which tests/golden/make_golden.py
    generated from the UNMODIFIED reference build (oracle/_ref/libyh_ref.so);
  * where that reference build is present (this container; it also travels to
    the GPU box), port vs reference BIT FOR BIT on seeded BSDF tuples, hit
    records and small renders, and the reference's own four self-tests.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle as orc
from tuples import make_tuples
from yocto_hair_b200 import hair, synth
from yocto_hair_b200.pathtrace import make_rays
from yocto_hair_b200.scene import Camera, Environment, Material, Object, Scene, Shape, make_params

GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def refo():
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libyh_ref.so not built (needs /root/reference)")
    return orc.RefOracle()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def test_port_rng_matches_golden(port):
    g = np.load(GOLDEN / "rng_kat.npz")
    for (seed, seq), floats in zip(g["pairs"], g["floats"]):
        assert np.array_equal(bits(port.rng_floats(int(seed), int(seq), floats.size)), bits(floats))
    # make_rng(961748941, 7) -> first 4 rand1f
    assert np.allclose(port.rng_floats(961748941, 7, 4), [0.616754532, 0.163862228, 0.111510277, 0.444151044], rtol=0, atol=1e-9)


def test_port_hair_matches_golden(port):
    g = np.load(GOLDEN / "hair_kat.npz")
    n = len(g["v"])
    mats = hair.make_materials(n, sigma_a=g["sigma_a"], beta_m=g["beta_m"], beta_n=g["beta_n"], alpha=g["alpha"],
                               eta=g["eta"], color=g["color"], eumelanin=g["eumelanin"], pheomelanin=g["pheomelanin"])
    b = port.hair_brdf(mats, g["v"], g["normals"], g["tangents"])
    assert same(hair.brdf_view(b)[:, :27], g["brdf"][:, :27]).all()
    assert same(port.hair_eval(b, g["wo"], g["wi"]), g["f"]).all()
    assert same(port.hair_pdf(b, g["wo"], g["wi"]), g["pdf"]).all()
    ws = port.hair_sample(b, g["wo"], g["rn"])
    assert same(ws, g["ws"]).all()
    assert same(port.hair_eval(b, g["wo"], g["ws"]), g["f_ws"]).all()
    assert same(port.hair_pdf(b, g["wo"], g["ws"]), g["pdf_ws"]).all()
    # the three hand-derived cases of SURVEY.md 8a (values printed there to 9 digits)
    assert np.allclose(g["f"][0], [0.00218053185, 0.00107384764, 0.000194572029], rtol=2e-7)
    assert np.allclose(g["pdf"][:3], [0.00544281909, 0.000484268297, 0.389571518], rtol=2e-7)
    assert np.allclose(g["ws"][2], [-0.319444478, 0.258741528, -0.911596477], rtol=2e-6)


def test_port_hair_bit_exact_vs_reference(port, refo):
    t = make_tuples(refo, 100_000)
    br, bp = (o.hair_brdf(t["materials"], t["v"], t["normals"], t["tangents"]) for o in (refo, port))
    assert same(hair.brdf_view(bp)[:, :27], hair.brdf_view(br)[:, :27]).all()
    assert same(port.hair_eval(br, t["wo"], t["wi"]), refo.hair_eval(br, t["wo"], t["wi"])).all()
    assert same(port.hair_pdf(br, t["wo"], t["wi"]), refo.hair_pdf(br, t["wo"], t["wi"])).all()
    ws = refo.hair_sample(br, t["wo"], t["rn"])
    assert same(port.hair_sample(br, t["wo"], t["rn"]), ws).all()
    ok = np.isfinite(ws).all(axis=1)
    assert same(port.hair_eval(br, t["wo"], ws)[ok], refo.hair_eval(br, t["wo"], ws)[ok]).all()


@pytest.mark.parametrize("which", range(4))
def test_self_tests_pass(which, port, refo):
    """white_furnace / white_furnace_sampled / sampling_weights / sampling_consistency
    (yocto_extension.cpp:555-693) inside the reference build and in the port."""
    assert refo.hair_selftest(which)
    assert port.hair_selftest(which)


SCENES = {
    "block": lambda: synth.hair_block_scene(6_000),
    "straight": lambda: synth.straight_hair_scene(5_000),
    "curly": lambda: synth.curly_hair_scene(8_000),
    "curls_instanced": lambda: synth.hair_curls_scene(6_000),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_port_traversal_and_render_bit_exact_vs_reference(name, port, refo):
    scene, cam = SCENES[name]()
    rays = np.concatenate([synth.random_rays(scene, 30_000), synth.camera_rays(cam, 128, 128, 20_000)])
    rr, rp = refo.scene(scene), port.scene(scene)
    flat = rays.view(np.float32).reshape(-1, 8)
    hr, hp = rr.intersect(flat, threads=4), rp.intersect(flat, threads=4)
    assert 0.02 < hr["hit"].mean() < 0.98
    assert np.array_equal(hr["hit"], hp["hit"])
    h = hr["hit"] == 1
    for f in ("object", "element"):
        assert np.array_equal(hr[f][h], hp[f][h])
    assert np.array_equal(bits(hr["distance"][h]), bits(hp["distance"][h])) and np.array_equal(bits(hr["uv"][h]), bits(hp["uv"][h]))
    assert np.array_equal(rr.intersect(flat[:5000], find_any=True)["hit"], rp.intersect(flat[:5000], find_any=True)["hit"])
    p = make_params(resolution=64)
    sr, sp = rr.state(cam, p), rp.state(cam, p)
    sr.trace_samples(3), sp.trace_samples(3)
    (r1, a1, s1, g1), (r2, a2, s2, g2) = sr.download(), sp.download()
    assert (sr.width, sr.height) == (sp.width, sp.height)
    assert np.array_equal(s1, s2) and np.array_equal(g1, g2)
    assert np.array_equal(bits(a1), bits(a2)) and np.array_equal(bits(r1), bits(r2))
    for x in (sr, sp, rr, rp):
        x.close()


def test_port_edge_cases_vs_reference(port, refo):
    """Tiny trees (one leaf), zero-length segments, no tangents, empty shape."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 0], [0, 1, 0.5], [0.5, 0.5, 1]], np.float32)
    rad = np.full(6, 0.05, np.float32)
    rng = np.random.default_rng(3)
    o = rng.uniform(-0.5, 1.5, (4000, 3)).astype(np.float32)
    d = rng.normal(size=(4000, 3)).astype(np.float32)
    d[:400] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 400)]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flat = make_rays(o, d).view(np.float32).reshape(-1, 8)
    for lines in ([], [[0, 1]], [[0, 1], [1, 2], [2, 3]], [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [5, 0]]):
        sc = Scene()
        sc.shapes.append(Shape(positions=pos, radius=rad, lines=np.array(lines, np.int32).reshape(-1, 2)))
        sc.materials.append(Material(sigma_a=(0.5, 0.8, 1.2)))
        sc.objects.append(Object(shape=0, material=0))
        sc.environments.append(Environment(emission=(1, 1, 1)))
        rr, rp = refo.scene(sc), port.scene(sc)
        hr, hp = rr.intersect(flat), rp.intersect(flat)
        assert np.array_equal(hr["hit"], hp["hit"])
        h = hr["hit"] == 1
        assert np.array_equal(hr["element"][h], hp["element"][h]) and np.array_equal(bits(hr["distance"][h]), bits(hp["distance"][h]))
        cam = Camera.from_lookat((0.5, 0.5, 3), (0.5, 0.5, 0))
        for bounces in (8, 0):
            p = make_params(resolution=32, bounces=bounces)
            sr, sp = rr.state(cam, p), rp.state(cam, p)
            sr.trace_samples(2), sp.trace_samples(2)
            assert np.array_equal(bits(sr.download()[1]), bits(sp.download()[1])) and np.array_equal(sr.download()[3], sp.download()[3])
            sr.close(), sp.close()
        rr.close(), rp.close()


def test_synthetic_generator_follows_the_pbrt_curve_rule():
    """yocto_pbrt.h:1751-1796: `sub` lines over `sub + 1` UNSHARED vertices per Bezier,
    unit tangents in the normal slot, radius interpolated along the curve."""
    for sub, make in ((2, synth.make_straight_hair), (4, synth.make_curly_hair)):
        sh = make(4000)
        nb = len(sh.lines) // sub
        assert len(sh.lines) == nb * sub and len(sh.positions) == nb * (sub + 1)
        assert len(sh.normals) == len(sh.positions) == len(sh.radius)
        assert np.array_equal(sh.lines[:, 1] - sh.lines[:, 0], np.ones(len(sh.lines), np.int32))
        assert not np.isin(sh.lines[:, 0], np.arange(nb) * (sub + 1) + sub).any()  # no line crosses Beziers
        assert np.allclose(np.linalg.norm(sh.normals, axis=1), 1, atol=1e-5)
        assert (sh.radius > 0).all() and len(sh.positions) // (sub + 1) <= 255 * nb


def test_port_counter_mode_properties(port, refo):
    """Counter mode (the library's sample-sharded stream rule, restated in the port):
    sample 0 equals the reference's first sample bit for bit; the union of the
    sample shards of world = 3 holds the same per-pixel sample set as world = 1."""
    scene, cam = synth.curly_hair_scene(4_000)
    p = make_params(resolution=48)
    rr, rp = refo.scene(scene), port.scene(scene)
    sr = rr.state(cam, p)
    sr.trace_samples(1)
    sp = rp.state(cam, p).set_sample_shard(0, 1)
    sp.trace_samples(1)
    assert np.array_equal(bits(sr.download()[1]), bits(sp.download()[1]))
    sp.trace_samples(5)
    _, a_full, s_full, _ = sp.download()
    acc = np.zeros_like(a_full, dtype=np.float64)
    cnt = np.zeros_like(s_full)
    for rank in range(3):
        st = rp.state(cam, p).set_sample_shard(rank, 3)
        st.trace_samples(2)
        _, a, s, _ = st.download()
        acc += a
        cnt += s
        st.close()
    assert np.array_equal(cnt, s_full) and (cnt == 6).all()
    assert np.allclose(acc, a_full, rtol=2e-6, atol=1e-6)  # same samples, another summation order
    # and the counter-mode estimate agrees with the reference's sequential-stream estimate
    sr.trace_samples(31), sp.trace_samples(26)
    m_ref, m_cnt = sr.download()[0][..., :3].mean(), sp.download()[0][..., :3].mean()
    assert abs(m_cnt / m_ref - 1) < 0.03
    for x in (sr, sp, rr, rp):
        x.close()


@pytest.mark.parametrize("shader", [0, 2, 3], ids=["naive", "eyelight", "normal"])
def test_port_other_shaders_bit_exact_vs_reference(shader, port, refo):
    """trace_naive / trace_eyelight / trace_normal restated in the port (yocto_pathtrace.cpp:1514-1657)."""
    for make in (lambda: synth.hair_block_scene(3_000), lambda: synth.straight_hair_scene(3_000)):
        scene, cam = make()
        for bounces in (8, 0):
            p = make_params(resolution=48, shader=shader, bounces=bounces)
            rr, rp = refo.scene(scene), port.scene(scene)
            sr, sp = rr.state(cam, p), rp.state(cam, p)
            sr.trace_samples(2), sp.trace_samples(2)
            (_, a1, s1, g1), (_, a2, s2, g2) = sr.download(), sp.download()
            assert np.array_equal(bits(a1), bits(a2)) and np.array_equal(s1, s2) and np.array_equal(g1, g2)
            for x in (sr, sp, rr, rp):
                x.close()


def textured_materials_scene():
    """surface_lobes_scene with a float texture in every material slot the tracer reads, vertex texcoords on the
    quads, normal maps, and a scattering sphere with textured albedo / density (same construction as
    tests/test_trace_gpu.py::test_material_textures_bit_identical)."""
    from yocto_hair_b200.scene import Object, Texture
    scene, cam = synth.surface_lobes_scene(1_500, volumes=False)
    vol, _ = synth.surface_lobes_scene(500, volumes=True)
    rng = np.random.default_rng(5)

    def tex(lo, hi, size=(12, 20)):
        scene.textures.append(Texture(colorf=rng.uniform(lo, hi, (*size, 3)).astype(np.float32)))
        return len(scene.textures) - 1

    scene.shapes[2].texcoords = np.array([[-0.4, 0.1], [2.7, 0.1], [-0.4, 3.3], [2.7, 3.3]], np.float32)
    m = dict(zip(["hair", "floor"] + list(synth.SURFACE_MATERIALS) + ["light"], scene.materials))
    m["floor"].color_tex, m["floor"].normal_tex = tex(0.2, 1.0), tex(0.25, 0.75)
    m["coated"].color_tex, m["coated"].roughness_tex, m["coated"].specular_tex = tex(0.1, 1.0), tex(0.3, 1.0), tex(0.0, 1.0)
    m["gold"].metallic_tex, m["gold"].roughness_tex = tex(0.0, 1.0), tex(0.0, 1.0)
    m["mirror"].normal_tex = tex(0.4, 0.6)
    m["frosted"].emission = (0.3, 0.2, 0.1)
    m["frosted"].emission_tex = m["frosted"].transmission_tex = tex(0.0, 1.0)
    m["veil"].opacity_tex, m["veil"].color_tex = tex(0.0, 1.0), tex(0.2, 0.9)
    m["light"].emission_tex = tex(0.5, 1.0)
    m["hair"].opacity, m["hair"].opacity_tex = 0.9, tex(0.6, 1.0)
    wax = vol.materials[2 + list(synth.VOLUME_MATERIALS).index("wax")]
    wax.scattering_tex, wax.color_tex = tex(0.3, 1.0), tex(0.4, 1.0)
    scene.materials.append(wax)
    scene.objects.append(Object(shape=1, material=len(scene.materials) - 1, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0.4, 0.38, 1.6)))
    return scene, cam


MATERIAL_SCENES = {
    "lobes": lambda: synth.surface_lobes_scene(1_500),
    "refraction_volumes": lambda: synth.surface_lobes_scene(1_500, volumes=True),
    "textures": textured_materials_scene,
}


@pytest.mark.parametrize("shader", [1, 0, 2], ids=["path", "naive", "eyelight"])
@pytest.mark.parametrize("name", sorted(MATERIAL_SCENES))
def test_port_material_model_bit_exact_vs_reference(name, shader, port, refo):
    """SURVEY 8f N3 in the plain-C restatement: surface lobes and their delta forms, refraction, homogeneous
    volumes, material textures and normal maps (yocto_pathtrace.cpp:294-531, :1069-1280, :1360-1497;
    yocto_math.h:4215-4812) - renders bit-identical to the reference build, RNG state included."""
    scene, cam = MATERIAL_SCENES[name]()
    p = make_params(resolution=96, shader=shader)
    rr, rp = refo.scene(scene), port.scene(scene)
    sr, sp = rr.state(cam, p), rp.state(cam, p)
    sr.trace_samples(3), sp.trace_samples(3)
    (_, a1, s1, g1), (_, a2, s2, g2) = sr.download(), sp.download()
    assert np.array_equal(s1, s2) and np.array_equal(g1, g2)
    diff = np.nonzero((bits(a1) != bits(a2)).any(axis=-1))
    assert diff[0].size == 0, "first differing pixels (y, x): %s" % list(zip(diff[0][:8], diff[1][:8]))
    for x in (sr, sp, rr, rp):
        x.close()


@pytest.mark.parametrize("shader", [1, 0, 2, 3], ids=["path", "naive", "eyelight", "normal"])
def test_port_point_shapes_bit_exact_vs_reference(shader, port, refo):
    """Point primitives in the plain-C restatement (point_bounds yocto_math.h:3030, intersect_point :3405, the leaf
    loop yocto_pathtrace.cpp:894-903, shading normal -outgoing :364): hit records and renders equal the reference's."""
    scene, cam = synth.points_scene(3_000, 1_000)
    rr, rp = refo.scene(scene), port.scene(scene)
    if shader == 1:
        rays = np.concatenate([synth.random_rays(scene, 20_000), synth.camera_rays(cam, 256, 256, 20_000)])
        h1 = rr.intersect(rays.view(np.float32).reshape(-1, 8), threads=8)
        h2 = rp.intersect(rays.view(np.float32).reshape(-1, 8), threads=8)
        assert (h1["object"][h1["hit"] == 1] == 0).mean() > 0.05
        for k in ("hit", "object", "element"):
            assert np.array_equal(h1[k], h2[k])
        h = h1["hit"] == 1
        assert np.array_equal(bits(h1["distance"][h]), bits(h2["distance"][h])) and np.array_equal(bits(h1["uv"][h]), bits(h2["uv"][h]))
    p = make_params(resolution=72, shader=shader)
    sr, sp = rr.state(cam, p), rp.state(cam, p)
    sr.trace_samples(2), sp.trace_samples(2)
    (_, a1, s1, g1), (_, a2, s2, g2) = sr.download(), sp.download()
    assert np.array_equal(s1, s2) and np.array_equal(g1, g2) and np.array_equal(bits(a1), bits(a2))
    for x in (sr, sp, rr, rp):
        x.close()


def test_port_tree_order_equals_the_reference(port, refo):
    """shape->bvh->primitives after build_bvh (the leaf order exact-t ties are broken by): the port's builder, with its
    restatement of libstdc++'s std::partition, leaves the primitives where the reference's does."""
    for scene in (synth.hair_block_scene(30_000)[0], synth.curly_hair_scene(60_000, sky=False)[0], synth.points_scene(5_000, 2_000, sky=False)[0]):
        a, b = refo.scene(scene), port.scene(scene)
        for k in range(len(scene.shapes)):
            assert np.array_equal(a.shape_order(k), b.shape_order(k))
        a.close(); b.close()
