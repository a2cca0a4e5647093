"""Parity at the sizes BASELINE.json names ("parity at BASELINE sizes"): the configurations the
bench measures are compared with the reference build itself, not only their small stand-ins.

  C3  1 M curly segments, 1280 x 1280      2 spp: accumulation, sample counts and per-pixel RNG bit-identical
  C2  250 k straight segments, 720 x 720   4 spp: the same, with the two area lights and the sky
  C2  converged image at 720 x 720         rel-MSE at the Monte Carlo noise floor (SURVEY.md 8d tolerance)
  C5  5 M segments                         >= 10^6 hit records bit-exact
The CPU side costs about a second per sample per megapixel on the box's host cores, so the sample counts are small;
the comparison is bit for bit, which a single differing path anywhere in the image would break."""
import numpy as np
import pytest

from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import make_params

from test_trace_gpu import assert_hits_identical, ray_set, rel_mse

pytestmark = pytest.mark.gpu


def bench_scene(kind, segments):
    """The scene bench.py builds for this workload (same generator, same sky)."""
    import bench
    return bench.build_scene(kind, segments, True)


@pytest.mark.parametrize("name,kind,segments,res,spp", [
    ("c3", "curly", 1_000_000, 1280, 2),
    ("c2", "straight", 250_000, 720, 4),
])
def test_render_at_baseline_size_is_bit_identical(name, kind, segments, res, spp, ref):
    if ref.kind != "reference":
        pytest.skip("needs oracle/_ref (the reference build)")
    scene, cam = bench_scene(kind, segments)
    p = make_params(resolution=res)
    rs = ref.scene(scene)
    sr = rs.state(cam, p)
    sr.trace_samples(spp)
    _, a_r, s_r, rng_r = sr.download()
    ds = DeviceScene(scene).init_bvh().init_lights()
    sd = ds.init_state(cam, p)
    sd.trace_samples(spp)
    _, a_d, s_d = sd.download()
    assert a_d.shape == (res, res, 4)
    assert np.array_equal(s_r, s_d)
    diff = np.nonzero((a_r.view(np.uint32) != a_d.view(np.uint32)).any(axis=-1))
    assert diff[0].size == 0, "%d pixels differ, first (y, x): %s" % (diff[0].size, list(zip(diff[0][:8], diff[1][:8])))
    assert np.array_equal(rng_r, sd.rng())
    sd.close(); ds.close(); sr.close(); rs.close()


def test_c2_converged_image_rel_mse_at_720(ref):
    """north_star's third check at the bench resolution: a GPU render with another seed against two reference renders
    of the 250 k-segment scene at 720 x 720.  Tolerance: relMSE(GPU, CPU_A) <= 1.5 x relMSE(CPU_B, CPU_A), mean
    luminance within 0.5 %."""
    scene, cam = bench_scene("straight", 250_000)
    res, spp = 720, 24
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
    floor, err = rel_mse(imgs["B"], imgs["A"]), rel_mse(gpu, imgs["A"])
    lum = lambda im: float((im[..., :3] @ np.array([0.2126, 0.7152, 0.0722])).mean())  # noqa: E731
    print("720x720 x %d spp: relMSE gpu-vs-A %.4g, B-vs-A %.4g, lum gpu %.5f A %.5f" % (spp, err, floor, lum(gpu), lum(imgs["A"])))
    assert err <= 1.5 * floor
    assert abs(lum(gpu) / lum(imgs["A"]) - 1) <= 5e-3
    st.close(); ds.close(); rs.close()


def test_c5_hit_records_bit_exact(ref):
    """BASELINE configs[4] geometry (5 M segments, the one tree that leaves L2): a million hit records, both builders'
    default, bit-exact against the reference's own BVH walk."""
    scene, cam = bench_scene("curly", 5_000_000)
    rays = ray_set(scene, cam, 600_000, 450_000)
    rs = ref.scene(scene)
    hr = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=32)
    ds = DeviceScene(scene).init_bvh().init_lights()
    assert ds.bvh_stats().num_primitives == 5_000_000
    hd = ds.intersect(rays)
    assert len(rays) >= 1_000_000 and 0.05 < hr["hit"].mean() < 0.95
    assert_hits_identical(hr, hd)
    ds.close(); rs.close()
