"""The reference's leaf order, grown on the device (yh_tieorder.cu), against the reference itself.

An exact-t tie between two primitives of a shape goes to the one the reference's BVH visits last, so the library
restates that tree - build_bvh / split_middle with std::partition's own permutation (yocto_pathtrace.cpp:564-650) -
level by level on the GPU.  Here the resulting primitive order is compared element for element with
shape->bvh->primitives of the reference build (oracle/_ref) or of the plain-C port."""
import numpy as np
import pytest

from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceScene
from yocto_hair_b200.scene import Scene, Shape, Material, Object, Environment

pytestmark = pytest.mark.gpu


def one_shape_scene(shape):
    sc = Scene()
    sc.shapes.append(shape)
    sc.materials.append(Material(color=(0.5, 0.5, 0.5)))
    sc.objects.append(Object(shape=0, material=0, frame=(1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)))
    sc.environments.append(Environment(emission=(1, 1, 1)))
    return sc


def degenerate_lines(n, seed=3):
    """Many identical centres (median fallback of split_middle), exact lattice positions (centres ON the split planes),
    a geometric run (deep one-sided splits) and a random rest."""
    rng = np.random.default_rng(seed)
    q = n // 4
    a = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (q, 1))
    b = (rng.integers(0, 8, (q, 3)) / 8.0).astype(np.float32)
    c = np.stack([2.0 ** -np.arange(q, dtype=np.float32) % 1.0 + 1e-3 * np.arange(q, dtype=np.float32), np.zeros(q, np.float32), np.ones(q, np.float32)], axis=1)
    d = rng.uniform(0, 1, (n - 3 * q, 3)).astype(np.float32)
    p0 = np.concatenate([a, b, c, d]).astype(np.float32)
    p1 = p0 + np.array([0.01, 0.0, 0.0], np.float32)
    pos = np.stack([p0, p1], axis=1).reshape(-1, 3)
    lines = np.arange(2 * n, dtype=np.int32).reshape(-1, 2)
    return Shape(positions=pos, normals=np.tile(np.array([[1, 0, 0]], np.float32), (2 * n, 1)), radius=np.full(2 * n, 0.002, np.float32), lines=lines)


CASES = {
    "curly_200k": lambda: synth.curly_hair_scene(200_000, sky=False)[0],
    "straight_50k": lambda: synth.straight_hair_scene(50_000, sky=False, area_lights=True)[0],  # + a two-triangle quad
    "block_sphere": lambda: synth.hair_block_scene(30_000)[0],                                 # lines + 4096 triangles
    "points": lambda: synth.points_scene(20_000, 2_000, sky=False)[0],
    "degenerate": lambda: one_shape_scene(degenerate_lines(40_000)),
    "tiny": lambda: one_shape_scene(degenerate_lines(5)),
    "four": lambda: one_shape_scene(degenerate_lines(4)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_order_equals_the_reference_tree(name, ref):
    scene = CASES[name]()
    rs = ref.scene(scene)
    ds = DeviceScene(scene).init_bvh().init_lights()
    for k in range(len(scene.shapes)):
        want, got = rs.shape_order(k), ds.reference_order(k)
        assert len(want) == len(got)
        assert np.array_equal(np.sort(got), np.arange(len(got)))  # a permutation
        first = np.nonzero(want != got)[0]
        assert first.size == 0, "shape %d: first difference at position %d of %d (%s vs %s)" % (k, first[0], len(got), want[first[0]:first[0] + 6], got[first[0]:first[0] + 6])
    ds.close(); rs.close()


def test_million_segments_order(ref):
    """BASELINE configs[2] geometry: the order of 1 M segments, built on the device in a few milliseconds."""
    scene = synth.curly_hair_scene(1_000_000, sky=False)[0]
    rs = ref.scene(scene)
    ds = DeviceScene(scene).init_bvh().init_lights()
    assert np.array_equal(rs.shape_order(0), ds.reference_order(0))
    ds.close(); rs.close()
