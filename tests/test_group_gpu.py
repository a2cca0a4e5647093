"""Several GPUs in one process behind the C ABI (yh_group_*).

A device may be listed more than once (each entry is one rank), so the whole N > 1
path - replicated scene, per-rank states, shards, the fused peer reduce + resolve -
runs on a one-GPU box too; with two or more GPUs the same tests use distinct
devices (real peer loads over NVLink) and the NCCL reduce is compared with the
peer kernel."""
import ctypes as C

import numpy as np
import pytest

from yocto_hair_b200 import _capi as capi
from yocto_hair_b200 import synth
from yocto_hair_b200.pathtrace import DeviceGroup, DeviceScene
from yocto_hair_b200.scene import make_params

pytestmark = pytest.mark.gpu


def device_list(n):
    have = C.c_int()
    capi.check(capi.load().yh_device_count(C.byref(have)))
    return [i % max(1, have.value) for i in range(n)], have.value


@pytest.mark.parametrize("ranks", [1, 2, 4])
def test_tile_sharded_group_is_bit_identical_to_the_reference(ranks, ref):
    scene, cam = synth.straight_hair_scene(20_000)  # hair + two area lights + sky
    p = make_params(resolution=160)
    rs = ref.scene(scene)
    sr = rs.state(cam, p)
    sr.trace_samples(3)
    r_ref, a_ref, s_ref = sr.download()[:3]
    devices, _ = device_list(ranks)
    grp = DeviceGroup(scene, devices=devices).init_state(cam, p, "tiles")
    st = grp.trace_samples(2)
    assert st.samples == grp.width * grp.height * 2
    grp.trace_samples(1)
    r, a, s = grp.download()
    assert np.array_equal(s, s_ref)
    assert np.array_equal(a.view(np.uint32), a_ref.view(np.uint32))
    assert np.array_equal(r.view(np.uint32), r_ref.view(np.uint32))
    # a second download after more samples starts from the members' buffers again (nothing is double counted)
    grp.trace_samples(1)
    sr.trace_samples(1)
    r2, a2, s2 = grp.download()
    assert np.array_equal(a2.view(np.uint32), sr.download()[1].view(np.uint32)) and (s2 == 4).all()
    # a camera edit re-runs init_state on every member
    grp.init_state(cam, make_params(resolution=96), "tiles")
    grp.trace_samples(1)
    assert grp.download()[0].shape == (96, 96, 4)
    grp.close(); sr.close(); rs.close()


def test_sample_sharded_group_adds_up_to_the_counter_mode_image():
    scene, cam = synth.curly_hair_scene(30_000)
    p = make_params(resolution=128)
    ds = DeviceScene(scene).init_bvh().init_lights()
    one = ds.init_state(cam, p).set_sample_shard(0, 1)
    one.trace_samples(8)
    _, a_one, s_one = one.download()
    devices, _ = device_list(4)
    grp = DeviceGroup(scene, devices=devices).init_state(cam, p, "samples")
    with pytest.raises(capi.YhError):
        grp.trace_samples(6)  # not a multiple of the rank count
    grp.trace_samples(8)
    r, a, s = grp.download()
    assert np.array_equal(s, s_one) and (s == 8).all()
    assert np.allclose(a, a_one, rtol=2e-6, atol=1e-6)  # same samples, other fp32 summation order
    # the group's sum is the rank-ordered sum of the four shards, bit for bit
    acc = None
    for rank in range(4):
        st = ds.init_state(cam, p).set_sample_shard(rank, 4)
        st.trace_samples(2)
        _, ar, _ = st.download()
        acc = ar.copy() if acc is None else acc + ar
        st.close()
    assert np.array_equal(acc.view(np.uint32), a.view(np.uint32))
    assert np.array_equal(r.view(np.uint32), (a / s[..., None].astype(np.float32)).view(np.uint32))
    grp.close(); one.close(); ds.close()


def test_group_errors_are_loud():
    scene, cam = synth.curly_hair_scene(2_000, sky=False)
    lib = capi.load()
    desc, keep = scene.to_desc()
    h = C.c_void_p()
    bad = (C.c_int * 1)(99)
    assert lib.yh_group_create(C.byref(desc), None, bad, 1, C.byref(h)) == capi.YH_ERR_INVALID
    assert lib.yh_group_create(C.byref(desc), None, None, 0, C.byref(h)) == capi.YH_ERR_INVALID
    grp = DeviceGroup(scene, devices=[0])
    with pytest.raises(capi.YhError) as e:
        grp.trace_samples(1)
    assert e.value.status == capi.YH_ERR_STATE
    scene.materials[0].specular = 0.5  # out of scope on a line shape: every member refuses
    with pytest.raises(capi.YhError) as e:
        DeviceGroup(scene, devices=[0, 0])
    assert e.value.status == capi.YH_ERR_UNSUPPORTED
    grp.close()
    del keep


def test_nccl_reduce_equals_the_peer_kernel():
    devices, have = device_list(2)
    if have < 2:
        pytest.skip("ncclCommInitAll needs distinct devices (this box has one GPU)")
    scene, cam = synth.curly_hair_scene(30_000)
    p = make_params(resolution=160)
    grp = DeviceGroup(scene, devices=list(range(min(have, 8)))).init_state(cam, p, "tiles")
    grp.trace_samples(3)
    r0, a0, s0 = grp.download()
    grp.set_reduce("nccl")
    r1, a1, s1 = grp.download()
    assert np.array_equal(s0, s1) and np.array_equal(a0.view(np.uint32), a1.view(np.uint32))
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
    grp.close()
