"""Whitted integrator (cast_ray, raytracer.c:556-641) on the GPU: SURVEY.md 8(f) N4.

cast_ray draws no random numbers, so GPU, oracle and reference are compared ray by ray.
Bar: `ray_count` (cast_ray invocations) equal exactly; colours within 1e-12 relative of the
reference's doubles -- both sides run the same double operations in the same order, the only
difference is libm vs the CUDA math library in pow()/atan2()/fmod() (a few ulp)."""
import os

import numpy as np
import pytest

from conftest import random_rays_in_room

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12


def close(a, b):
    return np.abs(a - b).max() <= RTOL * max(1.0, np.abs(b).max())


def test_cast_rays_match_the_reference_golden(gpu_api):
    """against the unmodified reference's cast_ray outputs (tests/golden/whitted.npz)"""
    g = np.load(os.path.join(GOLD, "whitted.npz"))
    rays = random_rays_in_room(np.random.default_rng(106), 3000)
    with gpu_api.Scene(gpu_api.scene_default(320, 180)) as sc:
        rgb, calls = sc.cast_rays(rays, max_depth=5)
    assert np.array_equal(calls.astype(np.int64), g["c1_calls"])
    assert close(rgb, g["c1_rgb"])
    # the arithmetic is the same operation for operation: most colours are bit-identical
    assert (rgb == g["c1_rgb"]).all(axis=1).mean() > 0.9
    with gpu_api.Scene(gpu_api.scene_sphere_field(400, 96, 54, mix=(0.3, 0.3, 0.3))) as sc:
        rgb, calls = sc.cast_rays(rays, max_depth=8)
    assert np.array_equal(calls.astype(np.int64), g["field_calls"])
    assert close(rgb, g["field_rgb"])


@pytest.mark.parametrize("depth", [0, 1, 5, 12])
def test_cast_rays_match_the_oracle_on_a_mesh_scene(gpu_api, ol, depth):
    """mesh + spheres (the reference has no live mesh path: the oracle restatement is the checker)"""
    W, H = 96, 54
    holder = gpu_api.mesh_room(gpu_api.heightfield_mesh(24, 20 * W / H * 0.98), W, H)
    rays = random_rays_in_room(np.random.default_rng(7 + depth), 1500)
    with gpu_api.Scene(holder) as sc:
        rgb, calls = sc.cast_rays(rays, max_depth=depth)
    ref_rgb, ref_calls = ol.cast_rays(holder, rays, max_depth=depth)
    assert np.array_equal(calls.astype(np.int64), ref_calls)
    assert close(rgb, ref_rgb)


def test_whitted_render_matches_the_oracle(gpu_api, ol):
    """render() with the Whitted integrator: same jitter (Philox), per-pixel sums and 8-bit frame"""
    W, H, SPP = 96, 54, 3
    objs = gpu_api.scene_sphere_field(120, W, H, mix=(0.2, 0.3, 0.3))
    cam = gpu_api.init_camera(W, H)
    with gpu_api.Scene(objs) as sc:
        fb, acc, ctr = sc.render(cam, gpu_api.make_desc(W, H, 2, 2 + SPP, max_depth=6, integrator=1), want_accum=True)
    ref_sum, (rays, _) = ol.render_sum(objs, cam, W, H, SPP, rng="philox", max_depth=6, sample_offset=2,
                                       integrator="whitted")
    assert ctr.rays == rays and ctr.paths == W * H * SPP
    np.testing.assert_allclose(acc, ref_sum, rtol=2e-7, atol=1e-7)  # the accumulation buffer is float32
    ref_fb = ol.tonemap(ref_sum, SPP)
    diff = np.abs(fb.astype(int) - ref_fb.astype(int))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999


def test_whitted_through_the_drop_in(gpu_api, ol, abi):
    """render_ex() of libraytracer_b200.so with RenderParams.integrator = RT_INTEGRATOR_WHITTED"""
    import ctypes as C
    W, H, SPP = 64, 36, 2
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    _, host = gpu_api.load()
    opt = abi.Options()
    opt.width, opt.height, opt.samples = W, H, SPP
    rp = abi.RenderParams()
    host.render_params_default(rp)
    rp.integrator = 1
    fb = np.zeros((H, W, 3), np.uint8)
    host.render_ex(fb.ctypes.data, objs.ctypes.data, len(objs), C.byref(cam), C.byref(opt), C.byref(rp))
    ref_sum, _ = ol.render_sum(objs, cam, W, H, SPP, rng="philox", max_depth=5, integrator="whitted")
    diff = np.abs(fb.astype(int) - ol.tonemap(ref_sum, SPP).astype(int))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.999
    assert fb.std() > 1  # not a constant frame


def test_whitted_rejects_deep_recursion(gpu_api):
    W, H = 32, 18
    with gpu_api.Scene(gpu_api.scene_default(W, H)) as sc:
        with pytest.raises(Exception):
            sc.render(gpu_api.init_camera(W, H), gpu_api.make_desc(W, H, 0, 1, max_depth=16, integrator=1))
