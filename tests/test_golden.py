"""Oracle against the committed golden vectors (tests/golden/*.npz, generated from the
unmodified reference by tests/golden/make_golden.py).  These run anywhere -- they need
neither /root/reference nor oracle/_ref.  Bit-exact unless stated."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import random_rays_in_room

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden as mg  # noqa: E402  (input generators only; seeds regenerate the inputs)


def gold(name):
    return np.load(os.path.join(GOLD, name))


def test_golden_primitives(ol):
    g = gold("prims.npz")
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rays, cs, rs = mg.sphere_cases()
    for i in range(len(rays)):
        t = C.c_double()
        hit = ol.oracle().oracle_intersect_sphere(p(rays[i]), p(cs[i]), float(rs[i]), C.byref(t))
        assert bool(hit) == bool(g["sphere_hit"][i])
        if hit:
            assert t.value == g["sphere_t"][i]
    rays, verts = mg.triangle_cases()
    for i in range(len(rays)):
        out = np.zeros(3)
        hit = ol.oracle().oracle_intersect_triangle(p(rays[i]), p(verts[i]), p(out))
        assert bool(hit) == bool(g["tri_hit"][i])
        if hit:
            assert np.array_equal(out, g["tri_tuv"][i])
    assert g["sphere_hit"].sum() > 100 and g["tri_hit"].sum() > 200


def test_golden_camera(ol, api):
    g = gold("camera.npz")
    for key in ("320x180", "1920x1080", "512x512"):
        w, h = map(int, key.split("x"))
        assert np.array_equal(ol.init_camera(w, h).as_array(), g[key])
        assert np.array_equal(api.init_camera(w, h).as_array(), g[key])
    cam = ol.init_camera(320, 180)
    uv = np.random.default_rng(104).uniform(0, 1, (64, 2))
    got = np.array([ol.camera_ray(cam, u, v) for u, v in uv])
    assert np.array_equal(got, g["cam_rays"])


def test_golden_nearest_hit(ol, api):
    g = gold("c1_hits.npz")
    objs = api.scene_default(320, 180)
    rays = random_rays_in_room(np.random.default_rng(105), 3000)
    got = ol.intersect_rays(objs, rays)
    assert np.array_equal(got["ids"], g["ids"])
    for k in ("points", "normals", "uvs"):
        assert np.array_equal(got[k], g[k]), k


def test_golden_paths(ol, api):
    g = gold("c1_paths.npz")
    objs = api.scene_default(320, 180)
    rays, streams = _path_cases(ol)
    for i in range(len(rays)):
        rad, used = ol.trace_path_stream(objs, rays[i], streams[i], max_depth=5, dielectric="split")
        assert used == g["used"][i]
        assert np.array_equal(rad, g["radiance"][i])


def _path_cases(ol, seed=103, n=200):
    """same generator as make_golden.path_cases, with the oracle's camera (bit-identical to
    the reference's, test_golden_camera) so the test does not need oracle/_ref"""
    rng = np.random.default_rng(seed)
    cam = ol.init_camera(320, 180)
    rays = np.array([ol.camera_ray(cam, rng.uniform(0, 1), rng.uniform(0, 1)) for _ in range(n)])
    streams = rng.integers(0, 2 ** 31, size=(n, 2048), dtype=np.int64).astype(np.int32)
    return rays, streams


@pytest.mark.parametrize("cfg", [(64, 36, 6, 5), (48, 27, 4, 8)])
def test_golden_frames(ol, api, cfg):
    W, H, S, depth = cfg
    g = gold("c1_frames.npz")
    objs = api.scene_default(W, H)
    cam = ol.init_camera(W, H)
    tag = f"{W}x{H}_s{S}_d{depth}"
    s, ctr = ol.render_sum(objs, cam, W, H, S, rng="libc", dielectric="split", max_depth=depth)
    assert np.array_equal(ol.tonemap(s, S), g["fb_" + tag])
    assert np.array_equal(s * (1.0 / S), g["mean_" + tag])
    assert tuple(ctr) == tuple(g["ctr_" + tag])


def test_golden_converged_noise_floor(ol, api):
    """the converged reference render used by the GPU PSNR gate is self-consistent: its
    independent quarter-spp twin agrees with it at the expected Monte Carlo level"""
    from conftest import psnr_u8
    g = gold("c1_converged_96x54.npz")
    hi = ol.tonemap(g["mean"].astype(np.float64), 1)
    lo = ol.tonemap(g["mean_quarter"].astype(np.float64), 1)
    p = psnr_u8(hi, lo)
    assert 28.0 < p < 50.0, p  # 1024 spp vs 2048 spp ~ 33 dB at 320x180


def test_oracle_cast_ray_matches_the_reference_golden(ol, api):
    """Whitted integrator: the restated cast_ray against the reference's own cast_ray outputs
    (tests/golden/whitted.npz, made by make_golden.py from oracle/_ref): bit-identical colours
    and identical cast_ray call counts"""
    g = np.load(os.path.join(GOLD, "whitted.npz"))
    rays = random_rays_in_room(np.random.default_rng(106), 3000)
    rgb, calls = ol.cast_rays(api.scene_default(320, 180), rays, max_depth=5)
    assert np.array_equal(rgb, g["c1_rgb"]) and np.array_equal(calls, g["c1_calls"])
    field = api.scene_sphere_field(400, 96, 54, mix=(0.3, 0.3, 0.3))
    rgb, calls = ol.cast_rays(field, rays, max_depth=8)
    assert np.array_equal(rgb, g["field_rgb"]) and np.array_equal(calls, g["field_calls"])
    assert calls.max() > 4 and (g["c1_rgb"] != g["c1_rgb"][0]).any()
