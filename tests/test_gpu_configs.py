"""Parity on the SHAPES of the named configs (BASELINE.json configs / SURVEY.md 8d), at frames the CPU
oracle finishes in seconds: the full scene of each config (all 10 000 spheres, depth 64, ...) with a
reduced frame and sample count.  GPU through the C ABI vs oracle/oracle.c with the same Philox streams:
per-pixel float sums to 1e-3 relative, ray_count exactly.  Plus the PSNR gate at the reference's own default
frame, 320x180 (main.c:24-30), against a converged render of the UNMODIFIED reference.
"""
import os

import numpy as np
import pytest

from conftest import psnr_u8

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _compare(gpu_api, ol, source, W, H, spp, depth, rtol=1e-3, atol=1e-4, sample_begin=0):
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, sample_begin, sample_begin + spp, max_depth=depth)
    with gpu_api.Scene(source) as sc:
        fb, acc, ctr = sc.render(cam, desc, want_accum=True)
    want, (rays, _) = ol.render_sum(source, cam, W, H, spp, rng="philox", dielectric="stochastic", max_depth=depth,
                                    sample_offset=sample_begin)
    assert ctr.rays == rays, "ray_count (trace_path invocations, raytracer.c:484) must match the oracle exactly"
    assert ctr.paths == W * H * spp
    np.testing.assert_allclose(acc, want, rtol=rtol, atol=atol)
    want_fb = ol.tonemap(want, spp)
    assert np.abs(fb.astype(int) - want_fb.astype(int)).max() <= 1
    return ctr


def test_c2_shape_10k_spheres_depth_8(gpu_api, ol):
    """C2: 10 000 random spheres, mixed materials, max depth 8 -- the whole integrator, not only nearest hits"""
    objs = gpu_api.scene_sphere_field(10000, 1920, 1080)
    assert len(objs) >= 10000
    ctr = _compare(gpu_api, ol, objs, 160, 90, 1, 8, sample_begin=5)
    assert 2.0 < ctr.rays / ctr.paths < 10.0


def test_c4_shape_dielectric_metal_heavy(gpu_api, ol):
    """C4: the 10 000-sphere field with 40 % dielectric / 40 % mirror at the 4K aspect, depth 8"""
    objs = gpu_api.scene_sphere_field(10000, 3840, 2160, mix=(0.2, 0.4, 0.4))
    _compare(gpu_api, ol, objs, 128, 72, 1, 8, rtol=2e-3, atol=2e-4)


def test_c5_shape_all_dielectric_depth_64(gpu_api, ol):
    """C5: deep-bounce stress -- 2 000 spheres, 90 % dielectric, max depth 64: bounce words up to 0x140,
    130 kernel launches per wave, paths that survive all 65 levels"""
    objs = gpu_api.scene_sphere_field(2000, 512, 512, mix=(0.1, 0.9, 0.0))
    ctr = _compare(gpu_api, ol, objs, 96, 96, 2, 64, rtol=2e-3, atol=2e-4)
    assert ctr.rays / ctr.paths > 6.0, "the stress scene must actually go deep"


def test_c5_depth_boundaries(gpu_api, ol):
    """depths around the encoding limits: 63/64/65 on a mirror + dielectric field (no early roulette exits)"""
    objs = gpu_api.scene_sphere_field(300, 256, 256, mix=(0.0, 0.5, 0.5))
    for depth in (63, 65, 128):
        _compare(gpu_api, ol, objs, 32, 32, 1, depth, rtol=3e-3, atol=3e-4)


def test_c3_shape_mesh_room_depth_5(gpu_api, ol):
    """C3: the height-field mesh room (here 2 x 64 x 64 triangles) + its spheres, depth 5, two samples"""
    W, H = 96, 54
    verts = gpu_api.heightfield_mesh(64, 20 * W / H * 0.98)
    holder = gpu_api.mesh_room(verts, W, H)
    _compare(gpu_api, ol, holder, W, H, 2, 5)


def test_psnr_c1_at_the_reference_default_frame(gpu_api, ol):
    """C1 at 320x180 (main.c:24-30) against 32 768 spp of the unmodified reference (8 srand() seeds x 4 096
    spp, tests/golden/make_converged.py).  Gate: >= 40 dB, or the reference's own noise floor + 1 dB
    (an independent 8 192-spp reference render against the 32 768-spp one)."""
    path = os.path.join(GOLD, "c1_converged_320x180.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/c1_converged_320x180.npz not generated")
    W, H, gpu_spp = 320, 180, 131072
    g = np.load(path)
    ref_full, ref_quarter = g["mean"].astype(np.float64), g["mean_quarter"].astype(np.float64)
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    with gpu_api.Scene(objs) as sc:
        _, acc, ctr = sc.render(cam, gpu_api.make_desc(W, H, 0, gpu_spp, max_depth=5), want_accum=True)
    gpu_mean = acc.astype(np.float64) / gpu_spp
    fb_ref, fb_q, fb_gpu = ol.tonemap(ref_full, 1), ol.tonemap(ref_quarter, 1), ol.tonemap(gpu_mean, 1)
    floor, got = psnr_u8(fb_q, fb_ref), psnr_u8(fb_gpu, fb_ref)
    bias = (gpu_mean.mean() - ref_full.mean()) / ref_full.mean()
    print(f"C1 320x180: PSNR(gpu {gpu_spp} spp, ref {int(g['spp'][0])} spp) = {got:.2f} dB; reference noise floor = {floor:.2f} dB; "
          f"mean bias {bias:+.4f}; {ctr.rays / ctr.gpu_ms / 1e3:.0f} Mrays/s")
    assert got >= min(40.0, floor + 1.0), (got, floor)
    assert abs(bias) < 0.01


# ---- meshes whose vertices are genuine doubles (apply_matrix, main.c:140-147) --------------------------

def _transformed_mesh(gpu_api, grid, W, H):
    """a height-field mesh rotated, scaled and shifted in double: its coordinates are no longer floats"""
    verts = gpu_api.heightfield_mesh(grid, 20 * W / H * 0.5)
    a, b = 0.3, -0.2
    ry = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]])
    rx = np.array([[1, 0, 0, 0], [0, np.cos(b), -np.sin(b), 0], [0, np.sin(b), np.cos(b), 0], [0, 0, 0, 1]])
    m = ry @ rx @ np.diag([0.7, 1.3, 0.9, 1.0])
    m[:3, 3] = [0.123456789, 1.0 / 3.0, -2.0 / 7.0]
    gpu_api.apply_matrix(verts, m)
    assert not np.array_equal(verts["pos"], verts["pos"].astype(np.float32).astype(np.float64))
    return verts


def test_double_vertices_nearest_hit_is_bit_exact(gpu_api, ol):
    """the caller's double vertices reach the exact triangle test unrounded: ids, t, points and normals equal
    the oracle's brute-force loop over the SAME doubles, bit for bit (round 1 narrowed them to float silently)"""
    W, H = 96, 54
    verts = _transformed_mesh(gpu_api, 48, W, H)
    holder = gpu_api.mesh_room(verts, W, H)
    rng = np.random.default_rng(23)
    rays = random_rays_in_room_local(rng, 6000)
    want = ol.intersect_rays(holder, rays)
    with gpu_api.Scene(holder, all_trees=True) as sc:
        assert sc.info.double_triangles == 1
        for mode in (1, 3, 4, 5, 0):
            got = sc.trace_rays(rays, use_bvh=mode)
            assert np.array_equal(got["ids"], want["ids"]), mode
            hit = want["ids"] >= 0
            assert np.array_equal(got["points"][hit], want["points"][hit]), mode
            assert np.array_equal(got["normals"][hit], want["normals"][hit]), mode
    # a float-representable mesh keeps the compact float records
    with gpu_api.Scene(gpu_api.mesh_room(gpu_api.heightfield_mesh(8, 10.0), W, H)) as sc:
        assert sc.info.double_triangles == 0


def _with_second_mesh(abi, holder, verts):
    """the scene of `holder` plus one more mesh object (material of the first mesh)"""
    import ctypes as C
    first = next(k for k in range(holder.n) if holder.objects[k].type == abi.GEOMETRY_MESH)
    h = abi.SceneHolder()
    h.n = holder.n + 1
    h.objects = (abi.SceneObject * h.n)()
    C.memmove(h.objects, holder.objects, C.sizeof(abi.SceneObject) * holder.n)
    mesh = abi.TriangleMesh()
    mesh.num_triangles = len(verts) // 3
    mesh.vertices = C.cast(verts.ctypes.data, C.POINTER(abi.Vertex))
    C.memmove(C.byref(h.objects[holder.n]), C.byref(holder.objects[first]), C.sizeof(abi.SceneObject))
    h.objects[holder.n].geometry.mesh = C.pointer(mesh)
    h._keep += [holder, verts, mesh]
    return h


def test_narrowed_upload_is_bit_exact(gpu_api, ol, abi, monkeypatch):
    """a pageable mesh large enough for the narrowed upload (60 B per triangle, converted to float by the host's
    staging threads): same nearest hits, points, normals and texture coordinates as the oracle's brute-force loop,
    and the same as the raw 120 B upload (RTB_UPLOAD_NARROW=0)"""
    W, H = 96, 54
    verts = gpu_api.heightfield_mesh(72, 20 * W / H * 0.9)  # 10 368 triangles: above the narrowing threshold
    holder = gpu_api.mesh_room(verts, W, H)
    rays = random_rays_in_room_local(np.random.default_rng(29), 3000)
    want = ol.intersect_rays(holder, rays)
    with gpu_api.Scene(holder) as sc:
        assert sc.info.double_triangles == 0
        got = sc.trace_rays(rays)
    monkeypatch.setenv("RTB_UPLOAD_NARROW", "0")
    with gpu_api.Scene(holder) as sc:
        raw = sc.trace_rays(rays)
    monkeypatch.delenv("RTB_UPLOAD_NARROW")
    hit = want["ids"] >= 0
    assert hit.sum() > 1000
    assert np.array_equal(got["ids"], want["ids"])
    assert np.array_equal(got["points"][hit], want["points"][hit])
    assert np.array_equal(got["normals"][hit], want["normals"][hit])
    for k in ("ids", "points", "normals", "uvs"):
        assert np.array_equal(got[k], raw[k]), k


def test_narrowed_and_double_meshes_in_one_scene(gpu_api, ol, abi):
    """one float-representable mesh (uploaded narrowed) and one mesh transformed in double (uploaded raw): the
    scene keeps double vertices for BOTH -- the narrowed piece's doubles are its floats, widened -- and every
    nearest hit equals the oracle's brute-force loop bit for bit"""
    W, H = 96, 54
    exact = gpu_api.heightfield_mesh(72, 20 * W / H * 0.9)
    moved = _transformed_mesh(gpu_api, 72, W, H)
    moved["pos"][:, 1] += 6.0  # above the first mesh
    holder = _with_second_mesh(abi, gpu_api.mesh_room(exact, W, H), moved)
    rays = random_rays_in_room_local(np.random.default_rng(31), 3000)
    want = ol.intersect_rays(holder, rays)
    with gpu_api.Scene(holder) as sc:
        assert sc.info.double_triangles == 1
        assert sc.info.n_triangles == (len(exact) + len(moved)) // 3
        got = sc.trace_rays(rays)
    hit = want["ids"] >= 0
    n_first = len(exact) // 3
    tri_hits = want["ids"][hit]
    assert (tri_hits >= 6).sum() > 500  # hits on both meshes (ids follow the reference's loop order)
    assert np.array_equal(got["ids"], want["ids"])
    assert np.array_equal(got["points"][hit], want["points"][hit])
    assert np.array_equal(got["normals"][hit], want["normals"][hit])
    # and the other order: the double mesh first
    holder2 = _with_second_mesh(abi, gpu_api.mesh_room(moved, W, H), exact)
    want2 = ol.intersect_rays(holder2, rays)
    with gpu_api.Scene(holder2) as sc:
        assert sc.info.double_triangles == 1
        got2 = sc.trace_rays(rays)
    hit2 = want2["ids"] >= 0
    assert np.array_equal(got2["ids"], want2["ids"])
    assert np.array_equal(got2["points"][hit2], want2["points"][hit2])
    assert np.array_equal(got2["normals"][hit2], want2["normals"][hit2])
    del n_first


def test_double_vertices_render_matches_oracle(gpu_api, ol):
    W, H = 64, 36
    verts = _transformed_mesh(gpu_api, 32, W, H)
    holder = gpu_api.mesh_room(verts, W, H)
    _compare(gpu_api, ol, holder, W, H, 2, 5)


def random_rays_in_room_local(rng, n):
    from conftest import random_rays_in_room
    return random_rays_in_room(rng, n)


# ---- the reference's own dielectric estimator: deterministic two-way split (raytracer.c:522-529) --------

@pytest.mark.parametrize("scene,depth", [("c1", 5), ("dielectric", 5), ("dielectric", 3), ("mesh", 4)])
def test_dielectric_split_matches_oracle(gpu_api, ol, scene, depth):
    """RTB_DIELECTRIC_SPLIT traces BOTH children at a dielectric vertex like upstream; against the oracle in the
    same mode (which is bit-identical to the reference under libc rand()): ray_count exact -- it now counts the
    whole split tree -- and sums to 1e-3"""
    W, H, SPP = 64, 36, 3
    if scene == "c1":
        src = gpu_api.scene_default(W, H)
    elif scene == "dielectric":
        src = gpu_api.scene_sphere_field(200, W, H, mix=(0.1, 0.6, 0.2))
    else:
        src = gpu_api.mesh_room(gpu_api.heightfield_mesh(24, 20 * W / H * 0.98), W, H)
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, 1, 1 + SPP, max_depth=depth, dielectric=1)
    with gpu_api.Scene(src) as sc:
        fb, acc, ctr = sc.render(cam, desc, want_accum=True)
        _, acc_st, ctr_st = sc.render(cam, gpu_api.make_desc(W, H, 1, 1 + SPP, max_depth=depth), want_accum=True)
    want, (rays, _) = ol.render_sum(src, cam, W, H, SPP, rng="philox", dielectric="split", max_depth=depth, sample_offset=1)
    assert ctr.rays == rays
    np.testing.assert_allclose(acc, want, rtol=1e-3, atol=1e-4)
    # the sphere field and the mesh room (its glass ball) hold M_REFRACTION surfaces, the default scene may not
    has_dielectric = scene in ("dielectric", "mesh") or bool((np.asarray(src["flags"]) & 8).any())
    if has_dielectric:
        assert ctr.rays > ctr_st.rays, "the split traces more rays than the one-child estimator"
    else:
        assert ctr.rays == ctr_st.rays  # no dielectric surface: same paths; sums equal up to the order of additions
        np.testing.assert_allclose(acc, acc_st, rtol=1e-5, atol=1e-6)


def test_dielectric_split_reports_overflow_instead_of_dropping_rays(gpu_api):
    """2^(depth+1) rays per path: beyond max_depth 5 an all-dielectric scene can outgrow the queue; that is an
    error, never a silently darker image"""
    W, H = 32, 18
    objs = gpu_api.scene_sphere_field(300, W, H, mix=(0.0, 1.0, 0.0))
    cam = gpu_api.init_camera(W, H)
    with gpu_api.Scene(objs) as sc:
        with pytest.raises(gpu_api.RtbError, match="SPLIT"):
            sc.render(cam, gpu_api.make_desc(W, H, 0, 1, max_depth=16, dielectric=1))
        with pytest.raises(gpu_api.RtbError, match="SPLIT"):
            sc.render(cam, gpu_api.make_desc(W, H, 0, 1, max_depth=17, dielectric=1))  # rejected up front
        fb, _, ctr = sc.render(cam, gpu_api.make_desc(W, H, 0, 1, max_depth=5, dielectric=1))  # exact worst case fits
        assert ctr.rays > 0


def test_drop_in_render_honours_dielectric_mode_and_total_samples(gpu_api):
    """RenderParams.dielectric_mode and .total_samples used to be ignored by render_ex()"""
    import ctypes as C
    import __graft_entry__ as entry
    pkg = entry.load_package()
    _, host = pkg.load()
    W, H, SPP = 64, 36, 4
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    opt = pkg.abi.Options()
    opt.width, opt.height, opt.samples = W, H, SPP

    def run(**kw):
        rp = pkg.abi.RenderParams()
        host.render_params_default(C.byref(rp))
        rp.num_gpus = 1
        for k, v in kw.items():
            setattr(rp, k, v)
        fb = np.zeros((H, W, 3), np.uint8)
        acc = np.zeros((H, W, 3), np.float32)
        rp.accum_out = acc.ctypes.data_as(C.POINTER(C.c_float))
        before = C.c_longlong.in_dll(host, "ray_count").value
        host.render_ex(fb.ctypes.data, objs.ctypes.data, len(objs), C.byref(cam), C.byref(opt), C.byref(rp))
        return fb, acc, C.c_longlong.in_dll(host, "ray_count").value - before

    fb_s, acc_s, rays_s = run()
    fb_d, acc_d, rays_d = run(dielectric_mode=1)
    assert rays_d >= rays_s  # equal when the scene has no dielectric surface
    # a caller that renders its share of a larger job: mean over total_samples, not over its own count
    fb_half, acc_half, _ = run(total_samples=2 * SPP)
    assert np.array_equal(acc_half, acc_s)
    want = np.minimum(1.0, (acc_s.astype(np.float64) / (2 * SPP)) ** 0.2) * 255.0
    assert np.abs(fb_half.astype(int) - np.floor(want).astype(int)).max() <= 1
    assert fb_half.astype(int).sum() < fb_s.astype(int).sum()


# ---- the named configs at their FULL frame sizes: size-independent properties ------------------------------

@pytest.mark.parametrize("name,W,H,count,mix,depth", [
    ("c2", 1920, 1080, 10000, (0.5, 0.2, 0.2), 8),
    ("c4", 3840, 2160, 10000, (0.2, 0.4, 0.4), 8),
    ("c5", 512, 512, 2000, (0.1, 0.9, 0.0), 64),
])
def test_full_frame_properties_of_the_sphere_configs(gpu_api, name, W, H, count, mix, depth):
    """at BASELINE.json's frame sizes and depths (a few samples instead of hundreds): the sum over samples is
    additive over sample ranges (what the multi-GPU sharding relies on), bit-reproducible, finite and non-negative;
    ray counts add up; every path casts between 1 and depth + 2 rays; the frame is the tonemap of the sums"""
    objs = gpu_api.scene_sphere_field(count, W, H, mix=mix)
    cam = gpu_api.init_camera(W, H)
    with gpu_api.Scene(objs) as sc:
        fb, ab, cab = sc.render(cam, gpu_api.make_desc(W, H, 0, 4, max_depth=depth), want_accum=True)
        _, a, ca = sc.render(cam, gpu_api.make_desc(W, H, 0, 2, max_depth=depth), want_accum=True)
        _, b, cb = sc.render(cam, gpu_api.make_desc(W, H, 2, 4, max_depth=depth), want_accum=True)
        fb2, ab2, cab2 = sc.render(cam, gpu_api.make_desc(W, H, 0, 4, max_depth=depth), want_accum=True)
    assert np.array_equal(ab, ab2) and np.array_equal(fb, fb2) and cab.rays == cab2.rays  # idempotent
    assert np.isfinite(ab).all() and (ab >= 0).all()
    np.testing.assert_allclose(a + b, ab, rtol=1e-5, atol=1e-6)
    assert ca.rays + cb.rays == cab.rays and cab.paths == W * H * 4
    assert 1.0 <= cab.rays / cab.paths <= depth + 2
    want = np.floor(255.0 * np.minimum(1.0, (ab.astype(np.float64) / 4) ** 0.2)).astype(int)
    assert np.abs(fb.astype(int) - want).max() <= 1  # pow() in CUDA vs numpy at truncation boundaries
