"""The drop-in claim of INTEGRATION.md section 2, guarded: the reference's UNCHANGED callers --
/root/reference/main.c (call site of render(), main.c:429) and /root/reference/test.c -- compiled with the
reference's own flags (Makefile:2) and headers, link against libraytracer_b200.so instead of raytracer.o.

oracle/Makefile does the compiling where /root/reference exists and leaves the binaries in oracle/_ref/
(git-ignored, shipped to the GPU box); nothing here reads /root/reference at run time.
"""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
MAIN_B200 = os.path.join(REF_DIR, "ref_main_b200")
TEST_B200 = os.path.join(REF_DIR, "ref_test_b200")

needs_bins = pytest.mark.skipif(not (os.path.exists(MAIN_B200) and os.path.exists(TEST_B200)),
                                reason="oracle/_ref binaries not built (make -C oracle, needs /root/reference)")


@needs_bins
def test_unchanged_callers_resolve_every_symbol_in_the_product_library():
    """no undefined symbol is left for raytracer.o: ldd/nm see the product libraries only"""
    for exe in (MAIN_B200, TEST_B200):
        needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True, check=True).stdout
        assert "libraytracer_b200.so" in needed
        r = subprocess.run(["ldd", "-r", exe], capture_output=True, text=True)
        assert "undefined symbol" not in r.stdout + r.stderr, r.stdout + r.stderr
    # main.c's imports: render, init_camera, the two counters
    # (the two counters are data: the executable holds copy relocations of the library's definitions)
    syms = subprocess.run(["nm", "-D", MAIN_B200], capture_output=True, text=True, check=True).stdout
    for s in ("render", "init_camera", "ray_count", "intersection_test_count"):
        assert re.search(rf"\b{s}\b", syms), s


@needs_bins
def test_reference_unit_test_runs_against_the_product_library():
    """test.c: vec3_cross passes (test.c:63); calculate_surface_normal reproduces upstream's own
    FAIL at test.c:78 (the reference computes the negated normal) -- same as raytracer.o"""
    r = subprocess.run([TEST_B200], capture_output=True, text=True, timeout=60)
    out = r.stdout + r.stderr
    assert r.returncode == 0
    assert re.search(r"OK.*test\.c:0063", out), out
    assert re.search(r"FAIL.*test\.c:0078", out), out


@needs_bins
def test_reference_main_fails_loudly_without_a_gpu(tmp_path):
    """no CPU fallback: without a CUDA device the unchanged main.c exits non-zero with the CUDA error
    (the reference's own error convention, main.c:42,192,418)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([MAIN_B200, "-w", "32", "-h", "18", "-s", "1", "-o", str(tmp_path / "x.png")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "render" in r.stderr and "cuda" in r.stderr.lower(), r.stderr
    assert not os.path.exists(tmp_path / "x.png")


@needs_bins
@pytest.mark.gpu
def test_reference_main_renders_on_the_gpu(gpu_api, tmp_path):
    """the unchanged reference program, linked against the product, renders its default scene on the B200:
    PNG written, ray count printed by main.c:436 equal to the one the C ABI reports for the same call"""
    W, H, S = 96, 54, 3
    out = tmp_path / "ref_main.png"
    r = subprocess.run([MAIN_B200, "-w", str(W), "-h", str(H), "-s", str(S), "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert f"{W} x {H} ({W * H}) pixels" in r.stdout
    rays = int(re.search(r"cast (\d+) rays", r.stdout).group(1))
    data = out.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    assert int.from_bytes(data[16:20], "big") == W and int.from_bytes(data[20:24], "big") == H
    # same scene, camera, seed and sample range through the Python binding of the same library
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    with gpu_api.Scene(objs) as sc:
        _, _, ctr = sc.render(cam, gpu_api.make_desc(W, H, 0, S, max_depth=5))
    assert rays == ctr.rays
