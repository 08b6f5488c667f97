"""render() on several GPUs behind the C ABI (csrc/rtb_multi.cu): samples sharded over the ranks, ONE
ncclReduce of the float sums, tonemap on rank 0; scene upload sharded and all-gathered over NVLink.

One-GPU boxes run the group-of-one cases; the two-rank cases need two visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need(gpu_api, n):
    if gpu_api.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def test_group_of_one_equals_the_plain_call(gpu_api):
    """rtb_render_multi on a one-rank group == rtb_render: same bits, same counters"""
    W, H, SPP = 96, 54, 6
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, 2, 2 + SPP, max_depth=5)
    with gpu_api.Scene(objs) as sc:
        fb0, acc0, c0 = sc.render(cam, desc, want_accum=True)
    with gpu_api.Comm.local(1) as comm:
        fb1, acc1, c1 = comm.render_host(objs, cam, desc, want_accum=True)
        with comm.scene(objs) as ms:
            c2 = ms.render(cam, desc, want_counters=True)
    assert np.array_equal(fb0, fb1) and np.array_equal(acc0, acc1)
    assert (c0.rays, c0.paths, c0.prim_tests) == (c1.rays, c1.paths, c1.prim_tests) == (c2.rays, c2.paths, c2.prim_tests)


@pytest.mark.parametrize("n_gpus", [2, 4])
def test_sample_sharded_group_is_the_same_estimator(gpu_api, n_gpus):
    """G GPUs render the SAME global sample indices as one GPU, split G ways: ray counts are equal exactly;
    the float sums differ only by the order of additions; the 8-bit frame by at most one level"""
    _need(gpu_api, n_gpus)
    W, H, SPP = 160, 90, 10  # 10 samples over 4 ranks: uneven shares (2, 3, 2, 3)
    objs = gpu_api.scene_sphere_field(600, W, H, mix=(0.3, 0.3, 0.3))
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, 0, SPP, max_depth=8)
    with gpu_api.Scene(objs) as sc:
        fb0, acc0, c0 = sc.render(cam, desc, want_accum=True)
    with gpu_api.Comm.local(n_gpus) as comm:
        assert comm.size == n_gpus and comm.local_ranks == n_gpus
        fb1, acc1, c1 = comm.render_host(objs, cam, desc, want_accum=True)
    assert c1.rays == c0.rays and c1.paths == c0.paths == W * H * SPP
    np.testing.assert_allclose(acc1, acc0, rtol=2e-5, atol=1e-5)
    assert np.abs(fb1.astype(int) - fb0.astype(int)).max() <= 1


def test_sharded_mesh_upload_builds_the_same_scene(gpu_api):
    """every rank uploads 1/G of the triangles, the marshalled records are all-gathered over NVLink:
    the frame equals the one-GPU frame of the same job (mesh large enough to take the sharded path)"""
    _need(gpu_api, 2)
    W, H, SPP = 128, 72, 4
    verts = gpu_api.heightfield_mesh(96, 20 * W / H * 0.98)  # 18 432 triangles > 2 x 4096
    holder = gpu_api.mesh_room(verts, W, H)
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, 0, SPP, max_depth=5)
    with gpu_api.Scene(holder) as sc:
        fb0, acc0, c0 = sc.render(cam, desc, want_accum=True)
    with gpu_api.Comm.local(2) as comm:
        fb1, acc1, c1 = comm.render_host(holder, cam, desc, want_accum=True)
        # each rank's scene alone reproduces the one-GPU nearest hits (the gathered arrays are complete)
        with comm.scene(holder) as ms:
            c2 = ms.render(cam, desc, want_counters=True)
    assert c1.rays == c0.rays == c2.rays and c1.prim_tests == c0.prim_tests
    np.testing.assert_allclose(acc1, acc0, rtol=2e-5, atol=1e-5)
    assert np.abs(fb1.astype(int) - fb0.astype(int)).max() <= 1


def test_drop_in_render_uses_the_whole_box(gpu_api):
    """render_ex() with RenderParams.num_gpus = 2 (what $RTB_NUM_GPUS gives the unchanged main.c)"""
    _need(gpu_api, 2)
    import ctypes as C
    import __graft_entry__ as entry
    pkg = entry.load_package()
    _, host = pkg.load()
    W, H, SPP = 96, 54, 8
    objs = gpu_api.scene_default(W, H)
    cam = gpu_api.init_camera(W, H)
    opt = pkg.abi.Options()
    opt.width, opt.height, opt.samples = W, H, SPP
    frames = []
    for g in (1, 2):
        rp = pkg.abi.RenderParams()
        host.render_params_default(C.byref(rp))
        rp.num_gpus = g
        fb = np.zeros((H, W, 3), np.uint8)
        host.render_ex(fb.ctypes.data, objs.ctypes.data, len(objs), C.byref(cam), C.byref(opt), C.byref(rp))
        frames.append(fb)
    assert np.abs(frames[0].astype(int) - frames[1].astype(int)).max() <= 1


_RANK_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
idt = torch.zeros(api.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, src=0)
comm = api.Comm.rank(bytes(idt.cpu().numpy().tobytes()), rank, world, local)
W, H, SPP = 128, 72, 6
verts = api.heightfield_mesh(96, 20 * W / H * 0.98)
holder = api.mesh_room(verts, W, H)
cam = api.init_camera(W, H)
desc = api.make_desc(W, H, 0, SPP, max_depth=5)
fb, acc, ctr = comm.render_host(holder, cam, desc, want_accum=True)
if rank == 0:
    with api.Scene(holder, device=local) as sc:
        fb0, acc0, c0 = sc.render(cam, desc, want_accum=True)
    assert ctr.rays == c0.rays and ctr.paths == W * H * SPP, (ctr.rays, c0.rays)
    np.testing.assert_allclose(acc, acc0, rtol=2e-5, atol=1e-5)
    assert np.abs(fb.astype(int) - fb0.astype(int)).max() <= 1
    print("OK", world, ctr.rays)
dist.barrier()
comm.close()
dist.destroy_process_group()
"""


def test_one_process_per_gpu_group(gpu_api, tmp_path):
    """the torchrun form (rtb_comm_create_rank, id broadcast by the launcher): two processes, two GPUs,
    rtb_render_multi == the one-GPU frame of the same job"""
    _need(gpu_api, 2)
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "rank_worker.py"
    script.write_text(_RANK_WORKER.format(root=root))
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "OK 2" in r.stdout


def test_sharded_upload_of_double_precision_vertices(gpu_api):
    """a mesh moved with apply_matrix (main.c:140-147) has genuine double coordinates: every rank finds out on its
    own share, the flag is max-reduced inside the gather group, and the doubles are all-gathered too"""
    _need(gpu_api, 2)
    W, H, SPP = 128, 72, 4
    verts = gpu_api.heightfield_mesh(96, 20 * W / H * 0.5)
    a = 0.37
    m = np.array([[np.cos(a), 0, np.sin(a), 0.1234567], [0, 1.1, 0, 1.0 / 3.0], [-np.sin(a), 0, np.cos(a), -0.7], [0, 0, 0, 1]])
    gpu_api.apply_matrix(verts, m)
    holder = gpu_api.mesh_room(verts, W, H)
    cam = gpu_api.init_camera(W, H)
    desc = gpu_api.make_desc(W, H, 0, SPP, max_depth=5)
    with gpu_api.Scene(holder) as sc:
        assert sc.info.double_triangles == 1
        fb0, acc0, c0 = sc.render(cam, desc, want_accum=True)
    with gpu_api.Comm.local(2) as comm:
        fb1, acc1, c1 = comm.render_host(holder, cam, desc, want_accum=True)
    assert c1.rays == c0.rays and c1.prim_tests == c0.prim_tests
    np.testing.assert_allclose(acc1, acc0, rtol=2e-5, atol=1e-5)


def test_more_gpus_than_the_box_has_is_an_error(gpu_api):
    with pytest.raises(gpu_api.RtbError, match="device ordinal"):
        gpu_api.Comm.local(gpu_api.device_count() + 1)
