"""dev tool: load_obj on the C3 mesh written out as OBJ (239 MB), by thread count; host only"""
import ctypes as C, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api; abi = pkg.abi
_, host = api.load()
W, H = 1920, 1080
verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
p = os.path.join(tempfile.gettempdir(), "rtb_hf708.obj")
t0 = time.time(); api.write_obj(p, verts); print(f"scene_write_obj {time.time() - t0:.2f} s, {os.path.getsize(p) / 1e6:.0f} MB, {os.cpu_count()} cores", flush=True)
for th in (1, 2, 4, 8, 16, 32, 0):
    best = 1e9
    for k in range(3):
        mesh = abi.TriangleMesh()
        t0 = time.time(); ok = host.load_obj_ex(os.fsencode(p), C.byref(mesh), th, 0); dt = time.time() - t0
        n = mesh.num_triangles; host.free_mesh(C.byref(mesh)); best = min(best, dt)
    print(f"load_obj_ex threads {th or 'default'}: {best * 1e3:.0f} ms, {n} triangles, ok={bool(ok)}", flush=True)
os.remove(p)
