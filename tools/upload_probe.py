"""dev tool: scene create (upload + marshal + BVH build) of C3's mesh from pageable memory, narrowed (60 B per
triangle, converted by the staging threads) against raw (120 B), by staging-thread count; pinned for reference"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api
import torch
W, H = 1920, 1080
verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
buf = torch.empty(verts.nbytes, dtype=torch.uint8).pin_memory()
pv = buf.numpy().view(verts.dtype); pv[:] = verts
def run(label, v):
    holder = api.mesh_room(v, W, H)
    ts = []
    for it in range(10):
        t0 = time.perf_counter(); sc = api.Scene(holder); t1 = time.perf_counter()
        sc.close()
        ts.append(1e3 * (t1 - t0))
    print(f"{label}: create ms " + " ".join(f"{t:.2f}" for t in ts) + f"   best {min(ts[1:]):.2f}", flush=True)
run("pinned (raw, one DMA)", pv)
for narrow in ("1", "0"):
    for th in ("6", "8", "12", "16"):
        os.environ["RTB_UPLOAD_NARROW"] = narrow
        os.environ["RTB_UPLOAD_THREADS"] = th
        run(f"pageable narrow={narrow} threads={th}", verts)
