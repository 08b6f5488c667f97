"""dev tool: where does the end-to-end time of the C3 step go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api
import torch
W, H, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 128
verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
buf = torch.empty(verts.nbytes, dtype=torch.uint8).pin_memory()
pv = buf.numpy().view(verts.dtype); pv[:] = verts
for label, v in (("pageable", verts), ("pinned", pv)):
    holder = api.mesh_room(v, W, H)
    cam = api.init_camera(W, H)
    desc = api.make_desc(W, H, 0, spp, max_depth=5)
    for it in range(3):
        t0 = time.perf_counter(); sc = api.Scene(holder); t1 = time.perf_counter()
        bms = sc.info.build_ms
        fb, acc, ctr = sc.render(cam, desc); t2 = time.perf_counter()
        sc.close(); t3 = time.perf_counter()
        print(f"{label} it{it}: create {1e3*(t1-t0):.1f} ms (build_ms {bms:.1f}) render wall {1e3*(t2-t1):.1f} ms gpu_ms {ctr.gpu_ms:.1f} close {1e3*(t3-t2):.1f} ms", flush=True)
