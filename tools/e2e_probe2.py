"""dev tool: end-to-end time of the drop-in render_scene() on C3, call by call, with a second scene alive
(the situation inside bench.py)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api, abi = pkg.api, pkg.abi
import torch
W, H, spp = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 128
keep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
buf = torch.empty(verts.nbytes, dtype=torch.uint8).pin_memory()
pv = buf.numpy().view(verts.dtype); pv[:] = verts
holder = api.mesh_room(pv, W, H)
cam = api.init_camera(W, H)
_, host = pkg.load()
opt = abi.Options(); opt.width, opt.height, opt.samples = W, H, spp
rp = abi.RenderParams(); host.render_params_default(rp); rp.max_depth = 5
fb = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
other = None
if keep:
    other = api.Scene(holder)
    other.render(cam, api.make_desc(W, H, 0, spp, max_depth=5))
for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host.render_scene(fb.data_ptr(), C.addressof(holder.objects), holder.n, C.byref(cam), C.byref(opt), C.byref(rp))
    t1 = time.perf_counter()
    free, total = torch.cuda.mem_get_info()
    print(f"keep={keep} it{it}: render_scene {1e3*(t1-t0):.1f} ms   device memory used {(total-free)/2**30:.1f} GiB", flush=True)
