"""dev tool: one process drives N GPUs (rtb_comm_create_local): time rtb_render_multi end to end on C3.
RTB_TIMING=1 prints the per-rank phases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
W, H = 1920, 1080
src = api.mesh_room(api.heightfield_mesh(708, 20 * W / H * 0.98), W, H)
cam = api.init_camera(W, H)
desc = api.make_desc(W, H, 0, spp, max_depth=5)
fb = np.zeros((H, W, 3), np.uint8)
with api.Comm.local(n) as comm:
    for k in range(4):
        t0 = time.perf_counter()
        _, _, ctr = comm.render_host(src, cam, desc, want_counters=(k == 0), fb=fb)
        dt = (time.perf_counter() - t0) * 1e3
        print(f"local x{n}: call {k}: {dt:.2f} ms" + (f"  rays {ctr.rays} -> {ctr.rays / dt / 1e3:.0f} Mrays/s e2e" if ctr is not None else ""), flush=True)
