"""dev tool: time the render kernel on the benchmark scenes (one GPU)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); api = pkg.api

def run(name, source, W, H, spp, depth, reps=2, kernel=0, tune=0, planes=0, tune2=0):
    cam = api.init_camera(W, H)
    t0 = time.time()
    sc = api.Scene(source)
    info = sc.info
    t_build = time.time() - t0
    desc = api.make_desc(W, H, 0, spp, max_depth=depth, kernel=kernel, tune=tune, planes=planes, tune2=tune2,
                         profile=int(os.environ.get("RTB_QB_PROFILE", "1")))  # 0: no per-kernel events, no node counts
    best = None
    for r in range(reps):
        fb, acc, ctr = sc.render(cam, desc, want_accum=False)
        if best is None or ctr.gpu_ms < best.gpu_ms: best = ctr
    c = best
    print(f"{name} k{kernel} t{tune} s{tune2} p{planes}: {W}x{H}x{spp} d{depth} prims={info.n_bvh_prims} big={info.n_big_prims} nodes={info.n_bvh_nodes} "
          f"depth={info.bvh_depth} build={info.build_ms:.1f}ms(wall {t_build*1e3:.0f}) gpu={c.gpu_ms:.1f}ms "
          f"rays={c.rays} Mrays/s={c.rays/c.gpu_ms/1e3:.1f} Mpaths/s={c.paths/c.gpu_ms/1e3:.1f} "
          f"nodes/ray={c.node_visits/max(1,c.rays_intersected):.1f} tests/ray={c.prim_tests/max(1,c.rays_intersected):.1f}", flush=True)
    sc.close()
    return fb

if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--which", default="c1,c2,c3,c5"); ap.add_argument("--kernels", default="1,2"); ap.add_argument("--tunes", default="0"); ap.add_argument("--planes", type=int, default=0); ap.add_argument("--tunes2", default="0")
    a = ap.parse_args()
    which = a.which.split(",")
    for kern, tune, tune2 in [(int(k), int(t), int(t2)) for k in a.kernels.split(",") for t in a.tunes.split(",") for t2 in a.tunes2.split(",")]:
        if "c1" in which:
            run("C1", api.scene_default(320, 180), 320, 180, 50, 5, kernel=kern, tune=tune, planes=a.planes, tune2=tune2)
        if "c2" in which:
            W, H = 1920, 1080; spp = max(1, int(256 * a.scale))
            run("C2", api.scene_sphere_field(10000, W, H), W, H, spp, 8, kernel=kern, tune=tune, planes=a.planes, tune2=tune2)
        if "c3" in which:
            W, H = 1920, 1080; spp = max(1, int(128 * a.scale))
            verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
            run("C3", api.mesh_room(verts, W, H), W, H, spp, 5, kernel=kern, tune=tune, planes=a.planes, tune2=tune2)
        if "c5" in which:
            W, H = 512, 512; spp = max(1, int(64 * a.scale))
            run("C5", api.scene_sphere_field(2000, W, H, mix=(0.1, 0.9, 0.0)), W, H, spp, 64, kernel=kern, tune=tune, planes=a.planes, tune2=tune2)
