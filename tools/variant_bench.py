"""dev tool: time trace-kernel variants on C3/C2 (one GPU).  tune word = variant<<16 | node_exit<<8 | refill"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import quick_bench as qb
api = qb.api

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="22,54,86,150,182")
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--which", default="c3,c2")
    ap.add_argument("--node-exit", default="16")
    ap.add_argument("--refill", default="8")
    a = ap.parse_args()
    W, H = 1920, 1080
    srcs = {}
    if "c3" in a.which:
        srcs["C3"] = (api.mesh_room(api.heightfield_mesh(708, 20 * W / H * 0.98), W, H), max(1, int(128 * a.scale)), 5)
    if "c2" in a.which:
        srcs["C2"] = (api.scene_sphere_field(10000, W, H), max(1, int(256 * a.scale)), 8)
    if "c5" in a.which:
        srcs["C5"] = (api.scene_sphere_field(2000, 512, 512, mix=(0.1, 0.9, 0.0)), max(1, int(64 * a.scale)), 64)
    for name, (src, spp, depth) in srcs.items():
        w, h = (512, 512) if name == "C5" else (W, H)
        for v in [int(x) for x in a.variants.split(",")]:
            for ne in [int(x) for x in a.node_exit.split(",")]:
                for rf in [int(x) for x in a.refill.split(",")]:
                    qb.run(name, src, w, h, spp, depth, reps=3, kernel=6, tune=(v << 16) | (ne << 8) | rf)
