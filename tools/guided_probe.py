"""dev tool: guided self-scheduling of the trace kernel's ray batches (RTB_WF_GUIDED) against fixed batches, on C3 at
128 / 32 / 16 spp (the shares of 1 / 4 / 8 GPUs), C2 and C5; production schedule (two streams, no per-kernel events)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ["RTB_QB_PROFILE"] = "0"
import quick_bench as qb
api = qb.api
W, H = 1920, 1080
c3 = api.mesh_room(api.heightfield_mesh(708, 20 * W / H * 0.98), W, H)
c2 = api.scene_sphere_field(10000, W, H)
c5 = api.scene_sphere_field(2000, 512, 512, mix=(0.1, 0.9, 0.0))
for rnd in range(2):
    for g in ("1", "33", "65", "34", "0"):
        os.environ["RTB_WF_GUIDED"] = g
        print(f"--- guided={g}", flush=True)
        for spp in (128, 32, 16):
            qb.run("C3", c3, W, H, spp, 5, reps=3, kernel=6)
        qb.run("C2", c2, W, H, 64, 8, reps=3, kernel=6)
        qb.run("C5", c5, 512, 512, 64, 64, reps=3, kernel=6)
