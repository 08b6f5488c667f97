"""dev tool: C5 (512x512, 64 spp, 90 % dielectric) by max_depth: how fast the deep bounces run (RTB_WF_STREAMS=1..4 for the plane groups)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
os.environ["RTB_QB_PROFILE"] = "0"
import quick_bench as qb
api = qb.api
W = H = 512
src = api.scene_sphere_field(2000, W, H, mix=(0.1, 0.9, 0.0))
for depth in (8, 16, 24, 32, 48, 64):
    qb.run("C5", src, W, H, 64, depth, reps=3, kernel=0)
