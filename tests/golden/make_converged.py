"""Converged reference renders for the PSNR gate, FROM THE UNMODIFIED REFERENCE (oracle/_ref).

The reference's render() is sequential under srand(seed) (its OpenMP build serialises on
rand()), so the samples are spread over independent seeds, one process per seed, and the
per-seed double means are averaged: N seeds x S spp is an N*S-spp render.

    make -C oracle && python tests/golden/make_converged.py

Stored per scene: `mean` (float32, N_FULL seeds), `mean_quarter` (float32, a disjoint set of
N_FULL/4 seeds: the reference's own noise floor), `spp` = [full, quarter], `rays`.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

SEED0 = 1666943821
N_FULL, S = 8, 4096


def scene(name):
    import oracle_lib as ol
    api = ol.pkg.api
    if name == "c1":
        return api.scene_default(96, 54), 96, 54
    if name == "c1_320":  # the reference's default frame (main.c:24-30): 320x180
        return api.scene_default(320, 180), 320, 180
    return api.scene_sphere_field(60, 64, 36, mix=(0.3, 0.4, 0.2)), 64, 36


def one(job):
    name, seed, spp = job
    import oracle_lib as ol
    objs, W, H = scene(name)
    cam = ol.ref_init_camera(W, H)
    mean, ctr = ol.ref_render_mean(objs, cam, W, H, spp, seed=seed, max_depth=5)
    return name, seed, mean, ctr[0]


SCENES = (("c1", "c1_converged_96x54.npz", S), ("dielectric", "dielectric_converged_64x36.npz", S // 2),
          ("c1_320", "c1_converged_320x180.npz", S))


def main():
    """python tests/golden/make_converged.py [scene ...] [--procs N]   (default: every scene, all cores)"""
    args = sys.argv[1:]
    procs = os.cpu_count() or 1
    if "--procs" in args:
        k = args.index("--procs")
        procs = int(args[k + 1])
        del args[k:k + 2]
    todo = [sc for sc in SCENES if not args or sc[0] in args]
    jobs = []
    for name, _, spp in todo:
        for k in range(N_FULL + N_FULL // 4):
            jobs.append((name, SEED0 + 7919 * k, spp))
    with mp.Pool(min(len(jobs), procs)) as pool:
        res = pool.map(one, jobs, chunksize=1)
    for name, fname, spp in todo:
        rs = [r for r in res if r[0] == name]
        full, quarter = rs[:N_FULL], rs[N_FULL:]
        mean = np.mean([r[2] for r in full], axis=0)
        mean_q = np.mean([r[2] for r in quarter], axis=0)
        np.savez_compressed(os.path.join(HERE, fname), mean=mean.astype(np.float32), mean_quarter=mean_q.astype(np.float32),
                            spp=np.array([spp * len(full), spp * len(quarter)]), rays=np.array(sum(r[3] for r in full)),
                            seeds=np.array([r[1] for r in rs]))
        print(fname, mean.shape, spp * len(full), spp * len(quarter), os.path.getsize(os.path.join(HERE, fname)))


if __name__ == "__main__":
    main()
