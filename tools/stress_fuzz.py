"""Repeat one fuzz scene many times per kernel-switch environment and count frames that differ from the oracle
(debugging aid for intermittent failures; run on the GPU box)."""
import importlib, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
rt = importlib.import_module("raytracing2-fork_b200")
import oracle, scenes

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
scene, u = scenes.random_scene(seed)
orc = oracle.OracleScene.from_scene(scene)
ref = {m: orc.render_frame(u, rng_mode=m) for m in (rt.RNG_REF_PCG, rt.RNG_PHILOX)}
envs = [None]   # the environment comes from the caller: one process per configuration survives a crash of another
cn = oracle.OrcCounters()
orc.render_frame(u, rng_mode=rt.RNG_PHILOX, counters=cn)
print("oracle segments", cn.segments, "paths", cn.paths)
for env in envs:
    bad = {0: 0, 1: 0}
    zeros = 0
    first = None
    for r in range(reps):
        for mode in (rt.RNG_REF_PCG, rt.RNG_PHILOX):
            be = rt.Backend(device=0, rng_mode=mode)
            be.upload(scene)
            be.render_frame(u)
            img = be.read_frame()
            d = img.view(np.uint32) != ref[mode].view(np.uint32)
            if d.any():
                bad[mode] += 1
                px = np.argwhere(d.any(axis=2))
                if first is None:
                    first = (mode, int(d.sum()), len(px), px[:4].tolist(), float(img[tuple(px[0])][0]), float(ref[mode][tuple(px[0])][0]),
                             "segments", be.counters()["segments"])
            if d.any() and mode == rt.RNG_PHILOX and bad[mode] <= 3:
                again = []
                for _ in range(6):
                    be.render_frame(u)
                    again.append(int((be.read_frame().view(np.uint32) != ref[mode].view(np.uint32)).sum()))
                print("   same context, 6 more renders, differing floats:", again, flush=True)
            if mode == rt.RNG_PHILOX:
                o = np.random.default_rng(seed).uniform(-5, 5, (2000, 3)).astype(np.float32)
                dd = np.random.default_rng(seed + 1).normal(size=(2000, 3)).astype(np.float32)
                be.trace_rays(o, dd)
            be.close()
    print({k: v for k, v in os.environ.items() if k.startswith("RT_")}, "bad frames pcg/philox:", bad[0], bad[1], "of", reps, "first:", first, flush=True)
