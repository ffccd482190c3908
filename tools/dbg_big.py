import importlib, sys, numpy as np
sys.path[:0] = ['/root/repo', '/root/repo/tests']
rt = importlib.import_module("raytracing2-fork_b200")
scene = rt.scene_big_sphere(n_quads=2236)
be = rt.Backend(device=0); be.upload(scene)
print(be.counters())
rng = np.random.default_rng(21)
o = rng.uniform(-4.9, 4.9, size=(200000, 3)).astype(np.float32)
c = np.array([0, -1, 0], np.float32)
far = np.linalg.norm(o - c, axis=1) > 3.3
oc = o[far]; dc = c - oc; dist = np.linalg.norm(dc, axis=1, keepdims=True); dc = (dc / dist).astype(np.float32)
t, d, u, v = be.trace_rays(oc, dc)
bad = np.where(t < 14)[0]
print("rays", len(oc), "bad", len(bad))
for b in bad[:10]:
    print(t[b], oc[b], dc[b], d[b], dist[b, 0], "hit point", oc[b] + dc[b] * d[b])
