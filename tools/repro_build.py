import importlib, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
rt = importlib.import_module("raytracing2-fork_b200")
scene = rt.scene_textured_sphere(n_quads=40, container="cornell", tex_size=128)
be = rt.Backend(device=0)
be.upload(scene)
print("built", be.counters())
