"""Scenes shared by the CPU and GPU parity tests and by tests/golden/make_golden_refshader.py."""
import importlib

import numpy as np

rt = importlib.import_module("raytracing2-fork_b200")


def material_zoo():
    """CHECKER, GLASS, partial-smoothness SPECULAR, edge highlight, GLASS_HIGHLIGHT (magenta in trace), a light
    slab and no container, so misses reach the procedural sky (compute.glsl:216-273, 521-546)."""
    s = rt.Scene()
    red = s.add_fixed_materials()
    glass = s.add_glass((0.9, 0.95, 1.0), 1.5)
    checker = s.add_checker(2.0)
    metal = s.add_specular((0.8, 0.6, 0.2), (1, 1, 1), 0.7, 0.5)
    m = np.zeros(1, dtype=rt.MATERIAL)
    m["color"] = (0.2, 0.9, 0.3, 0); m["materialType"] = rt.MAT_DIFFUSE; m["isEdgeHighlight"] = 1; m["textureIndex"] = -1
    edge = s.add_material(m)
    m2 = np.zeros(1, dtype=rt.MATERIAL)
    m2["color"] = (1, 1, 0, 0); m2["materialType"] = rt.MAT_GLASS_HIGHLIGHT; m2["textureIndex"] = -1
    gh = s.add_material(m2)
    s.add_cube((0, -1.5, 0), (12, 0.2, 12), (0, 0, 0), checker)
    s.add_cube((-2.5, 0, 0), (1.5, 1.5, 1.5), (0.2, 0.5, 0.1), glass)
    s.add_cube((0, 0, -1), (1.5, 1.5, 1.5), (0.0, 0.8, 0.3), metal)
    s.add_cube((2.5, 0, 0), (1.5, 1.5, 1.5), (0.4, 0.1, 0.0), edge)
    s.add_cube((0, 1.5, -3), (1, 1, 1), (0, 0, 0), gh)
    s.add_cube((0, 3.5, 0), (2, 0.1, 2), (0, 0, 0), red + 3)
    return s


def zoo_camera(width=96, height=64):
    return rt.make_camera(width, height, (0.0, 1.0, 12.0), pitch=0.05)


# ---- the cases pinned against the reference's own shader source (tests/golden/refshader_*.npy)
def refshader_cases():
    """name -> (scene, uniforms).  RT_RNG_REF_PCG frames (the shader's own random stream) and previews."""
    cases = {}
    classic = rt.scene_classic_cornell()
    cam = rt.make_camera(64, 64, (0.0, 0.0, 15.5))
    cases["classic_64x64_spp16_d8"] = (classic, rt.screenshot_uniforms(classic, cam, spp=16, max_bounce=8, env_light=False))
    camd = rt.make_camera(48, 48, (0.0, 0.0, 15.5), defocus=0.05)
    u = rt.screenshot_uniforms(classic, camd, spp=6, max_bounce=5, env_light=False)
    u["frameIndex"] = 3
    cases["classic_defocus_48x48_frame3"] = (classic, u)
    up = rt.interactive_uniforms(classic, cam)
    cases["classic_preview_64x64"] = (classic, up)
    ups = up.copy(); ups["basicShadingShadow"] = 1
    cases["classic_preview_shadow_64x64"] = (classic, ups)
    zoo = material_zoo()
    cases["zoo_96x64_spp8_d10_env"] = (zoo, rt.screenshot_uniforms(zoo, zoo_camera(), spp=8, max_bounce=10, env_light=True))
    cases["zoo_preview_96x64"] = (zoo, rt.interactive_uniforms(zoo, zoo_camera()))
    sph = rt.scene_textured_sphere(n_quads=24, container="cornell", tex_size=64)
    cams = rt.camera_for_box(sph, 80, 60)
    cases["sphere_cornell_80x60_spp8_d8"] = (sph, rt.screenshot_uniforms(sph, cams, spp=8, max_bounce=8, env_light=False))
    mir = rt.scene_textured_sphere(n_quads=16, container="mirror", tex_size=32)
    camm = rt.camera_for_box(mir, 64, 48)
    cases["sphere_mirror_64x48_spp4_d12"] = (mir, rt.screenshot_uniforms(mir, camm, spp=4, max_bounce=12, env_light=False))
    # createDiverseCornellBox (rayTracing.cpp:1071-1118): glass, mirror, checker and metal cubes in the classic room
    div = rt.Scene()
    red = div.add_fixed_materials()
    glass = div.add_glass((1.0, 1.0, 1.0), 1.5)
    checker = div.add_checker(1.0)
    metal = div.add_specular((0.9, 0.7, 0.3), (1, 1, 1), 0.85, 0.6)
    div.create_diverse_cornell_box(10.0, red, red + 1, red + 2, red + 3, glass, red + 4, checker, metal)
    cases["diverse_64x64_spp8_d12"] = (div, rt.screenshot_uniforms(div, cam, spp=8, max_bounce=12, env_light=False))
    # addSkyLightPlane (rayTracing.cpp:388-432; duplicate triangles on purpose) over a sphere, sky on misses
    sky = rt.Scene()
    sky.set_procedural_texture(0, 32)
    tex = sky.add_textured(0)
    sky.add_displaced_sphere(12, (0.0, 0.0, 0.0), 3.0, 0.05, tex)
    r2 = sky.add_fixed_materials()
    sky.add_sky_light_plane(r2 + 3)
    camk = rt.make_camera(72, 48, (0.0, 2.0, 14.0), pitch=0.1)
    cases["sky_sphere_72x48_spp6_d6_env"] = (sky, rt.screenshot_uniforms(sky, camk, spp=6, max_bounce=6, env_light=True))
    # addSideLitCornellBox (rayTracing.cpp:690-847), rotated variant
    side = rt.Scene()
    side.set_procedural_texture(0, 32)
    t2 = side.add_textured(0)
    side.add_displaced_sphere(10, (0.0, 0.0, 0.0), 3.0, 0.05, t2)
    r3 = side.add_fixed_materials()
    d = rt.defaults()
    side.add_side_lit_cornell_box(d.cornell_light_size, d.cornell_padding, r3 + 3, r3 + 2, True)
    camsd = rt.camera_for_box(side, 64, 48)
    cases["sphere_sidelit_64x48_spp6_d8"] = (side, rt.screenshot_uniforms(side, camsd, spp=6, max_bounce=8, env_light=False))
    return cases


def refshader_big_cases():
    """Full-size cases pinned by CRC only (tests/golden/refshader.json → "crc_only"): BASELINE config 1 exactly
    (512x512, one 64-spp frame, depth 8, classic Cornell box) and the config-2 scene (100 368 triangles, depth 20)
    at 240x135 x 8 spp.  Minutes on the CPU for the reference shader, so not regenerated by the CPU suite."""
    cases = {}
    classic = rt.scene_classic_cornell()
    cam = rt.make_camera(512, 512, (0.0, 0.0, 15.5))
    cases["config1_512x512_spp64_d8"] = (classic, rt.screenshot_uniforms(classic, cam, spp=64, max_bounce=8, env_light=False))
    sph = rt.scene_textured_sphere()
    cams = rt.camera_for_box(sph, 240, 135)
    cases["config2_scene_240x135_spp8_d20"] = (sph, rt.screenshot_uniforms(sph, cams, spp=8, max_bounce=20, env_light=False))
    return cases


def random_scene(seed: int):
    """A seeded random triangle soup with every material type, up to three small textures of 1, 3 and 4 channels
    (non-power-of-two sizes), a random camera looking at it, random uniforms.  Returns (scene, uniforms)."""
    rng = np.random.default_rng(seed)
    s = rt.Scene()
    ntex = int(rng.integers(0, 4))
    for t in range(ntex):
        ch = (1, 3, 4)[t % 3]
        w, h = int(rng.integers(3, 20)), int(rng.integers(3, 20))
        px = rng.integers(0, 256, size=(h, w, ch), dtype=np.uint8)
        s.set_texture(t, px[:, :, 0] if ch == 1 else px)
    mats = []
    for _ in range(int(rng.integers(2, 5))):
        mats.append(s.add_diffuse(*rng.uniform(0.1, 1.0, 3)))
    mats.append(s.add_light(*rng.uniform(0.5, 1.0, 3), float(rng.uniform(2.0, 12.0))))
    mats.append(s.add_specular(rng.uniform(0.2, 1.0, 3), (1, 1, 1), float(rng.uniform(0, 1)), float(rng.uniform(0, 1))))
    mats.append(s.add_specular((1, 1, 1), (1, 1, 1), 1.0, 1.0))
    mats.append(s.add_checker(float(rng.uniform(0.5, 3.0))))
    mats.append(s.add_glass(rng.uniform(0.7, 1.0, 3), float(rng.uniform(1.1, 1.8))))
    for t in range(ntex):
        mats.append(s.add_textured(t))
    mats.append(s.add_textured(ntex))      # texture index out of range: black (compute.glsl:349)
    m = np.zeros(1, dtype=rt.MATERIAL)
    m["color"] = (*rng.uniform(0.2, 1.0, 3), 0); m["materialType"] = rt.MAT_DIFFUSE; m["isEdgeHighlight"] = 1; m["textureIndex"] = -1
    mats.append(s.add_material(m))
    m2 = np.zeros(1, dtype=rt.MATERIAL)
    m2["color"] = (1, 1, 0, 0); m2["materialType"] = rt.MAT_GLASS_HIGHLIGHT; m2["textureIndex"] = -1
    mats.append(s.add_material(m2))
    n = int(rng.integers(20, 160))
    tris = np.zeros(n, dtype=rt.TRIANGLE)
    centre = rng.uniform(-3, 3, size=(n, 3))
    for k, name in enumerate(("a", "b", "c")):
        tris[name][:, :3] = (centre + rng.normal(scale=float(rng.uniform(0.3, 1.5)), size=(n, 3))).astype(np.float32)
    for name in ("aTex", "bTex", "cTex"):
        tris[name] = rng.uniform(-1.5, 2.5, size=(n, 2)).astype(np.float32)
    tris["materialIndex"] = rng.choice(mats, size=n)
    s.add_triangles(tris)
    # a floor and a big light so that most paths end somewhere
    s.add_cube((0, -4.5, 0), (14, 0.3, 14), (0, 0, 0), mats[0])
    s.add_cube((0, 6.0, 0), (5, 0.2, 5), (0, 0, 0), mats[int(np.flatnonzero(s.materials["materialType"][mats] == rt.MAT_LIGHT)[0])])
    pos = rng.uniform(-1, 1, 3) * (2.0, 1.5, 2.0) + (0.0, 0.5, 11.0)
    cam = rt.make_camera(int(rng.integers(24, 56)), int(rng.integers(16, 40)), tuple(float(x) for x in pos),
                         pitch=float(rng.uniform(-0.15, 0.15)), yaw=float(np.pi / 2 + rng.uniform(-0.2, 0.2)),
                         defocus=float(rng.choice([0.0, 0.0, 0.03])))
    if rng.random() < 0.25:
        u = rt.interactive_uniforms(s, cam)
        u["basicShadingShadow"] = int(rng.integers(0, 2))
    else:
        u = rt.screenshot_uniforms(s, cam, spp=int(rng.integers(1, 7)), max_bounce=int(rng.integers(1, 14)),
                                   env_light=bool(rng.integers(0, 2)))
        u["frameIndex"] = int(rng.integers(0, 1000))
    return s, u
