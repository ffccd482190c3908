"""Mint the golden fixtures of tests/golden/ from the CPU oracle.  Run once (here, in the authoring
container): python tests/golden/make_golden.py.  The reference ships no golden vectors for this path
(SURVEY §4, §8c), so these are oracle-minted; what pins the oracle to the reference itself is
tests/test_oracle_cpu.py::test_*_match_reference (real reference code through oracle/_ref/ref_host)."""
import importlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import oracle  # noqa: E402

rt = importlib.import_module("raytracing2-fork_b200")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def main():
    meta = {}
    scene = rt.scene_classic_cornell()
    orc = oracle.OracleScene.from_scene(scene)
    cam = rt.make_camera(512, 512, (0.0, 0.0, 15.5))
    u = rt.screenshot_uniforms(scene, cam, spp=64, max_bounce=8, env_light=False)
    tri, dst = orc.first_hit(u, rt.FIRST_HIT_CENTRE)
    meta["config1_first_hit_centre_crc"] = crc(tri)
    meta["config1_first_hit_centre_dst_crc"] = crc(dst)
    for mode, name in ((rt.RNG_REF_PCG, "pcg"), (rt.RNG_PHILOX, "philox")):
        t, _ = orc.first_hit(u, rt.FIRST_HIT_SAMPLE0, rng_mode=mode)
        meta[f"config1_first_hit_sample0_{name}_crc"] = crc(t)
    cam64 = rt.make_camera(64, 64, (0.0, 0.0, 15.5))
    u64 = rt.screenshot_uniforms(scene, cam64, spp=8, max_bounce=8, env_light=False)
    tri64, _ = orc.first_hit(u64, rt.FIRST_HIT_CENTRE)
    np.save(os.path.join(HERE, "classic_first_hit_64.npy"), tri64)
    for mode, name in ((rt.RNG_REF_PCG, "pcg"), (rt.RNG_PHILOX, "philox")):
        img = orc.render_frame(u64, rng_mode=mode)
        meta[f"classic_frame_64_{name}_crc"] = crc(img)
        shot, _ = orc.screenshot(u64, 2, rng_mode=mode)
        np.save(os.path.join(HERE, f"classic_shot_64_{name}.npy"), shot)
    # config 2 mesh, small image: first-hit ids over the textured sphere in the Cornell container
    s2 = rt.scene_textured_sphere(n_quads=224, container="cornell", tex_size=256)
    o2 = oracle.OracleScene.from_scene(s2)
    cam2 = rt.camera_for_box(s2, 240, 135)
    u2 = rt.screenshot_uniforms(s2, cam2, spp=4, max_bounce=6, env_light=False)
    t2, d2 = o2.first_hit(u2, rt.FIRST_HIT_CENTRE)
    meta["config2_first_hit_240x135_crc"] = crc(t2)
    meta["config2_first_hit_240x135_dst_crc"] = crc(d2)
    meta["config2_scene_triangles_crc"] = crc(s2.triangles["a"]) ^ crc(s2.triangles["b"]) ^ crc(s2.triangles["c"])
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
