"""CPU tests (no GPU): the C-ABI library loads and exports every symbol of include/rt_b200.h, host-side
logic (row / frame sharding, camera and controls, scene assembly, PNG out), and the N>1 exchange logic
with a world_size-2 gloo group (partial frame sums from the oracle, reduced / gathered like the
library does over NCCL)."""
import ctypes as C
import importlib
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

import oracle

rt = importlib.import_module("raytracing2-fork_b200")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_loads_and_exports_every_declared_symbol():
    lib = rt.backend_lib()
    header = open(os.path.join(REPO, "include", "rt_b200.h")).read()
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    declared -= {"rt_ctx"}
    assert declared, "no declarations found"
    assert set(rt.ABI_SYMBOLS) == declared, (sorted(declared - set(rt.ABI_SYMBOLS)), sorted(set(rt.ABI_SYMBOLS) - declared))
    for name in sorted(declared):
        assert hasattr(lib, name), f"librt_b200.so does not export {name}"
    assert b"sm_100a" in lib.rt_version()


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device rt_create must fail loudly (this container has none)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rt.BackendError) as e:
        rt.Backend(device=0)
    assert "no CUDA device" in str(e.value) or "rt_create failed (2)" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(REPO, "raytracing2-fork_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", "Makefile")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "rt_oracle" not in text and "import oracle" not in text, f


def test_wire_struct_sizes_match_header():
    header = open(os.path.join(REPO, "include", "rt_b200.h")).read()
    assert "80 bytes" in header and "96 bytes" in header and "192 bytes" in header
    assert rt.TRIANGLE.itemsize == 80 and rt.MATERIAL.itemsize == 96 and rt.UNIFORMS.itemsize == 192
    assert rt.UNIFORMS.fields["frameIndex"][1] == 60 and rt.UNIFORMS.fields["cameraPos"][1] == 64
    assert rt.UNIFORMS.fields["defocusDiskUp"][1] == 176
    assert rt.MATERIAL.fields["materialType"][1] == 72 and rt.MATERIAL.fields["isEdgeHighlight"][1] == 80
    assert rt.TRIANGLE.fields["materialIndex"][1] == 72


def test_split_rows_and_frames_partition():
    for h, band, world in [(1080, 8, 8), (2160, 8, 3), (40, 4, 3), (7, 0, 2), (5, 16, 4)]:
        seen = np.zeros(h, dtype=int)
        for r in range(world):
            rows = rt.split_rows(h, band, r, world)
            assert np.all(np.diff(rows) > 0)
            seen[rows] += 1
        assert np.all(seen == 1)
        sizes = [rt.split_rows(h, band, r, world).size for r in range(world)]
        assert max(sizes) - min(sizes) <= (band if band > 0 else 8)
    for frames, world in [(64, 8), (10, 4), (3, 8), (1, 1)]:
        allf = np.concatenate([rt.split_frames(frames, r, world) for r in range(world)])
        assert sorted(allf.tolist()) == list(range(frames))
        assert all((rt.split_frames(frames, r, world) % world == r).all() for r in range(world))


def test_reference_defaults_and_uniform_blocks():
    d = rt.defaults()
    assert (d.scr_width, d.scr_height, d.max_bounce_count, d.screenshot_frames) == (1000, 1000, 10, 10)
    assert (d.screenshot_rays_per_pixel, d.screenshot_max_bounce_count) == (64, 20)
    assert abs(d.hfov - np.pi / 6) < 1e-6 and abs(d.yaw - np.pi / 2) < 1e-6 and d.focus_distance == 20.0
    s = rt.scene_classic_cornell()
    cam = rt.make_camera(1000, 1000, tuple(d.camera_pos))
    ui = rt.interactive_uniforms(s, cam, 5.7, frame_index=9)
    assert ui["basicShading"][0] == 1 and ui["maxBounceCount"][0] == 10 and ui["numRaysPerPixel"][0] == 5
    assert ui["numTriangles"][0] == 38 and ui["frameIndex"][0] == 9 and ui["environmentalLight"][0] == 0
    us = rt.screenshot_uniforms(s, cam)
    assert us["basicShading"][0] == 0 and us["maxBounceCount"][0] == 20 and us["numRaysPerPixel"][0] == 64
    assert us["environmentalLight"][0] == 1 and us["frameIndex"][0] == 0
    assert np.allclose(us["viewportRight"][0][:3], (9.698, 0, 0), atol=2e-3)


def test_camera_controls_follow_camera_h():
    L = rt.host_lib()
    cam = rt.make_camera(800, 600, (0.0, 5.0, 10.0))
    p0 = np.array(cam.position[:])
    L.rth_camera_keyboard(C.byref(cam), C.c_uint8(0x80), C.c_float(0.5))      # FORWARD: position -= flat(front)*speed*dt
    assert np.allclose(np.array(cam.position[:]) - p0, (0, 0, -5.0), atol=1e-5)
    L.rth_camera_keyboard(C.byref(cam), C.c_uint8(0x10 | 0x08), C.c_float(0.1))  # RIGHT + UP
    assert np.allclose(np.array(cam.position[:]) - p0, (1.0, 1.0, -5.0), atol=1e-5)
    L.rth_camera_keyboard(C.byref(cam), C.c_uint8(0x02), C.c_float(2.0))      # DEFOCUS_UP: +0.1*dt
    assert abs(cam.defocusAngle - 0.2) < 1e-6 and np.linalg.norm(cam.defocusDiskRight[:]) > 0
    L.rth_camera_keyboard(C.byref(cam), C.c_uint8(0x01), C.c_float(100.0))    # clamps at 0
    assert cam.defocusAngle == 0.0
    w0 = np.linalg.norm(cam.viewportRight[:])
    L.rth_camera_scroll(C.byref(cam), C.c_float(10.0))                        # zoom in by e^(10*0.1)
    assert abs(np.linalg.norm(cam.viewportRight[:]) * np.e - w0) < 1e-3
    L.rth_camera_mouse(C.byref(cam), C.c_double(100.0), C.c_double(100.0))    # first event only latches
    yaw0 = cam.yaw
    L.rth_camera_mouse(C.byref(cam), C.c_double(180.0), C.c_double(100.0))
    assert cam.yaw > yaw0 and abs(np.linalg.norm(cam.front[:]) - 1) < 1e-6
    assert abs(L.rth_adjust_rays_per_pixel(C.c_float(199.5), 1, C.c_float(1.0)) - 200.0) < 1e-6   # clamp 1.1..200
    assert abs(L.rth_adjust_rays_per_pixel(C.c_float(2.0), 0, C.c_float(1.0)) - 1.1) < 1e-6


def test_named_scenes_and_png(tmp_path):
    s = rt.scene_textured_sphere(n_quads=224)
    assert s.triangles.size == 2 * 224 * 224 + 16 == 100368
    assert s.materials.size == 7 and len(s.textures) == 1 and s.textures[0].shape == (1024, 1024, 3)
    m = rt.scene_textured_sphere(n_quads=8, container="mirror", tex_size=16)
    assert m.triangles.size == 128 + 14
    assert (m.triangles["materialIndex"][-14:-2] == 6).all() and (m.triangles["materialIndex"][-2:] == 5).all()
    big = rt.scene_big_sphere(n_quads=20)
    assert big.triangles.size == 14 + 800
    img = (np.arange(6 * 5 * 3) % 251).astype(np.uint8).reshape(5, 6, 3)
    path = str(tmp_path / "t.png")
    rt.write_png(path, img)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(path)), img)
    s.save(str(tmp_path / "s.rtsc"))
    s2 = rt.Scene(); s2.load(str(tmp_path / "s.rtsc"))
    assert s2.triangles.tobytes() == s.triangles.tobytes() and s2.materials.tobytes() == s.materials.tobytes()
    assert np.array_equal(s2.textures[0], s.textures[0])


def test_scene_load_rejects_hostile_files(tmp_path):
    """rth_scene_load trusts nothing in the header: negative / huge counts, a texture header out of range and a
    truncated file are refused with an error (no exception crosses the C boundary) and leave the scene untouched."""
    import struct
    s = rt.scene_textured_sphere(n_quads=4, container="cornell", tex_size=8)
    good = str(tmp_path / "good.rtsc")
    s.save(good)
    blob = open(good, "rb").read()
    magic = struct.unpack_from("<q", blob, 0)[0]

    def attempt(data):
        path = str(tmp_path / "bad.rtsc")
        open(path, "wb").write(data)
        t = rt.scene_classic_cornell()
        before = t.triangles.tobytes()
        with pytest.raises(rt.BackendError):
            t.load(path)
        assert t.triangles.tobytes() == before    # the scene is only replaced by a file that was accepted whole

    attempt(blob[: len(blob) // 2])                                               # truncated
    attempt(struct.pack("<8q", magic, -5, 7, 0, 0, 1, 0, 0) + blob[64:])           # negative triangle count
    attempt(struct.pack("<8q", magic, 1 << 60, 7, 0, 0, 1, 0, 0) + blob[64:])      # absurd triangle count
    attempt(struct.pack("<8q", magic, 0, 0, 1 << 40, 0, 0, 0, 0))                  # absurd node section
    attempt(struct.pack("<8q", magic, 0, 0, 0, 0, 99, 0, 0))                       # more textures than slots
    attempt(struct.pack("<8q", magic, 0, 0, 0, 0, 1, 0, 0) + struct.pack("<4i", -4, 8, 3, 0))   # negative texture width
    attempt(struct.pack("<8q", magic, 0, 0, 0, 0, 1, 0, 0) + struct.pack("<4i", 8, 8, 9, 0))    # 9 channels
    attempt(b"not an rtsc file at all, just some bytes that are long enough to hold a header........")
    with pytest.raises(rt.BackendError):
        rt.write_png(str(tmp_path / "no_such_dir" / "x.png"), np.zeros((2, 2, 3), np.uint8))


# ------------------------------------------------------------------------------------------------ N > 1 on gloo
_WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path[:0] = [os.environ["RT_REPO"], os.path.join(os.environ["RT_REPO"], "tests")]
import oracle
rt = importlib.import_module("raytracing2-fork_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scene = rt.scene_classic_cornell()
orc = oracle.OracleScene.from_scene(scene)
cam = rt.make_camera(32, 24, (0.0, 0.0, 15.5))
u = rt.screenshot_uniforms(scene, cam, spp=4, max_bounce=5, env_light=False)
frames = 5
# frame-slice split: my frames only, then the sum the library does with ncclReduce
mine = rt.split_frames(frames, rank, world)
_, part = orc.screenshot(u, frames, rng_mode=rt.RNG_PHILOX, threads=1, frame_list=mine)
t = torch.from_numpy(part.astype(np.int64))
dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
# tile split: all frames, my rows only (zeros elsewhere), then the gather the library does with send/recv
_, full = orc.screenshot(u, frames, rng_mode=rt.RNG_PHILOX, threads=1)
rows = rt.split_rows(24, 4, rank, world)
compact = torch.from_numpy(full[rows].astype(np.int64).copy())
if rank == 0:
    gathered = np.zeros_like(full, dtype=np.int64)
    gathered[rows] = compact.numpy()
    for r in range(1, world):
        rr = rt.split_rows(24, 4, r, world)
        buf = torch.zeros((rr.size, 32, 3), dtype=torch.int64)
        dist.recv(buf, src=r)
        gathered[rr] = buf.numpy()
    ref, sums = orc.screenshot(u, frames, rng_mode=rt.RNG_PHILOX, threads=1)
    assert np.array_equal(t.numpy(), sums.astype(np.int64)), "frame-split reduce differs"
    assert np.array_equal(gathered, sums.astype(np.int64)), "tile-split gather differs"
    assert np.array_equal(oracle.finalize(t.numpy().astype(np.uint32), frames), ref)
    print("GLOO_OK")
else:
    dist.send(compact, dst=0)
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_exchange_logic_on_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, RT_REPO=REPO, OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as sock:          # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_batch_planning_fills_the_budget_without_exceeding_it():
    """rt_plan_batches = the loop nest rt_screenshot runs instead of one dispatch per frame: Philox keeps all samples
    of a frame together when they fit and then packs whole frames; the reference stream (sequential per pixel) runs
    one sample of as many frames as fit.  Never more slots than the budget unless one lane of the image alone
    already exceeds it."""
    P1080, P4k = 1920 * 1080, 3840 * 2160
    budget = 128 << 20
    assert rt.plan_batches(budget, P1080, 64, 4) == (64, 1)             # config 2, one GPU: a whole frame per batch
    assert rt.plan_batches(budget, P1080 // 8, 64, 4) == (64, 4)        # an eighth of the rows (tile split, 8 GPUs)
    assert rt.plan_batches(budget, P4k, 64, 16) == (16, 1)              # 4K: 16 samples of one frame at a time
    assert rt.plan_batches(budget, 512 * 512, 64, 1) == (64, 1)         # config 1
    assert rt.plan_batches(budget, 1000 * 1000, 64, 10) == (64, 2)      # the reference's default screenshot
    assert rt.plan_batches(budget, 1000 * 1000, 64, 10, rt.RNG_REF_PCG) == (1, 10)
    assert rt.plan_batches(budget, P1080, 64, 100, rt.RNG_REF_PCG) == (1, 64)
    rng = np.random.default_rng(3)
    for _ in range(300):
        b = int(rng.integers(1, 1 << 31)); P = int(rng.integers(1, 1 << 24)); spp = int(rng.integers(1, 300))
        frames = int(rng.integers(1, 200)); mode = int(rng.integers(0, 2))
        s, f = rt.plan_batches(b, P, spp, frames, mode)
        assert 1 <= s <= spp and 1 <= f <= frames
        assert mode != rt.RNG_REF_PCG or s == 1
        assert f == 1 or s == spp or mode == rt.RNG_REF_PCG             # frames are only packed whole
        assert s * f * P <= max(min(b, 1 << 30), P)                     # slot ids stay below 2^30
    with pytest.raises(rt.BackendError):
        rt.plan_batches(budget, P1080, 0, 4)
