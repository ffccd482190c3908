"""ctypes view of oracle/liboracle.so and oracle/_ref/ref_host — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rt = importlib.import_module("raytracing2-fork_b200")
ORACLE_DIR = os.path.join(REPO, "oracle")
REF_HOST = os.path.join(ORACLE_DIR, "_ref", "ref_host")


class OrcCounters(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("paths", C.c_uint64), ("node_visits", C.c_uint64),
                ("tri_tests", C.c_uint64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, os.path.join(ORACLE_DIR, "liboracle.so")])
        L = C.CDLL(path)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_node_count.restype = C.c_int64
        for f in ("orc_random", "orc_philox_draw", "orc_cos01", "orc_sin01", "orc_exp", "orc_acos",
                  "orc_pow_gamma", "orc_ray_bounds"):
            getattr(L, f).restype = C.c_float
        for f in ("orc_cos01", "orc_sin01", "orc_exp", "orc_acos", "orc_pow_gamma"):
            getattr(L, f).argtypes = [C.c_float]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleScene:
    def __init__(self, tris: np.ndarray, mats: np.ndarray, textures=(), build=True):
        self.L = lib()
        tris = np.ascontiguousarray(tris, dtype=rt.TRIANGLE)
        mats = np.ascontiguousarray(mats, dtype=rt.MATERIAL)
        self.n = tris.size
        self.h = C.c_void_p(self.L.orc_scene_create(_p(tris), C.c_int64(tris.size), _p(mats), C.c_int32(mats.size)))
        if not self.h:
            raise RuntimeError("orc_scene_create failed")
        for i, t in enumerate(textures):
            px = np.ascontiguousarray(t, dtype=np.uint8)
            ch = 1 if px.ndim == 2 else px.shape[2]
            assert self.L.orc_scene_set_texture(self.h, i, _p(px), px.shape[1], px.shape[0], ch) == 0
        if build:
            assert self.L.orc_scene_build_bvh(self.h) == 0

    @classmethod
    def from_scene(cls, scene, build=True):
        return cls(scene.triangles, scene.materials, scene.textures, build)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def nodes(self) -> np.ndarray:
        n = self.L.orc_scene_node_count(self.h)
        out = np.zeros(n, dtype=rt.REF_NODE)
        assert self.L.orc_scene_get_nodes(self.h, _p(out)) == 0
        return out

    def permuted(self):
        tris = np.zeros(self.n, dtype=rt.TRIANGLE)
        orig = np.zeros(self.n, dtype=np.int32)
        assert self.L.orc_scene_get_permuted(self.h, _p(tris), _p(orig)) == 0
        return tris, orig

    def trace_rays(self, origins, dirs, use_bvh=True):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        tri = np.zeros(n, np.int32); dst = np.zeros(n, np.float32)
        bu = np.zeros(n, np.float32); bv = np.zeros(n, np.float32)
        rc = self.L.orc_trace_rays(self.h, _p(o), _p(d), C.c_int64(n), int(use_bvh), _p(tri), _p(dst), _p(bu), _p(bv))
        assert rc == 0, rc
        return tri, dst, bu, bv

    def first_hit(self, u, mode=0, rng_mode=0, use_bvh=True, threads=0):
        u = np.ascontiguousarray(u, dtype=rt.UNIFORMS)
        w, h = int(u["width"][0]), int(u["height"][0])
        tri = np.zeros((h, w), np.int32); dst = np.zeros((h, w), np.float32)
        rc = self.L.orc_first_hit(self.h, _p(u), mode, rng_mode, int(use_bvh), threads, _p(tri), _p(dst))
        assert rc == 0, rc
        return tri, dst

    def render_frame(self, u, rng_mode=0, threads=0, region=None, counters=None):
        u = np.ascontiguousarray(u, dtype=rt.UNIFORMS)
        w, h = int(u["width"][0]), int(u["height"][0])
        x0, y0, x1, y1 = region if region else (0, 0, w, h)
        img = np.zeros((h, w, 4), np.float32)
        cn = counters if counters is not None else OrcCounters()
        rc = self.L.orc_render_frame(self.h, _p(u), rng_mode, threads, x0, y0, x1, y1, _p(img), C.byref(cn))
        assert rc == 0, rc
        return img

    def screenshot(self, u, frames, rng_mode=0, threads=0, frame_list=None, counters=None):
        u = np.ascontiguousarray(u, dtype=rt.UNIFORMS)
        w, h = int(u["width"][0]), int(u["height"][0])
        out = np.zeros((h, w, 3), np.uint8)
        sums = np.zeros((h, w, 3), np.uint32)
        fl = None if frame_list is None else np.ascontiguousarray(frame_list, dtype=np.int32)
        cn = counters if counters is not None else OrcCounters()
        rc = self.L.orc_screenshot(self.h, _p(u), frames, rng_mode, threads, _p(fl),
                                   0 if fl is None else fl.size, _p(sums), _p(out), C.byref(cn))
        assert rc == 0, rc
        return out, sums


def finalize(sums: np.ndarray, frames: int) -> np.ndarray:
    sums = np.ascontiguousarray(sums, dtype=np.uint32)
    h, w = sums.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    lib().orc_finalize(_p(sums), w, h, frames, _p(out))
    return out


def camera_uniforms(width, height, pos, hfov, pitch, yaw, focus, defocus, zoom) -> np.ndarray:
    u = np.zeros(1, dtype=rt.UNIFORMS)
    lib().orc_camera_uniforms(width, height, (C.c_float * 3)(*pos), C.c_float(hfov), C.c_float(pitch),
                              C.c_float(yaw), C.c_float(focus), C.c_float(defocus), C.c_float(zoom), _p(u))
    return u


# ---------------------------------------------------------------------------------------------- ref_host
def have_ref_host() -> bool:
    return os.path.exists(REF_HOST)


def read_rtsc(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype="<i8", count=8)
    assert hdr[0] == 0x43535452
    n_t, n_m, n_n, n_p, n_tex = (int(x) for x in hdr[1:6])
    off = 64
    tris = np.frombuffer(raw, dtype=rt.TRIANGLE, count=n_t, offset=off); off += 80 * n_t
    mats = np.frombuffer(raw, dtype=rt.MATERIAL, count=n_m, offset=off); off += 96 * n_m
    nodes = np.frombuffer(raw, dtype=rt.REF_NODE, count=n_n, offset=off); off += 48 * n_n
    perm = np.frombuffer(raw, dtype=rt.TRIANGLE, count=n_p, offset=off); off += 80 * n_p
    tex = []
    for _ in range(n_tex):
        w, h, ch, _z = np.frombuffer(raw, dtype="<i4", count=4, offset=off); off += 16
        tex.append(np.frombuffer(raw, dtype=np.uint8, count=w * h * ch, offset=off).reshape(h, w, ch)); off += w * h * ch
    return dict(tris=tris, mats=mats, nodes=nodes, perm=perm, tex=tex)


def ref_scene(kind, in_path, out_path):
    subprocess.check_call([REF_HOST, "scene", kind, in_path or "-", out_path])
    return read_rtsc(out_path)


def ref_camera(width, height, pos, hfov, pitch, yaw, focus, defocus, zoom, out_path) -> np.ndarray:
    args = [REF_HOST, "camera", str(width), str(height)] + [repr(float(x)) for x in (*pos, hfov, pitch, yaw, focus, defocus, zoom)] + [out_path]
    subprocess.check_call(args)
    return np.fromfile(out_path, dtype=rt.UNIFORMS)


def ref_rng(seed, n, out_path):
    subprocess.check_call([REF_HOST, "rng", str(seed), str(n), out_path])
    raw = np.fromfile(out_path, dtype=np.uint32).reshape(n, 2)
    return raw[:, 0].copy(), raw[:, 1].copy().view(np.float32)
