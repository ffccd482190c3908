#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on B200 (BASELINE.json metric).

A step = one offline screenshot of BASELINE config 2 (Cornell box + 100 352-triangle textured mesh,
1920x1080, 256 spp = 4 frames x 64 spp, max depth 20, Philox RNG): raygen -> [extend -> shade] x depth ->
accumulate -> resolve per frame, then finalize.  `value` times that with the scene resident in HBM
(rt_screenshot_device); `e2e` times the whole job from HOST buffers through the C-ABI (triangle /
material / texture upload, LBVH build, rt_screenshot with the RGB8 image copied back).

N > 1 (torchrun, one rank per GPU): frame-slice split — the screenshot has 4*N frames, rank r renders
frames f with f % N == r, the 8-bit frame sums are reduced to rank 0 over NCCL inside the library and
rank 0 finalizes; per-GPU work is fixed (weak scaling), all inside the timed region.

--impl reference: the reference's path on the host CPU.  The reference implements it only as GLSL
(no GL in this image), so this arm times the CPU oracle port on all host threads on a bounded sample
of the same workload (`cpu_baseline.kind` = "port").
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOAD = dict(width=1920, height=1080, spp_per_frame=64, frames=4, max_bounce=20, n_quads=224, tex_size=1024)
CPU_SAMPLE = dict(crop_w=480, crop_h=270, frames=1, spp=4)  # bounded CPU sample of the same workload


# The other BASELINE configs (parity-test cases; `--workload` measures them for the record, the default
# bench line is always config 2).
OTHER_WORKLOADS = {
    "config1": dict(width=512, height=512, frames=1, max_bounce=8),
    "config3": dict(width=1920, height=1080, frames=8, max_bounce=16),
    "config4": dict(width=3840, height=2160, frames=16, max_bounce=8),
    "config2_robot": dict(width=1920, height=1080, frames=4, max_bounce=20),   # needs tools/make_assets.sh
}


def build_workload(rt, width, height):
    name = os.environ.get("RT_BENCH_WORKLOAD", "config2")
    if name == "config1":
        scene = rt.scene_classic_cornell()
        cam = rt.make_camera(width, height, (0.0, 0.0, 15.5))
    elif name == "config3":
        scene = rt.scene_textured_sphere(n_quads=WORKLOAD["n_quads"], container="mirror", tex_size=WORKLOAD["tex_size"])
        cam = rt.camera_for_box(scene, width, height)
    elif name == "config2_robot":
        scene = rt.scene_from_rtsc(os.path.join(REPO, "assets", "_gen", "robot.rtsc"), container="cornell")
        cam = rt.camera_for_box(scene, width, height)
    elif name == "config4":
        scene = rt.scene_big_sphere(n_quads=2236)
        cam = rt.make_camera(width, height, (0.0, 0.0, 15.5))
    else:
        scene = rt.scene_textured_sphere(n_quads=WORKLOAD["n_quads"], container="cornell", tex_size=WORKLOAD["tex_size"])
        cam = rt.camera_for_box(scene, width, height)
    u = rt.screenshot_uniforms(scene, cam, spp=WORKLOAD["spp_per_frame"], max_bounce=WORKLOAD["max_bounce"],
                               env_light=False)
    return scene, cam, u


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("RT_BENCH_CLOCK_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def pinned_copy(torch, a: np.ndarray) -> np.ndarray:
    """The same bytes in page-locked host memory (numpy view of a pinned torch tensor)."""
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy()
    v[:] = np.frombuffer(a.tobytes(), dtype=np.uint8)
    return v.view(a.dtype).reshape(a.shape), t  # keep `t` alive as long as the view is used


# ------------------------------------------------------------------------------------------------ CPU arm
_CPU_ARM = {}


def cpu_arm(rt):
    """The CPU implementation that gets timed, prepared once: the reference's OWN compute shader source
    (oracle/_ref/libref_shader_libm.so: compute.glsl rewritten syntactically by oracle/glsl2cpp.py, compiled with
    g++ -O2 against the reference's glm, libm elementary functions; built where /root/reference exists and shipped
    prebuilt) over the reference's BVH — `kind: "reference"`.  If that library is missing the oracle port is
    timed instead — `kind: "port"`.  Either way the segment count comes from the oracle on the identical frame
    (RT_RNG_REF_PCG: the oracle's frame is bit-identical to the shader's, tests/test_refshader_cpu.py)."""
    if _CPU_ARM:
        return _CPU_ARM
    import oracle  # the CPU baseline leg is one of the two places allowed to execute oracle/
    scene, cam, u = build_workload(rt, WORKLOAD["width"], WORKLOAD["height"])
    orc = oracle.OracleScene.from_scene(scene)
    uu = u.copy()
    uu["numRaysPerPixel"] = CPU_SAMPLE["spp"]
    W, H = WORKLOAD["width"], WORKLOAD["height"]
    x0, y0 = (W - CPU_SAMPLE["crop_w"]) // 2, (H - CPU_SAMPLE["crop_h"]) // 2
    region = (x0, y0, x0 + CPU_SAMPLE["crop_w"], y0 + CPU_SAMPLE["crop_h"])
    _CPU_ARM.update(orc=orc, u=uu, region=region, shader=None, kind="port", segments=None)
    try:
        import refshader
        if refshader.available(False):
            tris, _ = orc.permuted()   # the oracle's BVH is pinned to the reference builder's (tests/test_oracle_cpu.py)
            _CPU_ARM["shader"] = refshader.Loaded(scene, spec_math=False, bvh=(orc.nodes(), tris))
            _CPU_ARM["kind"] = "reference"
            cn = oracle.OrcCounters()
            for f in range(CPU_SAMPLE["frames"]):
                uu["frameIndex"] = f
                orc.render_frame(uu, rng_mode=rt.RNG_REF_PCG, region=region, counters=cn)
            _CPU_ARM["segments"] = int(cn.segments)
    except Exception as e:  # a missing or stale harness must not take the bench down
        print(f"bench.py: reference shader harness unavailable ({e}); timing the oracle port", file=sys.stderr)
        _CPU_ARM.update(shader=None, kind="port", segments=None)
    return _CPU_ARM


def cpu_sample(rt, threads=0):
    """One bounded CPU sample: a centred crop of the config-2 image at reduced spp (throughput in Mrays/s does not
    depend on spp or crop size).  Returns (Mrays/s, seconds, segments, threads)."""
    import oracle
    arm = cpu_arm(rt)
    uu, region = arm["u"], arm["region"]
    nthreads = threads if threads > 0 else (os.cpu_count() or 1)
    if arm["shader"] is not None:
        t0 = time.perf_counter()
        for f in range(CPU_SAMPLE["frames"]):
            uu["frameIndex"] = f
            arm["shader"].render(uu, region=region, threads=threads)
        dt = time.perf_counter() - t0
        return arm["segments"] / dt * 1e-6, dt, arm["segments"], nthreads
    cn = oracle.OrcCounters()
    t0 = time.perf_counter()
    for f in range(CPU_SAMPLE["frames"]):
        uu["frameIndex"] = f
        arm["orc"].render_frame(uu, rng_mode=rt.RNG_PHILOX, threads=threads, region=region, counters=cn)
    dt = time.perf_counter() - t0
    return cn.segments / dt * 1e-6, dt, int(cn.segments), nthreads


def sample_text(rt=None):
    base = (f"centred {CPU_SAMPLE['crop_w']}x{CPU_SAMPLE['crop_h']} crop of the 1920x1080 config-2 image, "
            f"{CPU_SAMPLE['frames']} frame x {CPU_SAMPLE['spp']} spp, depth {WORKLOAD['max_bounce']}, ")
    if _CPU_ARM.get("kind") == "reference":
        return base + ("the reference's own compute.glsl compiled for the CPU (glsl2cpp.py + glm, g++ -O2, libm) over the "
                       "reference's BVH (BVH.h), the shader's PCG stream, std::thread over rows; segments counted by the oracle "
                       "on the identical frame")
    return base + "Philox, oracle port of compute.glsl with the reference's own BVH (BVH.h), std::thread over rows"


def run_reference(args, rt):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt, segs, nthreads = cpu_sample(rt)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": nthreads, "kind": _CPU_ARM.get("kind", "port"),
                         "sample": sample_text()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    if os.environ.get("RT_BENCH_WORKLOAD", "config2") != "config2":
        return {"workload": "BASELINE " + os.environ["RT_BENCH_WORKLOAD"] + f" ({WORKLOAD['width']}x{WORKLOAD['height']}, "
                f"{WORKLOAD['frames']} frames x 64 spp per GPU, depth {WORKLOAD['max_bounce']})", "frames_total": WORKLOAD["frames"] * n_gpus}
    return {
        "workload": "BASELINE config 2: Cornell box (addCornellBox 0.17/0.3, light 15.0) + synthetic textured "
                    "displaced sphere 100 352 triangles + 16 container triangles, 1920x1080, 256 spp = 4 frames x 64 spp "
                    "per GPU, max depth 20, environmentalLight 0",
        "rng": "philox4x32-10 keyed (pixel, frame, sample, bounce, draw)",
        "frames_total": WORKLOAD["frames"] * n_gpus,
        "split": "none" if n_gpus == 1 else "frame-slice (f % N == rank) + ncclReduce of the 8-bit frame sums to rank 0",
        "cache": "path state of one batch (64 samples x 2.07 M pixels = 133 M paths x 128 B = 17 GB) is far larger than "
                 "L2; the 10 MB BVH is L2-resident by nature of the workload; no explicit flush",
    }


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rt):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA backend is the product, there is no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = WORKLOAD["width"], WORKLOAD["height"]
    scene, cam, u = build_workload(rt, W, H)
    frames = (args.frames if args.frames > 0 else WORKLOAD["frames"]) * world
    tris, mats, texs = scene.triangles, scene.materials, scene.textures

    split = rt.SPLIT_NONE if world == 1 else (rt.SPLIT_TILES if args.split == "tiles" else rt.SPLIT_FRAMES)
    be = rt.Backend(device=local_rank, rng_mode=rt.RNG_PHILOX, split_mode=split, rank=rank, world_size=world,
                    kernel_timing=True)
    # The library launches on the stream it is given; torch's default stream is the legacy stream 0,
    # which rt_set_stream treats as "use the ctx's own stream", and events recorded on stream 0 would
    # not wait for that non-blocking stream.  So: one explicit side stream for library and events.
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    be.set_stream(stream.cuda_stream)
    if world > 1:
        obj = [be.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        be.comm_init(obj[0])
    be.upload(scene)

    # parity gate without the oracle: first-hit ids of the workload scene against the committed golden
    gate = None
    if rank == 0 and args.workload != "config2":
        gate = "n/a (golden is for config 2)"
    elif rank == 0:
        try:
            meta = json.load(open(os.path.join(REPO, "tests", "golden", "golden.json")))
            cam_s = rt.camera_for_box(scene, 240, 135)
            us = rt.screenshot_uniforms(scene, cam_s, spp=4, max_bounce=6, env_light=False)
            tri, dst = be.first_hit(us, rt.FIRST_HIT_CENTRE)
            ok = (zlib.crc32(tri.tobytes()) & 0xffffffff) == meta["config2_first_hit_240x135_crc"] and \
                 (zlib.crc32(dst.tobytes()) & 0xffffffff) == meta["config2_first_hit_240x135_dst_crc"]
            gate = "first-hit ids+dst 240x135 == tests/golden crc: " + ("PASS" if ok else "FAIL")
            if not ok:
                raise SystemExit("bench.py: parity gate failed — " + gate)
        except FileNotFoundError:
            gate = "golden missing"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        be.screenshot_device(u, frames)
    barrier()
    be.reset_counters()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RT_BENCH_NO_CLOCKS"):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev0.record(stream)
    step_ev[0].record(stream)
    for i in range(args.steps):
        be.screenshot_device(u, frames)
        step_ev[i + 1].record(stream)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    step_ms = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    c = be.counters()
    t = torch.tensor([ms_total, float(c["segments"]), float(c["kernel_launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, segments, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        segments, launches = float(c["segments"]), float(c["kernel_launches"])
    ms_per_step = ms_total / args.steps
    value = segments / (ms_total * 1e-3) * 1e-6
    extend_ms, extend_launches = c["extend_ms"], c["extend_launches"]

    # ---- end to end from host buffers through the C-ABI (upload + build + screenshot + image readback)
    ptris, _k1 = pinned_copy(torch, tris)
    pmats, _k2 = pinned_copy(torch, mats)
    ptexs = [pinned_copy(torch, t_) for t_ in texs]
    h2d = ptris.nbytes + pmats.nbytes + sum(p[0].nbytes for p in ptexs) + 192 * frames
    d2h = W * H * 3 if rank == 0 else 0

    phases = {"upload": 0.0, "build": 0.0, "screenshot": 0.0}

    def e2e_step():
        t_a = time.perf_counter()
        be.set_triangles(ptris)
        be.set_materials(pmats)
        for i, (p, _) in enumerate(ptexs):
            be.set_texture(i, p)
        t_b = time.perf_counter()
        be.build()
        t_c = time.perf_counter()
        out = be.screenshot(u, frames, want_output=(rank == 0))
        t_d = time.perf_counter()
        phases["upload"] += t_b - t_a; phases["build"] += t_c - t_b; phases["screenshot"] += t_d - t_c
        return out

    e2e_step()
    barrier()
    be.reset_counters()
    for k in phases:
        phases[k] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        shot = e2e_step()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    ce = be.counters()
    te = torch.tensor([wall, float(ce["segments"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = te.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = te.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        wall, seg_e2e = float(tm[0]), float(ts[1])
    else:
        seg_e2e = float(ce["segments"])
    e2e_value = seg_e2e / wall * 1e-6

    # ---- roofline of the dominant kernel (k_extend): algorithmic bytes from an instrumented replay of
    # one step (identical rays: the RNG is counter based), duration from the timed loop's own events
    roof = None
    if rank == 0:
        bi = rt.Backend(device=local_rank, rng_mode=rt.RNG_PHILOX, instrument=True)
        bi.set_stream(stream.cuda_stream)
        bi.upload(scene)
        bi.screenshot_device(u, 1)  # one frame is 1/frames of a step; per-segment averages are what we need
        torch.cuda.synchronize()
        ci = bi.counters()
        bi.close()
        seg_i = max(ci["segments"], 1)
        n_inner, n_tri = ci["node_visits"] / seg_i, ci["tri_tests"] / seg_i
        bytes_per_seg = n_inner * 64 + n_tri * 48 + 96
        seg_per_launch = c["segments"] / max(extend_launches, 1)
        peak, peak_src = measured_hbm_peak()
        avg_launch_s = extend_ms * 1e-3 / max(extend_launches, 1)
        achieved = bytes_per_seg * seg_per_launch / avg_launch_s * 1e-9 if avg_launch_s > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(REPO, "profiles", "extend_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        # SURVEY 8d's second term: flops(r) = n_inner*48 + n_tri*56 against the non-FMA FP32 issue peak
        # (SMs x 128 lanes x SM clock; parity forbids FMA contraction, so one flop per lane per cycle)
        props = torch.cuda.get_device_properties(local_rank)
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        fp32_peak = props.multi_processor_count * 128 * sm_mhz * 1e6 * 1e-12
        flops_per_seg = n_inner * 48 + n_tri * 56
        fp32_achieved = flops_per_seg * seg_per_launch / avg_launch_s * 1e-12 if avg_launch_s > 0 else None
        roof = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "node_visits_per_segment": n_inner, "tri_tests_per_segment": n_tri,
                "bytes_per_segment": bytes_per_seg, "segments_per_launch": seg_per_launch,
                "fp32": {"achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s (non-FMA)",
                         "frac": (fp32_achieved / fp32_peak) if fp32_achieved else None,
                         "flops_per_segment": flops_per_seg},
                "avg_launch_ms": avg_launch_s * 1e3, "extend_share_of_step": extend_ms / ms_total if ms_total else None,
                "note": "algorithmic bytes = n_inner*64 + n_tri*48 + 96 per segment (SURVEY 8d; n_inner = visits of the "
                        "64-byte 4-wide nodes k_extend walks); the 8 MB BVH of this workload is L2-resident, so this "
                        "traffic is served on chip — the bound that applies is L2 latency / instruction issue, see "
                        "DESIGN.md §6"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, segs, nthreads = cpu_sample(rt)
        port = None
        if _CPU_ARM.get("kind") == "reference":   # for the record: the oracle port on the same crop (Philox), same threads
            import oracle
            cn = oracle.OrcCounters()
            t0 = time.perf_counter()
            _CPU_ARM["orc"].render_frame(_CPU_ARM["u"], rng_mode=rt.RNG_PHILOX, region=_CPU_ARM["region"], counters=cn)
            port = {"value": cn.segments / (time.perf_counter() - t0) * 1e-6, "unit": "Mrays/s", "kind": "port"}
        cpu = {"value": v, "unit": "Mrays/s", "cores": nthreads, "kind": _CPU_ARM.get("kind", "port"), "sample": sample_text(),
               "oracle_port": port,
               "seconds": dt, "segments": segs}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(workload_config(world), frames_total=frames,
                                                  split=("none" if world == 1 else args.split)),
            "seconds_per_screenshot": ms_per_step * 1e-3, "step_ms": [round(x, 2) for x in step_ms],
            "segments_per_step": segments / args.steps,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "seconds_per_step": wall / n_e2e, "steps": n_e2e,
                    "host_phases_ms_per_step": {k: round(v / n_e2e * 1e3, 2) for k, v in phases.items()},
                    "device_ms_per_step": e0.elapsed_time(e1) / n_e2e,
                    "includes": "rt_scene_set_triangles/materials/texture from pinned host memory, rt_scene_build, "
                                "rt_screenshot with RGB8 readback"},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "parity_gate": gate,
            "bvh": {"build_ms": c["build_ms"], "nodes": c["bvh_nodes"], "bytes": c["bvh_bytes"], "depth": c["bvh_depth"]},
            "checksum": int(zlib.crc32(shot.tobytes()) & 0xffffffff) if shot is not None else None,
        }
        print(json.dumps(line), flush=True)
    be.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (default 4 = 256 spp)")
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config2_robot"],
                    help="BASELINE config to measure (default and contract: config2)")
    ap.add_argument("--split", default="frames", choices=["frames", "tiles"],
                    help="N > 1: frame-slice split + ncclReduce (default) or image-tile split + gather")
    args = ap.parse_args()
    if args.workload != "config2":
        os.environ["RT_BENCH_WORKLOAD"] = args.workload
        WORKLOAD.update(OTHER_WORKLOADS[args.workload])
    rt = importlib.import_module("raytracing2-fork_b200")
    if args.impl == "reference":
        run_reference(args, rt)
    else:
        run_gpu(args, rt)


if __name__ == "__main__":
    main()
