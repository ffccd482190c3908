#!/usr/bin/env python
"""Where does the tile split lose time?  One GPU plays rank r of an N-way image-tile split of a bench workload (no
communicator: rt_screenshot_partial) and reports the device time of its share, split by kernel class, next to the
unsplit job.  usage: tools/tile_split_probe.py --workload config4 --frames 16 --world 8 --ranks 0,3,7"""
import argparse, importlib, json, os, sys, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0,3,7")
    ap.add_argument("--band-rows", type=int, default=0)
    a = ap.parse_args()
    import torch
    rt = importlib.import_module("raytracing2-fork_b200")
    scene, cam, u = bench.build_scene(rt, a.workload)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def run(split, rank, world):
        be = rt.Backend(device=0, rng_mode=rt.RNG_PHILOX, split_mode=split, rank=rank, world_size=world,
                        band_rows=a.band_rows, kernel_timing=True)
        be.set_stream(stream.cuda_stream)
        be.upload(scene)
        be.screenshot_partial(u, a.frames)          # warm-up
        be.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        be.L.rt_screenshot_partial(be.h, u.ctypes.data_as(__import__("ctypes").c_void_p), a.frames)
        e1.record(stream)
        torch.cuda.synchronize()
        c = be.counters()
        be.close()
        return {"rank": rank, "world": world, "ms": e0.elapsed_time(e1), "extend_ms": c["extend_ms"], "other_kernels_ms": c["shade_ms"],
                "segments": int(c["segments"]), "extend_launches": int(c["extend_launches"]), "launches": int(c["kernel_launches"])}

    out = [run(rt.SPLIT_NONE, 0, 1)]
    for r in [int(x) for x in a.ranks.split(",")]:
        out.append(run(rt.SPLIT_TILES, r, a.world))
    for o in out:
        o["ms_x_world"] = o["ms"] * o["world"]
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
