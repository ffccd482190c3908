#!/usr/bin/env python
"""One screenshot of F frames of a bench workload, nothing else — the process `ncu` wraps to measure the DRAM bytes
k_extend moves per segment (tools/extend_traffic.sh).  Prints one JSON line: workload, frames, segments, launches."""
import argparse, importlib, json, os, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import bench  # noqa: E402  (workload table and scene builders only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--frames", type=int, default=1)
    a = ap.parse_args()
    rt = importlib.import_module("raytracing2-fork_b200")
    scene, cam, u = bench.build_scene(rt, a.workload)
    be = rt.Backend(device=0, rng_mode=rt.RNG_PHILOX)
    be.upload(scene)
    be.reset_counters()
    be.screenshot_device(u, a.frames)
    be.screenshot_fetch()  # synchronises the stream
    c = be.counters()
    print(json.dumps({"workload": a.workload, "frames": a.frames, "segments": int(c["segments"]),
                      "extend_launches": int(c["extend_launches"]), "bvh_width": int(c["bvh_width"]),
                      "bvh_bytes": int(c["bvh_bytes"])}), flush=True)
    be.close()


if __name__ == "__main__":
    main()
