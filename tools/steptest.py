import importlib, sys, time, os
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import torch, numpy as np
rt = importlib.import_module("raytracing2-fork_b200")
import bench
scene, cam, u = bench.build_workload(rt, 1920, 1080)
for timing in (True, False):
    be = rt.Backend(device=0, kernel_timing=timing)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); be.set_stream(stream.cuda_stream)
    be.upload(scene)
    for mode in ("async", "sync_each", "sync_sleep"):
        be.screenshot_device(u, 4); torch.cuda.synchronize(); be.reset_counters()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(14)]
        out = []
        for i in range(6):
            ev[2*i].record(stream)
            be.screenshot_device(u, 4)
            ev[2*i+1].record(stream)
            if mode != "async": torch.cuda.synchronize()
            if mode == "sync_sleep": time.sleep(0.5)
        torch.cuda.synchronize()
        print("timing", timing, mode, [round(ev[2*i].elapsed_time(ev[2*i+1]), 1) for i in range(6)], flush=True)
    be.close()
