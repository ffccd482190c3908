#!/usr/bin/env python
"""Merge an ncu pass made by tools/extend_traffic.sh into profiles/extend_traffic.json:
per workload, dram__bytes_read + dram__bytes_write of every k_extend launch of one screenshot, summed, divided by the
segments that screenshot traced (the library's own counter).  bench.py prints it as roofline.traffic (x segments per
launch).  usage: tools/extend_traffic.py gpurun_out/traffic_config2 [more prefixes]"""
import csv, json, os, subprocess, sys, time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles", "extend_traffic.json")


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
            "sector": 1.0}.get(u, 1.0)


def main():
    try:
        table = json.load(open(OUT))
        if "dram_bytes_per_launch" in table:  # round-1 format
            table = {}
    except Exception:
        table = {}
    head = subprocess.run(["git", "-C", REPO, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    for prefix in sys.argv[1:]:
        run = json.loads(open(prefix + ".json").read().strip().splitlines()[-1])
        rows = list(csv.reader(open(prefix + ".csv")))
        hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        col = {h: i for i, h in enumerate(rows[hi])}
        acc = {"k_extend": {"dram": 0.0, "ms": 0.0, "l2_read_sectors": 0.0, "launches": set()},
               "k_shade": {"dram": 0.0, "ms": 0.0, "l2_read_sectors": 0.0, "launches": set()}}
        for r in rows[hi + 1:]:
            if len(r) < len(col):
                continue
            name = r[col["Kernel Name"]]
            if "k_extend<1" in name or "k_extend<(bool)1" in name:
                continue  # the instrumented variant is not the timed kernel
            k = "k_extend" if "k_extend" in name else ("k_shade" if "k_shade" in name else None)
            if k is None:
                continue
            v = float(r[col["Metric Value"]].replace(",", "")) * unit_scale(r[col["Metric Unit"]])
            m = r[col["Metric Name"]]
            a = acc[k]
            a["launches"].add(r[col["ID"]])
            if m.startswith("dram__bytes"):
                a["dram"] += v
            elif m.startswith("gpu__time"):
                a["ms"] += v
            elif m.startswith("lts__t_sectors"):
                a["l2_read_sectors"] += v
        seg = run["segments"]
        e = acc["k_extend"]
        table[run["workload"]] = {
            "dram_bytes_per_segment": e["dram"] / seg,
            "k_extend_dram_gbs_under_ncu": e["dram"] / (e["ms"] * 1e-3) * 1e-9 if e["ms"] else None,
            "k_shade_dram_bytes_per_segment": acc["k_shade"]["dram"] / seg,
            "k_shade_dram_gbs_under_ncu": acc["k_shade"]["dram"] / (acc["k_shade"]["ms"] * 1e-3) * 1e-9 if acc["k_shade"]["ms"] else None,
            "segments": seg, "frames": run["frames"], "extend_launches": len(e["launches"]),
            "bvh_width": run["bvh_width"], "bvh_bytes": run["bvh_bytes"],
            "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every k_extend launch of one "
                      f"{run['frames']}-frame screenshot (tools/extend_traffic.sh {run['workload']} {run['frames']}); "
                      f"raw: profiles/r2_traffic_{run['workload']}.csv",
            "git": head, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        }
        print(run["workload"], json.dumps(table[run["workload"]], indent=1))
    json.dump(table, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
