#!/usr/bin/env python
"""Turn the raw ncu outputs of tools/final_evidence.sh into the summaries committed under profiles/.

  tools/profile_report.py launches gpurun_out/final_launches.csv            -> per-kernel table (markdown, stdout)
  tools/profile_report.py full gpurun_out/final_c2_full.ncu-rep [k_regex]   -> counters (tools/ncu_summary.py) + per
        launch the instructions that collect the most warp-stall samples and the share of instructions / samples of
        every source line above 1.5 % (ncu --page source, SASS correlated with -lineinfo)
"""
import collections, csv, io, os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    d = collections.OrderedDict()
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "").replace("rt::", "")
        e = d.setdefault(k, [0.0, 0])
        e[0] += float(r[-1])
        e[1] += 1
    tot = sum(v[0] for v in d.values())
    print(f"{len(rows)} launches, {tot / 1e6:.2f} ms in total (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][0]):
        print(f"| {k} | {v[1]} | {v[0] / 1e6:.2f} | {100 * v[0] / tot:.1f}% | {v[0] / v[1] / 1e3:.1f} |")
    for name in ("k_extend", "k_shade"):
        seq = [float(r[-1]) / 1e6 for r in rows if name in r[4]]
        print(f"\n{name}, first {min(20, len(seq))} launches (ms): " + " ".join(f"{x:.2f}" for x in seq[:20]))


def _clean(name):
    name = (name or "").replace("(bool)", "").replace("(int)", "").replace("void ", "").replace("rt::", "")
    return name.split("(")[0]


def _page(rep, view, regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--kernel-name",
                          "regex:" + regex], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, regex="k_extend|k_shade"):
    sys.stdout.write(subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
    # plain SASS view: one table per launch (ncu prints each twice), one row per instruction
    sass, name = [], None
    for r in _page(rep, "sass", regex):
        if r and r[0] == "Kernel Name":
            name = r[1]
        elif r and r[0] == "Address":
            sass.append({"name": _clean(name), "rows": []})
        elif r and len(r) > 8 and r[0].startswith("0x") and r[2].isdigit():
            sass[-1]["rows"].append((r[1].strip(), int(r[2]), int(r[5])))
    # CUDA source view correlated through -lineinfo: rows whose first column is a line number carry the line's totals
    lines, cur, cur_file = [], None, None
    for r in _page(rep, "cuda,sass", regex):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1])
        elif r[0] == "Function Name":
            if cur is None or cur["name"] != r[1] or cur_file in cur["files"]:
                cur = {"name": r[1], "files": set(), "lines": collections.OrderedDict()}
                lines.append(cur)
            cur["files"].add(cur_file)
        elif cur is not None and len(r) > 7 and r[0].isdigit():
            try:
                e = cur["lines"].setdefault((cur_file, int(r[0]), r[1].strip()[:100]), [0, 0])
                e[0] += int(r[4])
                e[1] += int(r[7])
            except ValueError:
                pass
    seen, order = set(), []
    for L in sass:
        ts, ti = sum(s for _, s, _ in L["rows"]) or 1, sum(i for _, _, i in L["rows"]) or 1
        if (L["name"], ts, ti) not in seen:
            seen.add((L["name"], ts, ti))
            order.append((L, ts, ti))
    seen = set()
    uniq_lines = []
    for L in lines:
        ts, ti = sum(v[0] for v in L["lines"].values()), sum(v[1] for v in L["lines"].values())
        if (L["name"], ts, ti) not in seen:
            seen.add((L["name"], ts, ti))
            uniq_lines.append((L, ts or 1, ti or 1))
    for li, (L, ts, ti) in enumerate(order):
        print(f"\n## launch {li}: {L['name']} - {ti:,} warp instructions, {ts:,} stall samples\n")
        print("instructions collecting the most stall samples (a sample lands on the instruction that WAITS):\n")
        print("| SASS | samples | instructions |\n|---|---|---|")
        for txt, s, i in sorted(L["rows"], key=lambda x: -x[1])[:12]:
            print(f"| `{txt[:80]}` | {100 * s / ts:.1f}% | {100 * i / ti:.2f}% |")
        if li < len(uniq_lines):
            LL, ls, lins = uniq_lines[li]
            print("\nsource lines above 1.5 % of the instructions or of the samples:\n")
            print("| file:line | instructions | samples | source |\n|---|---|---|---|")
            for (f, ln, src), (s, i) in sorted(LL["lines"].items(), key=lambda kv: -kv[1][1]):
                if 100 * i / lins >= 1.5 or 100 * s / ls >= 1.5:
                    print(f"| {f}:{ln} | {100 * i / lins:.1f}% | {100 * s / ls:.1f}% | `{src.replace('|', '/')}` |")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 3 and sys.argv[1] == "full":
        full(sys.argv[2], *(sys.argv[3:4]))
    else:
        sys.exit(__doc__)
