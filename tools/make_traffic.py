#!/usr/bin/env python
"""profiles/traffic.json from an ncu capture, by script (not by hand):

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/make_traffic.py /tmp/raw.csv <kernel regex> <case> <formulation> <what> <scenarios in the captured launch> <source note>

Takes dram__bytes_read.sum + dram__bytes_write.sum of the launches whose name matches, averages them, divides by the
scenarios of the captured launch and stores the per-scenario figure under "<case>:<formulation>:<what>"; bench.py
multiplies it by the scenarios of its own launch (roofline.traffic is per launch, like roofline.achieved)."""
import csv
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(argv):
    raw, pattern, case, form, what, scen = argv[1], argv[2], argv[3], argv[4], int(argv[5]), int(argv[6])
    note = argv[7] if len(argv) > 7 else raw
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    rd, wr, dur, names = [], [], [], []
    for r in rows[2:]:
        if not re.search(pattern, r[ci["Kernel Name"]]):
            continue
        rd.append(float(r[ci["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[ci["dram__bytes_read.sum"]]])
        wr.append(float(r[ci["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[ci["dram__bytes_write.sum"]]])
        dur.append(float(r[ci["gpu__time_duration.sum"]].replace(",", "")))
        names.append(r[ci["Kernel Name"]])
    if not rd:
        raise SystemExit(f"no launch matches {pattern}")
    n = len(rd)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj = {k: v for k, v in tj.items() if "dram_bytes_per_scenario" in v}          # drop hand-entered legacy entries
    tj[f"{case}:{form}:{what}"] = {
        "kernel": names[0], "launches_averaged": n, "scenarios_per_captured_launch": scen,
        "dram_read_per_scenario": sum(rd) / n / scen, "dram_write_per_scenario": sum(wr) / n / scen,
        "dram_bytes_per_scenario": (sum(rd) + sum(wr)) / n / scen,
        "ncu_duration": f"{sum(dur) / n:.3f} {units[ci['gpu__time_duration.sum']]}", "source": note}
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(tj[f"{case}:{form}:{what}"], indent=1))


if __name__ == "__main__":
    main(sys.argv)
