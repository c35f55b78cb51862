#!/usr/bin/env python3
"""Turn the raw captures of one profiling run into the tracked artifacts under profiles/.

    python tools/make_profiles.py r01 gpurun_out/r01_step.ncu-rep gpurun_out/r01_launches.csv

* <tag>_ncu_full_summary.txt   key metrics of every launch in the `ncu --set full` report
* <tag>_traffic.json           DRAM bytes per launch of every kernel (read by bench.py for roofline.traffic)
* <tag>_hotspots_k_morph_pair_erode.txt / <tag>_sass_k_morph_pair_erode.txt   stall hot spots and SASS of the top kernel
* <tag>_launches.csv / <tag>_launches_summary.txt   the `gpu__time_duration.sum` launch list and its per-kernel shares
"""
import contextlib
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import launch_summary  # noqa: E402
import ncu_hotspots  # noqa: E402
import ncu_summary  # noqa: E402


def capture(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def main(tag, rep, launches=None, streams=64):
    prof = os.path.join(ROOT, "profiles")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    tmp = os.path.join(prof, "." + tag + "_raw.csv")
    with open(tmp, "w") as f:
        f.write(raw)
    with open(os.path.join(prof, tag + "_ncu_full_summary.txt"), "w") as f:
        f.write(capture(ncu_summary.main, tmp))
    rows = list(csv.reader(io.StringIO(raw)))
    os.remove(tmp)
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]
    traffic = {}
    for r in rows[2:]:
        name = re.sub(r"\(bool\)", "", re.sub(r"^void ", "", r[idx["Kernel Name"]]))
        name = re.sub(r"\((?!.*<).*", "", name) if "<" not in name else name[:name.index(">") + 1]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = float(r[idx[m]].replace(",", ""))
            u = units[idx[m]]
            tot += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic.setdefault(name, tot)           # first launch of each kernel (attempt-1 launches come first)
    with open(os.path.join(prof, tag + "_traffic.json"), "w") as f:
        json.dump({"streams": streams,
                   "source": "ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum per launch), "
                             "profiles/%s_ncu_full_summary.txt" % tag,
                   "dram_bytes_per_launch": traffic}, f, indent=1)
    names = [r[idx["Kernel Name"]] for r in rows[2:]]
    # the source page lists every launch; pick the first erosion launch
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    order = [row[1] for row in csv.reader(src.splitlines()) if row and row[0] == "Kernel Name"]
    ki = next(i for i, n in enumerate(order) if "k_morph_pair" in n and "(bool)0, (bool)0" in n)
    with open(os.path.join(prof, tag + "_hotspots_k_morph_pair_erode.txt"), "w") as f:
        f.write(capture(ncu_hotspots.main, rep, 40, ki))
    so = os.path.join(ROOT, "lane_tracker_b200", "liblane_tracker_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z12k_morph_pairILb0ELb0EEv8MorphJobS0_6LtDimsiimmPKiS3_", so],
                          capture_output=True, text=True).stdout
    lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    with open(os.path.join(prof, tag + "_sass_k_morph_pair_erode.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if launches:
        dst = os.path.join(prof, tag + "_launches.csv")
        if os.path.abspath(launches) != os.path.abspath(dst):
            shutil.copyfile(launches, dst)
        with open(os.path.join(prof, tag + "_launches_summary.txt"), "w") as f:
            f.write(capture(launch_summary.main, dst))
    print("kernels in report:", len(names), "; erosion launch index", ki, "; SASS lines", len(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
