#!/usr/bin/env python3
"""Turn the raw captures of one profiling run (tools/profile_round.sh) into the tracked artifacts under profiles/.

    python tools/make_profiles.py r02 gpurun_out/r02a_step.ncu-rep gpurun_out/r02a_launches.csv [gpurun_out/r02a_bench.json]

* <tag>_ncu_full_summary.txt   key metrics of every launch in the `ncu --set full` report (one warm step, S = 64)
* <tag>_traffic.json           DRAM bytes per launch of every kernel and per morphology stage (read by bench.py for
                               roofline.traffic)
* <tag>_hotspots/<kernel>.txt  stall hot spots (ncu source page) of every kernel in the step
* <tag>_sass/<kernel>.txt      SASS listing (cuobjdump of the in-tree library) + opcode histogram of every such kernel
* <tag>_launches.csv / <tag>_launches_summary.txt   the `gpu__time_duration.sum` launch list and its per-kernel shares
* <tag>_bench.json             the bench line of the same run
"""
import collections
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

LIB = os.path.join(ROOT, "lane_tracker_b200", "liblane_tracker_b200.so")
STAGE_OF = {"k_morph<55, 0, 0": "erode55", "k_morph<29, 0, 0": "erode55", "k_morph<55, 1, 1": "tophat55",
            "k_morph<29, 1, 1": "tophat55"}


def capture(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def short_name(full):
    """'void <unnamed>::k_morph<(int)55, (bool)0, (bool)0, (int)2>(...)' -> 'k_morph<55, 0, 0, 2>'"""
    n = re.sub(r"^void ", "", full)
    n = n.replace("<unnamed>::", "")
    n = re.sub(r"\((?:int|bool)\)", "", n)
    depth, out = 0, []
    for ch in n:                         # cut at the argument list: the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out).strip()


def file_name(short):
    return re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")


def library_sass():
    """{short kernel name: [sass lines]} of the in-tree library."""
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = funcs.setdefault(short_name(dem), [])
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            cur.append(re.sub(r"/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
    return funcs


def opcode_histogram(lines):
    c = collections.Counter()
    for l in lines:
        t = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l).split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        c[op.rstrip(";").split(".")[0]] += 1
    return c


def main(tag, rep, launches=None, bench=None, streams=64):
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
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic, per_stage, dur = collections.OrderedDict(), collections.OrderedDict(), collections.OrderedDict()
    for r in rows[2:]:
        name = short_name(r[idx["Kernel Name"]])
        if name in traffic:               # the capture may run into the next step: first launch of each kernel
            continue
        tot = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[name] = tot
        dur[name] = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        for key, stage in STAGE_OF.items():
            if name.startswith(key):
                per_stage[stage] = per_stage.get(stage, 0.0) + tot
    with open(os.path.join(prof, tag + "_traffic.json"), "w") as f:
        json.dump({"streams": streams,
                   "source": "ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum per launch), "
                             "profiles/%s_ncu_full_summary.txt" % tag,
                   "dram_bytes_per_launch": traffic,
                   "dram_bytes_per_stage": per_stage,
                   "dram_bytes_per_step": sum(traffic.values()),
                   "dram_bytes_per_frame": sum(traffic.values()) / streams,
                   "ncu_us_per_launch": dur}, f, indent=1)
    # hot spots + SASS of every kernel of the step
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    order = [short_name(row[1]) for row in csv.reader(src.splitlines()) if row and row[0] == "Kernel Name"]
    sass = library_sass()
    for sub in ("_hotspots", "_sass"):
        d = os.path.join(prof, tag + sub)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    done = set()
    for ki, name in enumerate(order):
        if name in done:
            continue
        done.add(name)
        with open(os.path.join(prof, tag + "_hotspots", file_name(name) + ".txt"), "w") as f:
            f.write(capture(ncu_hotspots.main, rep, 40, ki))
        lines = sass.get(name)
        if lines:
            hist = opcode_histogram(lines)
            with open(os.path.join(prof, tag + "_sass", file_name(name) + ".txt"), "w") as f:
                f.write("// %s: %d SASS instructions (cuobjdump -sass of liblane_tracker_b200.so, sm_100a)\n" % (name, len(lines)))
                f.write("// opcodes: " + ", ".join("%s %d" % kv for kv in hist.most_common()) + "\n")
                f.write("\n".join(lines) + "\n")
    if launches:
        dst = os.path.join(prof, tag + "_launches.csv")
        if os.path.abspath(launches) != os.path.abspath(dst):
            shutil.copyfile(launches, dst)
        with open(os.path.join(prof, tag + "_launches_summary.txt"), "w") as f:
            f.write(capture(launch_summary.main, dst))
    if bench:
        shutil.copyfile(bench, os.path.join(prof, tag + "_bench.json"))
    print("kernels in report:", len(rows) - 2, "; distinct:", len(done), "; DRAM MB/frame: %.1f" % (sum(traffic.values()) / streams / 1e6))


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], a[1], a[2] if len(a) > 2 else None, a[3] if len(a) > 3 else None)
