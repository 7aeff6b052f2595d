"""Writes profiles/<name>.md from an .ncu-rep of the PreSync kernel: headline metrics, stall reasons
and the share of executed instructions / stall samples per device function.
usage: python tools/make_profile_md.py REPORT.ncu-rep OUT.md "title" ["notes file"]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, title = sys.argv[1:4]
notes = open(sys.argv[4]).read() if len(sys.argv) > 4 else ""


def fn_ranges(path, label):
    lines = open(path).read().split("\n")
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"(?:__device__|__global__)[^(]*?(\w+)\(", l)
        if m and m.group(1) not in ("__launch_bounds__",):
            starts.append((i, m.group(1)))
        elif l.startswith("presync_kernel("):
            starts.append((i - 2, "presync_kernel_body"))
    args = []
    for j, (i, n) in enumerate(starts):
        end = starts[j + 1][0] - 1 if j + 1 < len(starts) else len(lines)
        args.append(f"{label}:{i}-{end}={n}")
    return args


args = fn_ranges(os.path.join(ROOT, "rs-sync_b200/csrc/engine.cu"), "engine.cu") + \
    fn_ranges(os.path.join(ROOT, "rs-sync_b200/csrc/device_math.cuh"), "device_math.cuh") + \
    ["sm_30_intrinsics.hpp:1-9999=shfl_intrinsics", "sm_80_rt.hpp:1-9999=redux_intrinsics",
     "sm_100_rt.hpp:1-9999=f32x2_intrinsics", "math_functions.hpp:1-99999=min_max", "rng.h:1-99=rng_h",
     "sm_70_rt.hpp:1-9999=nanosleep"]
summ = subprocess.run([sys.executable, os.path.join(ROOT, "profiles/ncu_summary.py"), rep], capture_output=True, text=True).stdout
summ = "\n".join(summ.split("\n")[1:])
lines_out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles/ncu_lines.py"), rep,
                            os.path.join(ROOT, "rs-sync_b200/lib/engine.o"), "presync_kernelILi7ELb0E", "--top", "0",
                            "--stalls"] + args, capture_output=True, text=True).stdout
tab = []
for l in lines_out.split("\n"):
    if l.startswith("RANGE"):
        p = l.split()
        if float(p[3].rstrip("%")) >= 0.3:
            tab.append(f"{p[1]:28s} inst {p[3]:>7s}  samples {p[5]:>7s}")
    elif l.strip():
        tab.append(l)
with open(out, "w") as f:
    f.write(f"# {title}\n\nCapture: `ncu --set full --import-source on --clock-control none -k regex:presync_kernel -s 1 -c 1 "
            f"python tools/prof_presync.py C2 2` (one launch of the C2 grid: 3300 frames x 200 rays x 201 offsets = "
            f"663 300 warp tasks).\n\n```\n{summ}```\n\n## Share of executed warp instructions / of stall samples by "
            f"function (profiles/ncu_lines.py)\n\n```\n" + "\n".join(tab) + "\n```\n\n" + notes)
