#!/usr/bin/env python
"""Attribute an ncu SASS profile to CUDA source lines.
usage: python profiles/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [--top N] [--ranges a-b,c-d:label ...]
Matches the SASS rows of `ncu --page source --csv` (address order) with `nvdisasm --print-line-info`
of the same object and sums executed warp instructions / stall samples per source line."""
import csv, io, os, re, subprocess, sys, tempfile, collections

def disasm(obj, kern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    out, on, cur = [], False, ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out

def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    sass = disasm(obj, kern)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    body = rows[2:]
    ie, ns, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall_cols = {n: hdr.index(n) for n in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math",
                                            "stall_branch_resolving", "stall_no_inst", "stall_dispatch",
                                            "stall_lg", "stall_mio", "stall_barrier", "stall_sleep") if n in hdr}
    stall_tot = collections.Counter()
    per_line_stall = collections.defaultdict(collections.Counter)
    assert len(body) == len(sass), (len(body), len(sass))
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r, (off, text, loc) in zip(body, sass):
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        op = op.split(".")[0]
        e, s = int(r[ie].replace(",", "")), int(r[ns].replace(",", ""))
        per_line[loc][0] += e
        per_line[loc][1] += s
        per_line[loc][2][op] += e
        for n, ci in stall_cols.items():
            v = int(r[ci].replace(",", "") or 0)
            stall_tot[n] += v
            per_line_stall[loc][n] += v
        tot_i += e
        tot_s += s
    print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
    print(f"{'inst%':>6} {'smp%':>6}  location                      top opcodes")
    for loc, (e, s, ops) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        o = " ".join(f"{k}:{v * 100 // max(e, 1)}%" for k, v in ops.most_common(4))
        print(f"{e / tot_i * 100:6.2f} {s / max(tot_s, 1) * 100:6.2f}  {loc[0]}:{loc[1]:<5d}  {o}")
    if "--stalls" in sys.argv:
        for n in ("stall_long_sb", "stall_short_sb", "stall_wait"):
            print(f"-- top lines for {n} (total {stall_tot[n]:,} = {stall_tot[n] / max(tot_s, 1) * 100:.1f}% of samples)")
            for loc, c in sorted(per_line_stall.items(), key=lambda kv: -kv[1][n])[:12]:
                print(f"   {c[n] / max(stall_tot[n], 1) * 100:5.1f}%  {loc[0]}:{loc[1]}")
    # per-file line ranges given as file:a-b=label
    for a in sys.argv[4:]:
        m = re.match(r"([\w.]+):(\d+)-(\d+)=(.*)", a)
        if not m:
            continue
        f, lo, hi, label = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        e = sum(v[0] for k, v in per_line.items() if k[0] == f and lo <= k[1] <= hi)
        s = sum(v[1] for k, v in per_line.items() if k[0] == f and lo <= k[1] <= hi)
        ops = collections.Counter()
        for k, v in per_line.items():
            if k[0] == f and lo <= k[1] <= hi:
                ops.update(v[2])
        o = " ".join(f"{k}:{v * 100 // max(e, 1)}%" for k, v in ops.most_common(8))
        print(f"RANGE {label:28s} inst {e / tot_i * 100:6.2f}%  samples {s / max(tot_s, 1) * 100:6.2f}%   {o}")

if __name__ == "__main__":
    main()
