#!/usr/bin/env python
"""Instruction / stall-sample budget of one kernel by source line, from an exported ncu SASS page and the cubin's line info.

  python tools/ncu_by_source.py SOURCE_SASS.csv.gz SECTION KERNEL_SUBSTRING [--depth D] [--so PATH]

The .ncu-rep never leaves the GPU box (64 MiB limit), only its `--page source --print-source sass` CSV does; the line table is
recovered here: cuobjdump -xelf the engine .so, nvdisasm -gi -c, match instruction i of the kernel with row i of the CSV.
--depth 0 groups by the line of the kernel body (outermost inline frame), 1 by the next inline level, ...; -1 = innermost line.
"""
import argparse, csv, gzip, os, re, subprocess, sys, tempfile, collections

ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("section", type=int); ap.add_argument("kernel")
ap.add_argument("--depth", type=int, default=0)
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plasticinelab_b200", "libplb_b200.so"))
ap.add_argument("--top", type=int, default=40)
args = ap.parse_args()

rows = list(csv.reader(gzip.open(args.csv, "rt")))
secs, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        hdr, j, body = rows[i + 1], i + 2, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        secs.append((rows[i][1], hdr, body)); i = j
    else:
        i += 1
name, hdr, body = secs[args.section]
ci = {h: k for k, h in enumerate(hdr)}

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(args.so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
start = [k for k, l in enumerate(sass) if l.startswith(".text.") and args.kernel in l]
if not start:
    sys.exit("kernel not found in the cubin")
inst, stack, pending = [], [], []
for l in sass[start[0] + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        if pending:
            stack, pending = pending, []
        inst.append((l.split("*/", 1)[1].strip(), list(stack)))
if len(inst) != len(body):
    print(f"warning: {len(inst)} instructions in the cubin vs {len(body)} in the ncu page (different build?)", file=sys.stderr)
S, E = ci["# Samples"], ci["Instructions Executed"]
tot_s = sum(int(r[S]) for r in body); tot_e = sum(int(r[E]) for r in body)
agg = collections.OrderedDict()
for (txt, st), r in zip(inst, body):
    if not st:
        key = ("?", 0)
    elif args.depth < 0:
        key = st[0]
    else:
        key = st[max(len(st) - 1 - args.depth, 0)]          # stack is innermost first
    a = agg.setdefault(key, [0, 0, 0]); a[0] += int(r[E]); a[1] += int(r[S]); a[2] += 1
print(f"{name[:100]}\n{len(body)} SASS instructions, {tot_e} warp instructions executed, {tot_s} stall samples; depth {args.depth}")
print(f"{'file:line':34s} {'sass':>5s} {'warp inst':>11s} {'%':>6s} {'samples':>8s} {'%':>6s}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
    print(f"{key[0] + ':' + str(key[1]):34s} {a[2]:5d} {a[0]:11d} {100 * a[0] / tot_e:6.1f} {a[1]:8d} {100 * a[1] / max(tot_s, 1):6.1f}")
