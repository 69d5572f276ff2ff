"""Attribute ncu per-SASS-instruction counters to source lines.

usage: ncu_lines.py <sass.csv from `ncu -i rep --page source --csv --print-source sass`> <nvdisasm -g -c dump> <function substring> [block index]
Prints the source lines with the most executed warp instructions and stall samples."""
import collections
import csv
import re
import sys

sass_csv, dis, fn = sys.argv[1:4]
blk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
addr2line = {}
cur = None
infn = False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = fn in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = heads[blk]
end = heads[blk + 1] if blk + 1 < len(heads) else len(rows)
hdr = rows[h]
ci = hdr.index("Instructions Executed")
si = hdr.index("# Samples")
body = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
base = int(body[0][0], 16)
inst = collections.Counter()
samp = collections.Counter()
for r in body:
    a = int(r[0], 16) - base
    key = addr2line.get(a, ("?", 0))
    inst[key] += int(r[ci] or 0)
    samp[key] += int(r[si] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print(f"{len(body)} SASS instructions, {ti} warp instructions executed, {ts} samples")
src = {}
for (f, l), n in sorted(samp.items(), key=lambda kv: -kv[1])[:45]:
    if f not in src:
        try:
            src[f] = open("/root/repo/falling_sand_engine_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print(f"{f}:{l:5d} samples {100*n/ts:5.1f}%  inst {100*inst[(f,l)]/ti:5.1f}%  {text}")

if len(sys.argv) > 5:  # group by function ranges of fse_tick_rows.cuh: name:start,...
    ranges = [(n.split(":")[0], int(n.split(":")[1])) for n in sys.argv[5].split(",")]
    g_inst, g_samp = collections.Counter(), collections.Counter()
    for (f, l), n in inst.items():
        name = f
        if f == "fse_tick_rows.cuh":
            name = "?"
            for nm, st in ranges:
                if l >= st:
                    name = nm
        g_inst[name] += n
        g_samp[name] += samp[(f, l)]
    print("-- by function")
    for nm, n in sorted(g_inst.items(), key=lambda kv: -kv[1]):
        print(f"{nm:28s} inst {100*n/ti:5.1f}%  samples {100*g_samp[nm]/ts:5.1f}%")
