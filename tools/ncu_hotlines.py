#!/usr/bin/env python
"""Per-function / per-source-line / per-opcode shares of the executed warp instructions of one kernel, from an .ncu-rep captured
with --import-source on and the object file the library was built from (compiled with -lineinfo):
   python tools/ncu_hotlines.py gpurun_out/x.ncu-rep feature_tracker_b200/csrc/klt.o KltKernelILi1ELi1ELi16E [features]
ncu's SASS page carries the per-instruction execution counts, nvdisasm -gi the (inlined) source lines of the same instructions; the
two are joined by instruction offset.  `features` scales the counts to "per feature"."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kernel = sys.argv[1:4]
    per = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.strip().startswith(".section") and kernel in l and ".text." in l][0]
    instrs, frames, pending = [], [], []
    for l in dis[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            if pending:
                frames, pending = pending, []
            instrs.append((int(m.group(1), 16), m.group(2).strip(), list(frames)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = [(int(r[ia], 16), int(r[ie]), int(r[it])) for r in rows[2:] if len(r) == len(hdr)]
    base = data[0][0]
    by = {a - base: (e, t) for a, e, t in data}
    if len(by) != len(instrs):
        print(f"warning: {len(by)} profiled instructions vs {len(instrs)} disassembled (different build?)")
    # function ranges of the .cu / .cuh files next to the object
    funcs = {}
    src_dir = os.path.dirname(os.path.abspath(obj))
    for f in os.listdir(src_dir):
        if f.endswith((".cu", ".cuh")):
            fl = []
            for i, l in enumerate(open(os.path.join(src_dir, f)).read().split("\n"), 1):
                m = re.match(r"\s*(?:__device__|__global__|struct|template)?.*?\b(?:struct\s+)?([A-Z][A-Za-z0-9]+)\b\s*(?:\(|\{|$)", l)
                if m and (l.startswith(("__device__", "__global__", "struct")) or (l.startswith("    __device__") and "(" in l)):
                    fl.append((i, m.group(1)))
            funcs[f] = fl

    def func_of(f, line):
        name = None
        for s, n in funcs.get(f, []):
            if s <= line + 1:
                name = n
        return f"{f}:{name}"

    tot = sum(e for e, _ in by.values())
    thr = sum(t for _, t in by.values())
    agg_outer, agg_inner, agg_line, ops, thr_outer, thr_inner = (collections.Counter() for _ in range(6))
    for off, txt, fr in instrs:
        e, t = by.get(off, (0, 0))
        inner = fr[0] if fr else ("?", 0)
        # the innermost frame that lies in the kernel's own translation unit names the "phase"
        outer = next((func_of(f, l) for f, l in fr if f.endswith(".cu")), "?")
        agg_outer[outer] += e
        thr_outer[outer] += t
        thr_inner[func_of(*inner)] += t
        agg_inner[func_of(*inner)] += e
        agg_line[(outer, inner)] += e
        o = txt.split()
        ops[(o[1] if o[0].startswith("@") else o[0]).split(".")[0]] += e
    print(f"kernel {kernel}: {tot / per:.0f} warp instructions (per unit), {thr / max(tot, 1):.1f} active lanes on average")
    print("-- by phase (innermost frame in the .cu file)")
    for k, v in agg_outer.most_common(16):
        print(f"  {v / tot * 100:6.2f}%  {v / per:10.0f}  {thr_outer[k] / max(v, 1):5.1f} lanes  {k}")
    print("-- by innermost function")
    for k, v in agg_inner.most_common(16):
        print(f"  {v / tot * 100:6.2f}%  {v / per:10.0f}  {thr_inner[k] / max(v, 1):5.1f} lanes  {k}")
    print("-- by opcode")
    for k, v in ops.most_common(24):
        print(f"  {v / tot * 100:6.2f}%  {k}")
    print("-- hottest (phase, source line) pairs")
    for (o, i), v in agg_line.most_common(40):
        print(f"  {v / tot * 100:6.2f}%  {o:40s} {i[0]}:{i[1]}")


if __name__ == "__main__":
    main()
