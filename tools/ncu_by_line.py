#!/usr/bin/env python
"""Join an ncu report's SASS-level sampling data with nvdisasm line info -> per-source-line table.

    python tools/ncu_by_line.py gpurun_out/prof.ncu-rep gym_anm_b200/lib/libanm_b200.so 'anm_env_kernelILi16ELi6' [top]

(ncu's CSV source page has no CUDA-C correlation; nvdisasm -g provides it.)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_map(so, kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    txt = ""
    for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):  # one cubin per translation unit
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        if kernel_substr in txt:
            break
    m, cur, infn, fn = {}, None, False, None
    for ln in txt.splitlines():
        if ln.startswith("//---------------------"):
            infn = kernel_substr in ln
            continue
        if not infn:
            continue
        mm = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if mm:
            cur = (mm.group(1), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if mm:
            m[int(mm.group(1), 16)] = (cur, mm.group(2).strip())
    return m


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(so, kern)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel block whose name matches
    i = 0
    while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
        i += 1
    H = rows[i + 1]
    ci = {h: k for k, h in enumerate(H)}
    base = None
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_i = 0
    for r in rows[i + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) < len(H) or not r[0].startswith("0x"):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        off = addr - base
        src = lm.get(off, ((None, -1), "?"))[0] or ("?", -1)
        smp, ins = int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]])
        e = by_line[src]
        e[0] += smp
        e[1] += ins
        e[2][r[1].split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]] += smp
        tot_s += smp
        tot_i += ins
    print("kernel %s: %d samples, %d warp instructions" % (kern, tot_s, tot_i))
    print("%-28s %8s %6s %10s %6s  top opcodes by samples" % ("file:line", "samples", "%", "instr", "%"))
    for src, (smp, ins, ops) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %8d %5.1f%% %10d %5.1f%%  %s" % ("%s:%d" % src, smp, 100.0 * smp / max(tot_s, 1), ins,
                                                       100.0 * ins / max(tot_i, 1),
                                                       " ".join("%s:%d" % kv for kv in ops.most_common(4))))


if __name__ == "__main__":
    main()
