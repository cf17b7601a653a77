"""Join an ncu SASS source page with nvdisasm line info and aggregate stall samples / instructions
per CUDA source line and per user-given line ranges.

usage: ncu_by_line.py <report.ncu-rep> <kernel regex> <lib.so> <cubin stem, e.g. rollout> <source.cu> [launch-skip]
       [first_line-last_line]   (also list every SASS instruction attributed to that source line range)
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kre, lib, stem, src = sys.argv[1:6]
    skip = sys.argv[6] if len(sys.argv) > 6 else "0"
    span = [int(x) for x in sys.argv[7].split("-")] if len(sys.argv) > 7 else None
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(stem + ".") and f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    fn = None
    line = None
    seq = collections.defaultdict(list)
    for l in sass.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn:
            seq[fn].append((line, m.group(2).strip()))
    # "<sass regex>@@<ncu kernel regex>" when the two differ (ncu matches the plain function name)
    ncu_re = kre
    if "@@" in kre:
        kre, ncu_re = kre.split("@@")
    key = [k for k in seq if re.search(kre, k)][0]
    ins = seq[key]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + ncu_re,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    assert len(data) >= len(ins), (len(data), len(ins))
    data = data[: len(ins)]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_i = 0
    for d, (ln, txt) in zip(data, ins):
        s = int(d[ci["# Samples"]] or 0)
        n = int(d[ci["Instructions Executed"]] or 0)
        tot_s += s
        tot_i += n
        b = by_line[ln]
        b[0] += s
        b[1] += n
        for c in stall_cols:
            v = int(d[ci[c]] or 0)
            if v:
                b[2][c] += v
    text = open(src).read().split("\n")
    print(f"kernel {key[:60]}: {tot_s} samples, {tot_i} warp instructions")
    allst = collections.Counter()
    for ln, (s, n, st) in by_line.items():
        allst.update(st)
    print("stall breakdown %:", {k: round(100 * v / max(tot_s, 1), 1) for k, v in allst.most_common(8)})
    for ln, (s, n, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:40]:
        t = ""
        if ln and ln[0] == os.path.basename(src) and ln[1] - 1 < len(text):
            t = text[ln[1] - 1].strip()[:80]
        print(f"{str(ln):28s} {100*s/max(tot_s,1):5.1f}% samp {100*n/max(tot_i,1):5.1f}% inst {st.most_common(2)} | {t}")


    if span:
        print(f"\nSASS of {os.path.basename(src)}:{span[0]}-{span[1]} (samples, executed, top stalls)")
        inside = [i for i, (ln, _) in enumerate(ins) if ln and ln[0] == os.path.basename(src) and span[0] <= ln[1] <= span[1]]
        if inside:
            for i in range(min(inside), max(inside) + 1):
                d = data[i]
                ln, txt = ins[i]
                s_ = int(d[ci["# Samples"]] or 0)
                n_ = int(d[ci["Instructions Executed"]] or 0)
                st = sorted(((int(d[ci[c]] or 0), c.replace("stall_", "")) for c in stall_cols), reverse=True)[:2]
                st = " ".join(f"{c}:{v}" for v, c in st if v)
                print(f"{ln[1] if ln else 0:5d} {s_:6d} {n_:9d}  {txt[:70]:70s} {st}")


if __name__ == "__main__":
    main()
