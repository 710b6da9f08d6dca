"""Developer tool: attribute an ncu capture's warp-stall samples and executed instructions to CUDA source lines.

ncu's CSV source page lists SASS without file:line; `nvdisasm -g` of the shipped cubin lists the same SASS with line info.
The two listings are matched instruction by instruction (the opcode sequence must agree), then summed per source line.

Usage: ncu_hot_lines.py <report.ncu-rep> <cubin name inside libhorae_gpu.so, e.g. fused_scan> <substring of the mangled kernel name> [top=25] [launch=-1]
(launch = index of the captured launch inside the report, default the last one)
Needs the CUDA toolkit (ncu, cuobjdump, nvdisasm) and a library built with -lineinfo (the Makefile's default)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "horaedb_b200", "csrc", "libhorae_gpu.so")


def main():
    rep, cubin, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    which = int(sys.argv[5]) if len(sys.argv) > 5 else -1
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, check=True, stdout=subprocess.DEVNULL)
        path = [f for f in os.listdir(td) if f.startswith(cubin + ".")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, path)], capture_output=True, text=True, check=True).stdout.split("\n")
    start = [i for i, ln in enumerate(dis) if ln.startswith(".text.") and kernel in ln][0]
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.") or dis[i].startswith(".section")), len(dis))
    cur, ins = None, []
    for ln in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((cur, m.group(2)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]     # one block per captured launch
    b0 = starts[which if which >= 0 else len(starts) - 2]
    b1 = starts[starts.index(b0) + 1]
    hdr, data = rows[b0 + 1], rows[b0 + 2:b1]
    isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    if len(ins) != len(data) or any(a[1].split()[:1] != b[isrc].split()[:1] for a, b in zip(ins, data)):
        sys.exit(f"SASS of the library ({len(ins)} instructions) does not match the capture ({len(data)}): rebuild the commit that was profiled")
    by = collections.defaultdict(lambda: [0, 0])
    for (loc, _), r in zip(ins, data):
        by[loc][0] += int(r[ismp])
        by[loc][1] += int(r[iex])
    ts, te = sum(v[0] for v in by.values()), sum(v[1] for v in by.values())
    print(f"{len(ins)} SASS instructions, {te / 1e6:.1f} M warp instructions executed, {ts} stall samples")
    cache = {}
    for (f, n), (s, e) in sorted(by.items(), key=lambda x: -x[1][0])[:top]:
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except OSError:
                cache[f] = []
        text = cache[f][n - 1].strip()[:100] if n <= len(cache[f]) else ""
        print(f"{os.path.basename(f)}:{n:4d}  samples {100 * s / ts:5.1f} %  instructions {100 * e / te:5.1f} %  | {text}")


if __name__ == "__main__":
    main()
