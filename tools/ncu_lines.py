"""Stall samples and executed instructions of one kernel in an .ncu-rep, summed per CUDA source line.
ncu's CSV source page is SASS only; the line of every SASS instruction comes from `nvdisasm --print-line-info` on the cubin of
the SAME build (the two listings hold the kernel's instructions in the same order: checked opcode by opcode).
    python tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [--top 40] [--by samples|executed]"""
import argparse
import collections
import csv
import io
import linecache
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("lib")
    ap.add_argument("kernel", help="substring of the mangled kernel name, e.g. k_lanczos_clusterILi2ELi7")
    ap.add_argument("--cubin", default="engine.sm_100a.cubin")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by", default="samples", choices=["samples", "executed"])
    ap.add_argument("--src-root", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "autoinst_b200", "csrc"))
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 5:
            cur["rows"].append(r)
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", args.cubin, os.path.abspath(args.lib)], cwd=td, capture_output=True)
        sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, args.cubin)], capture_output=True, text=True).stdout
    lines = sass.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and args.kernel in l and l.rstrip().endswith(":"))
    mangled = lines[start][6:-1]
    instrs, cur_line = [], ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("//--------------------- .text"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            instrs.append((m.group(2).strip(), cur_line))
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    k = next(k for k in kern if len(k["rows"]) == len(instrs) and all(op(a[0]) == op(r[k["hdr"].index("Source")].strip())
                                                                      for a, r in zip(instrs[:200], k["rows"][:200])))
    h = k["hdr"]
    i_s, i_e = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    samp, exe = collections.Counter(), collections.Counter()
    for (txt, ln), r in zip(instrs, k["rows"]):
        samp[ln] += int(r[i_s]); exe[ln] += int(r[i_e])
    ts, te = sum(samp.values()), sum(exe.values())
    print(f"# {k['name']}\n# {mangled}: {len(instrs)} SASS instructions, {ts} stall samples, {te} warp instructions executed")
    key = samp if args.by == "samples" else exe
    for (f, n), _ in key.most_common(args.top):
        text = linecache.getline(os.path.join(args.src_root, f), n).strip()[:100]
        print(f"samples {100 * samp[(f, n)] / ts:5.1f}%  executed {100 * exe[(f, n)] / te:5.1f}%  {f}:{n}  {text}")


if __name__ == "__main__":
    main()
