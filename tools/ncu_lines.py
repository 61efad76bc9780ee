#!/usr/bin/env python
"""Per-source-line view of an ncu capture taken with --import-source on (developer tool):
python tools/ncu_lines.py <rep> [top_n]  -> lines sorted by stall samples, with executed instructions, lanes, shared-memory
excess wavefronts (bank conflicts) and the dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            lines.append((cur_file, int(r[0]), r[1], d))
    tot_s = sum(int(d["# Samples"]) for *_, d in lines)
    tot_e = sum(int(d["Instructions Executed"]) for *_, d in lines)
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    print(f"total samples {tot_s}, instructions {tot_e}")
    for f, ln, src, d in sorted(lines, key=lambda x: -int(x[3]["# Samples"]))[:top]:
        s, e = int(d["# Samples"]), int(d["Instructions Executed"])
        thr = int(d["Thread Instructions Executed"])
        exc = int(d.get("L1 Wavefronts Shared Excessive", "0") or 0)
        st = sorted(((c[6:], int(d[c] or 0)) for c in stall_cols), key=lambda x: -x[1])[:3]
        st = " ".join(f"{n}:{v}" for n, v in st if v)
        print(f"{f}:{ln:4d} smp {100 * s / tot_s:4.1f}% exec {100 * e / tot_e:4.1f}% lanes {thr / max(e, 1):4.1f} "
              f"bankx {exc:8d} [{st}] | {src.strip()[:70]}")


if __name__ == "__main__":
    main()
