"""Per-source-line instruction / stall / shared-memory-wavefront shares of one kernel from an .ncu-rep captured with
--import-source on (ncu -i REP --page source --print-source cuda,sass --csv).  Usage: python scripts/ncu_lines.py REP [top]"""
import collections, csv, subprocess, sys, io, re

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = load(rep)
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ix = {n: i for i, n in enumerate(hdr)}
    iI, iS = ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"]
    iW, iE = ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Excessive"]
    per = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
    cur = None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if len(r) > iE and r[0].isdigit():
            try:
                v = per[(cur, int(r[0]))]
                v[0] += int(r[iI]); v[1] += int(r[iS]); v[2] += int(r[iW] or 0); v[3] += int(r[iE] or 0); v[4] = r[1].strip()[:90]
            except ValueError:
                pass
    tot = [sum(v[k] for v in per.values()) for k in range(4)]
    print("instructions %.4e  stall samples %d  smem wavefronts %.4e  excessive %.4e (%.1f %%)" % (tot[0], tot[1], tot[2], tot[3], 100.0 * tot[3] / max(tot[2], 1)))
    # group by enclosing function: read the source files for "void name(" markers
    funcs = {}
    for f in set(k[0] for k in per):
        try:
            path = next(p for p in ("gym_cloth_b200/csrc/" + f,) if open(p))
        except Exception:
            continue
        marks = []
        for n, line in enumerate(open(path), 1):
            m = re.search(r"(?:void|int|bool|T|double|unsigned|uint16_t \*)\s+(\w+)\s*\([^;]*\)\s*(?:const\s*)?\{", line)
            if m and ("__device__" in line or "__global__" in line or "template" in line or line.startswith("    __device__")):
                marks.append((n, m.group(1)))
        funcs[f] = marks
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (f, ln), v in per.items():
        name = f
        for n, nm in funcs.get(f, []):
            if n <= ln:
                name = nm
        for k in range(4):
            agg[name][k] += v[k]
    print("%-28s %8s %8s %8s %8s" % ("function", "inst%", "stall%", "wave%", "excess%"))
    for nm, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-28s %8.2f %8.2f %8.2f %8.2f" % (nm, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), 100.0 * v[2] / max(tot[2], 1), 100.0 * v[3] / max(tot[3], 1)))
    print()
    for (f, ln), v in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
        print("%-20s %5d inst %5.2f%% stall %5.2f%% wave %5.2f%% | %s" % (f, ln, 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), 100.0 * v[2] / max(tot[2], 1), v[4]))

if __name__ == "__main__":
    main()
