"""SASS excerpt of the production kernel cloth_step_kernel<float,128,25,false,false> for profiles/: the TMA bulk-copy /
mbarrier sites, the instruction mix, and the hottest serial loop (the limit-queue pop of limit_replay).
Usage: python scripts/sass_excerpt.py > profiles/rXX_sass_excerpt.txt   (needs the built objects, no GPU)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "gym_cloth_b200", "csrc", "_build", "cloth_inst_f32_w25.o")
KERNEL = "_ZN9clothb20017cloth_step_kernelIfLi128ELi25ELb0ELb0EEEvNS_9DevParamsIT_EENS_8StepArgsIS2_EE"

sass = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if "Function : " + KERNEL in l)
end = next((i for i in range(start + 1, len(sass)) if "Function : " in sass[i]), len(sass))
ins = []
for l in sass[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
res = subprocess.run(["cuobjdump", "-res-usage", OBJ], capture_output=True, text=True).stdout.split("\n")
usage = next((res[i + 1].strip() for i, l in enumerate(res) if KERNEL in l and i + 1 < len(res)), "")
print("kernel: clothb200::cloth_step_kernel<float, 128, 25, false, false>   (%s)" % os.path.relpath(OBJ, ROOT))
print("resources:", usage)
print("instructions: %d" % len(ins))
mix = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t in ins)
print("static mix:", ", ".join("%s %d" % kv for kv in mix.most_common(24)))
for name, pat in (("TMA bulk copies (cp.async.bulk -> UBLKCP)", r"UBLKCP"), ("mbarrier (SYNCS)", r"SYNCS"),
                  ("warp reductions (REDUX / CREDUX)", r"REDUX"), ("tensor cores (none expected: HMMA/UTCMMA/tcgen05)", r"HMMA|UTC|TCGEN")):
    hits = [(a, t) for a, t in ins if re.search(pat, t)]
    print("\n== %s: %d sites" % (name, len(hits)))
    for a, t in hits[:12]:
        print("  /*%05x*/ %s" % (a, t))
# backward branches -> loops; the limit-queue pop is the loop that holds both an ATOMS.OR (new flags) and a REDUX.MIN (next pop)
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
        j = addr[int(m.group(1), 16)]
        body = [x[1] for x in ins[j:i + 1]]
        if any("ATOMS.OR" in b for b in body) and any("REDUX" in b and "MIN" in b for b in body):
            if best is None or len(body) < best[1] - best[0]:
                best = (j, i + 1)
if best:
    print("\n== limit_replay pop loop (cloth_device.cuh, limit_replay): %d instructions per popped spring" % (best[1] - best[0]))
    for a, t in ins[best[0]:best[1]]:
        print("  /*%05x*/ %s" % (a, t))
