"""Per-kernel SASS digest of the shipped library: architecture, registers, and the counts of the instructions that prove what
each kernel is built from (TMA bulk tensor loads UTMALDG, mbarrier SYNCS, named barriers BAR, cp.async LDGSTS, strong /
release-ordered global accesses, shuffles, FP64 / FP32 FMA).  usage: python tools/sass_digest.py > profiles/r2_sass_digest.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "papteam_opticalflow_b200", "libpyflow_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
ops = ["UTMALDG", "SYNCS", "BAR", "LDGSTS", "MEMBAR", "STG.E.STRONG", "LDG.E.STRONG", "REDG", "ATOMG", "SHFL", "LDS", "STS", "FFMA", "DFMA", "DADD", "DMUL", "NANOSLEEP", "HMMA", "UTCHMMA"]
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print("# SASS digest of `papteam_opticalflow_b200/libpyflow_b200.so` (round 2)\n")
print("`cuobjdump -sass` of the shipped library; architectures present: %s. No tensor-core instruction (HMMA / UTCHMMA / tcgen05) anywhere: nothing on this"
      " path is a contraction. Counts are static instruction counts per kernel.\n" % ", ".join(arch))
print("| kernel | regs | " + " | ".join(ops) + " |")
print("|---|---:|" + "---:|" * len(ops))
blocks = re.split(r"\n\s*Function : ", sass)
rows = []
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    cnt = collections.Counter()
    for line in b.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for o in ops:
            if op == o or op.startswith(o + ".") or (o in ("STG.E.STRONG", "LDG.E.STRONG") and op.startswith(o)):
                cnt[o] += 1
    d = demangle(name)
    d = re.sub(r"^void ", "", d); d = d.replace("pf::", ""); d = re.sub(r"\(.*$", "", d)
    rows.append((d, regs.get(name, (0, 0))[0], cnt))
for d, r, cnt in sorted(rows):
    if not d.startswith("k_"):
        continue
    print("| `%s` | %d | " % (d, r) + " | ".join(str(cnt[o]) if cnt[o] else "" for o in ops) + " |")
