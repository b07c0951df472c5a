"""Per-kernel SASS evidence of the Blackwell-native paths in libmplu.so (runs without a GPU):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tile load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, HMMA = legacy mma.sync (must be 0)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mixed-precision_lu_factorization_b200", "libmplu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "DFMA", "MUFU.RCP"]
cnt = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cnt[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        cnt[name]["instr"] += 1
        for k in keys:
            if op.startswith(k):
                cnt[name][k] += 1
print("# cuobjdump -sass libmplu.so, per kernel: instruction count and the mnemonics that prove the sm_100a paths")
print("# (tools/sass_summary.py; UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit)")
print(f"{'kernel':88s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in keys))
for name, c in cnt.items():
    short = name if len(name) <= 88 else name[:85] + "..."
    print(f"{short:88s} {c['instr']:6d} " + " ".join(f"{c[k]:8d}" for k in keys))
