"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMA paths (profiles/r02_sass_tcgen05.txt).
usage: python scripts/sass_counts.py [lib.so] > profiles/r02_sass_tcgen05.txt   (cuobjdump from /usr/local/cuda)"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "mfm_b200/libmfm_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MNEM = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FHFMA", "F2FP", "LDGSTS", "UBLKCP", "ATOMG", "REDG", "RED."]
kern = None; counts = collections.OrderedDict(); n_ins = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    if kern is None or "/*" not in line: continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m: continue
    op = m.group(1); n_ins[kern] += 1
    for k in MNEM:
        if op.startswith(k): counts[kern][k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: instruction counts per kernel (static SASS, sm_100a)")
print("# UTC*MMA = tcgen05.mma (UTCHMMA: kind::f16 / kind::tf32 dense MMA issue), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = warp-level mma.sync, FHFMA = fma.rn.f32.f16 (mixed-precision FMA of the operand splitter)")
rows = []
for (k, c), d in zip(counts.items(), dem):
    if not any(c[m] for m in MNEM): continue
    short = re.sub(r"\s+", " ", d)
    short = short if len(short) < 150 else short[:147] + "..."
    rows.append((short, n_ins[k], c))
cols = [m for m in MNEM if any(r[2][m] for r in rows)]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for short, n, c in sorted(rows, key=lambda r: -sum(r[2][m] for m in ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "UTMALDG"))):
    print(f"| `{short}` | {n} | " + " | ".join(str(c[m]) for m in cols) + " |")
