"""Per-kernel SASS opcode evidence for libdcb_b200.so: `cuobjdump -sass` -> for every kernel the Blackwell-specific opcodes
(tcgen05 = UTCHMMA / LDTM / UTCBAR..., TMA = UTMALDG / UTMASTG / UBLKCP, PRMT networks, IMAD.WIDE hash) and the instruction count.
    python tools/sass_histogram.py > profiles/sass_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "deepcubea_b200", "libdcb_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "PRMT", "IMAD.WIDE",
       "ATOMG", "ATOM", "RED", "MATCH", "VOTE", "LDG", "STG", "LDS", "STS", "BAR", "UCGABAR", "ELECT", "REDUX", "SHFL")
arch = re.search(r"arch = (sm_\w+)", out)
print("# %s: %s, %d bytes; cuobjdump -sass opcode evidence per kernel (tools/sass_histogram.py)" % (os.path.basename(lib), arch.group(1) if arch else "?", os.path.getsize(lib)))
cur, ops = None, None
kernels = []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur, ops = m.group(1), collections.Counter()
        kernels.append((cur, ops))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and ops is not None:
        ops[m.group(1)] += 1
for name, ops in sorted(kernels, key=lambda kv: -sum(kv[1].values())):
    total = sum(ops.values())
    agg = collections.Counter()
    for op, c in ops.items():
        for k in KEY:
            if op == k or op.startswith(k + "."):
                agg[k] += c
                break
    full = [op for op in ops if op.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "UBLKCP"))]
    dn = demangle(name)
    depth = 0
    for pos in range(len(dn) - 1, -1, -1):                 # drop the trailing parameter list only
        depth += dn[pos] == ")"
        depth -= dn[pos] == "("
        if depth == 0 and dn[pos] == "(":
            dn = dn[:pos]
            break
    dn = dn[:160]
    print("\n%s\n  %d instructions; %s" % (dn, total, ", ".join("%s %d" % (k, agg[k]) for k in KEY if agg[k])))
    if full:
        print("  blackwell opcodes: " + ", ".join("%s x%d" % (op, ops[op]) for op in sorted(full)))
