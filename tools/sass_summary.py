"""Counts the Blackwell-specific SASS mnemonics per kernel of libpinsage_b200.so -> profiles/sass_summary.txt:
tcgen05.mma (UTCHMMA / UTCQMMA / UTCIMMA), LDTM / STTM (tcgen05.ld / st: TMEM), UTMALDG / UTMASTG (TMA),
UTCBAR, SYNCS (mbarrier), ATOMS (shared-memory atomics), 256-bit global loads, REDUX.
Usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "movie-recommendation-engine_b200/libpinsage_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
pat = {"tcgen05.mma": r"UTC[HQI]MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"UTMALDG", "UTMASTG": r"UTMASTG",
       "UTCBAR": r"UTCBAR", "SYNCS": r"SYNCS", "ATOMS": r"ATOMS", "LDG.256": r"LDG\.E[^ ]*\.256", "REDUX": r"REDUX", "MATCH": r"MATCH"}
rows, cur, k = [], None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = [names[k], 0, collections.Counter()]
        rows.append(cur)
        continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        cur[1] += 1
        for key, p in pat.items():
            if re.search(p, line):
                cur[2][key] += 1
print(f"# cuobjdump -sass {lib} (sm_100a): instruction count and Blackwell-specific mnemonics per kernel")
for name, n, c in sorted(rows):
    if c:
        print(f"{name[:110]:110s} instr={n:5d}  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
print("# kernels without any of the mnemonics above:", sum(1 for r in rows if not r[2]), "of", len(rows))
