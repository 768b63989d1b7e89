"""profiles/rNN_sass_evidence.txt: per kernel of the shipped library, occurrences of the Blackwell-native SASS mnemonics
(cuobjdump -sass; B200_PROFILING.md "What proves a Blackwell-native kernel")."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "robo-vln_b200/librobovln_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pats = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
                                ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"(?<![A-Z])HMMA"),
                                ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR")])
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); counts[fn] = collections.Counter(); continue
    if fn:
        for k, p in pats.items():
            if re.search(p, line): counts[fn][k] += 1
print(f"SASS evidence (cuobjdump -sass {lib}, sm_100a): occurrences per kernel of the Blackwell-native mnemonics")
print("UTCHMMA = tcgen05.mma (kind::f16), UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy, LDTM/STTM = tcgen05.ld/st, "
      "UTCBAR = tcgen05.commit, HMMA = mma.sync, SYNCS = mbarrier, UCGABAR = cluster barrier\n")
for fn, c in counts.items():
    if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "HMMA")): continue
    print(demangle(fn)[:150])
    print("    " + "  ".join(f"{k}={c[k]}" for k in pats if c[k]))
