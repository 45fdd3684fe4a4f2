#!/usr/bin/env python
"""SASS instruction census of the built extension (no GPU needed): for every kernel, the number of
instructions and of the mnemonics that prove which hardware paths it uses -- UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA tensor load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), MUFU (special function unit), ATOM/RED.

    python benchmarks/sass_census.py > profiles/sass_census.txt
    python benchmarks/sass_census.py --excerpt dft_gemm_kernel > profiles/sass_dft_gemm_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

COLS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU",
        "STG/ST", "ATOM/RED"]


def excerpt(so: str, kernel: str) -> None:
    """Only the tcgen05 / TMA / mbarrier / TMEM / global-store instructions of one kernel."""
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"UTMA|UTC|SYNCS|LDTM|STTM|\bSTG|\bST\.E|FENCE|MEMBAR|ERRBAR")
    print(f"{kernel} -- tcgen05 / TMA / mbarrier / TMEM-load / global-store instructions "
          f"(full listing: cuobjdump -sass {os.path.relpath(so)}; python benchmarks/sass_census.py --excerpt {kernel})")
    keep = False
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            keep = kernel in m.group(1)
            continue
        if keep and pat.search(line):
            print(line.rstrip())


def main():
    default_so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dfno_b200", "_build",
                              "dfno_b200_C.so")
    if len(sys.argv) > 2 and sys.argv[1] == "--excerpt":
        return excerpt(sys.argv[3] if len(sys.argv) > 3 else default_so, sys.argv[2])
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                             "dfno_b200", "_build", "dfno_b200_C.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = ins.match(line)
        if m and cur:
            op = m.group(1)
            c = counts[cur]
            c["inst"] += 1
            base = op.split(".")[0]
            if base in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU"):
                c[base] += 1
            elif base in ("STG", "ST"):
                c["STG/ST"] += 1
            elif base in ("ATOM", "ATOMG", "RED", "ATOMS"):
                c["ATOM/RED"] += 1
    names = subprocess.run(["cu++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print(f"SASS instruction census of {os.path.relpath(so)} (cuobjdump -sass, sm_100a)")
    print("UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, "
          "SYNCS = mbarrier ops\n")
    print(f"{'kernel':110s} {'inst':>7s} " + " ".join(f"{c:>7s}" for c in COLS))
    for mangled, name in sorted(zip(order, names), key=lambda t: t[1]):
        c = counts[mangled]
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        print(f"{short[:110]:110s} {c['inst']:7d} " + " ".join(f"{c[k]:7d}" for k in COLS))


if __name__ == "__main__":
    main()
