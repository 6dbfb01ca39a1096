"""Static SASS evidence of the built library (no GPU needed): per-kernel mnemonic counts of the instructions
that prove which hardware path a kernel uses (bulk-copy engine, tcgen05 / TMEM, mma.sync, packed fp32),
plus registers / spills from `cuobjdump -res-usage`.

    python tools/sass_evidence.py [path/to/libnfn_b200.so] > profiles/r02_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "normalizingflownetwork_b200", "libnfn_b200.so")

# (label, substrings that must all appear in the demangled name)
CFG2 = "ChainSpec<2, true, 0, 1, 2, 0, 1, 2, 0, 1, 2, 0>"
CFG3 = "ChainSpec<4, true, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0>"
CFG4 = "ChainSpec<1, true, 1, 1, 1, 1, 1>"
KERNELS = [
    ("cfg2 chain fwd+bwd, warp tiles (headline)", ["chain_kernel_w<", CFG2, ">, true, nfn::MathFast"]),
    ("cfg2 chain forward, warp tiles", ["chain_kernel_w<", CFG2, ">, false, nfn::MathFast"]),
    ("cfg3 density grid forward, warp tiles", ["chain_kernel_w<", CFG3, ">, false, nfn::MathFast"]),
    ("cfg4 chain fwd+bwd, warp tiles", ["chain_kernel_w<", CFG4, ">, true, nfn::MathFast"]),
    ("cfg2 chain fwd+bwd, cp.async CTA tiles (round 1, opt-in)", ["chain_kernel<", CFG2, ">, true, nfn::MathFast"]),
    ("cfg2 Dense(16->48)+chain fwd+bwd, tcgen05", ["tc5::dense_tc5_kernel<", CFG2, "16, true, nfn::MathFast"]),
    ("cfg2 Dense(16->48)+chain fwd+bwd, mma.sync", ["dense_chain_kernel<", CFG2, "16, true, nfn::MathFast"]),
    ("cfg5 MDN head fwd+bwd (K=20, d=2)", ["nfn::mdn_kernel<2, true, 4, true, nfn::MathFast"]),
    ("Dense(16->100)+MDN fwd+bwd", ["dense_mdn_kernel<20, 2, 16, true, nfn::MathFast"]),
    ("Dense(16->P)+KMN fwd+bwd", ["dense_kmn_kernel<", "true, nfn::MathFast"]),
    ("hidden layer backward 16x16 tanh, mma.sync", ["dense_act_bwd_mma<16, 16, 1>"]),
    ("variational weight sample + KL", ["variational_fwd"]),
    ("split-phase peer exchange (stand-alone launch)", ["peer_allreduce_kernel"]),
]
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH", "SYNCS", "FENCE", "LDGSTS", "LDG", "STG", "LDS", "STS", "RED", "ATOM",
         "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "USETMAXREG", "HMMA", "FFMA2", "FADD2", "FMUL2",
         "FFMA", "MUFU", "SHFL", "BAR", "ACQBULK", "ERRBAR"]


def sh(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True).stdout


def main():
    sass = sh("cuobjdump", "-sass", LIB)
    funcs, cur = {}, None
    ins = re.compile(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), collections.Counter())
            continue
        m = ins.match(line) if cur is not None else None
        if m:
            cur[m.group(1)] += 1
    names = list(funcs)
    dem = dict(zip(names, sh("c++filt", *names).splitlines()))
    res = {}
    lines = sh("cuobjdump", "-res-usage", LIB).splitlines()
    for i, line in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", line)
        if m and i + 1 < len(lines):
            res[m.group(1)] = lines[i + 1].strip()
    print("# SASS evidence (round 2), `python tools/sass_evidence.py` = `cuobjdump -sass / -res-usage` of the built library\n")
    print("%d kernels in `libnfn_b200.so` (sm_100a).  Static instruction counts per kernel; families are summed over"
          " their suffixes (`LDS` = `LDS` + `LDS.128` + ...).  No GPU involved.\n" % len(names))
    print("Reading aid: in the warp-tile chain kernels every `STG...STRONG.SYS` / `LDG...STRONG.SYS` belongs to the ONE CTA of the"
          " grid that carries the split-phase peer exchange (unrolled pushes to / polls of up to 8 peers); the tile loop itself"
          " moves `t` / `dt` only through `UTMALDG` / `UTMASTG` (or `UBLKCP` for row widths without a swizzle mode) and writes"
          " `logp` with one plain `STG` per row.\n")
    for label, needles in KERNELS:
        hit = [n for n in names if all(s in dem[n] for s in needles)]
        if not hit:
            print("## %s\n\n(not in this build)\n" % label)
            continue
        n = hit[0]
        c = funcs[n]
        fam = collections.Counter()
        for op, k in c.items():
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    fam[w] += k
                    break
        print("## %s\n" % label)
        print("`%s`  " % re.sub(r"\((?!anonymous).*$", "", dem[n]).replace("void ", ""))
        print("%d instructions; %s\n" % (sum(c.values()), res.get(n, "").replace("  ", " ")))
        print("| " + " | ".join(w for w in WATCH if fam[w]) + " |")
        print("|" + "---|" * sum(1 for w in WATCH if fam[w]))
        print("| " + " | ".join(str(fam[w]) for w in WATCH if fam[w]) + " |")
        detail = [op for op in sorted(c) if any(op.startswith(p) for p in ("UTMA", "UBLKCP", "UTC", "LDTM", "STTM", "SYNCS",
                                                                             "HMMA", "FFMA2", "FADD2", "FMUL2", "MUFU", "USETMAXREG", "STG", "LDG", "LDGSTS",
                                                                             "RED", "ATOM"))]
        if detail:
            print("\n" + ", ".join("`%s` x%d" % (op, c[op]) for op in detail))
        print()


if __name__ == "__main__":
    main()
