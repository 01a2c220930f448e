#!/usr/bin/env python
"""Regenerate profiles/rNN_sass_evidence.txt from the in-tree library (no GPU needed):

    python profiles/sass_evidence.py r02

For every kernel of libmcs_b200.so: counts of the SASS mnemonics that prove Blackwell-native / special code paths
(tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, cp.async -> LDGSTS, redux.sync ->
REDUX, match.any -> MATCH, cluster barrier -> UCGABAR_ARV / UCGABAR_WAIT, shared atomics -> ATOMS), plus the first
UTMALDG / UTCHMMA lines of the dense kernel."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "montecarlosolvers_b200", "libmcs_b200.so")
WANT = re.compile(r"^(UTCHMMA|UTMALDG|LDTM|UTCBAR|LDGSTS|REDUX|MATCH|UCGABAR|ATOMS|SYNCS|MEMBAR|CCTL|ERRBAR)")


def main(rnd):
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    per, lines, name = collections.OrderedDict(), {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", m.group(1))
            per[name] = collections.Counter()
            lines[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and name and WANT.match(m.group(1)):
            per[name][m.group(1)] += 1
            if m.group(1).startswith(("UTMALDG", "UTCHMMA")) and len(lines[name]) < 6:
                lines[name].append(ln[:110])
    path = os.path.join(HERE, rnd + "_sass_evidence.txt")
    with open(path, "w") as f:
        f.write("# SASS evidence (cuobjdump -sass montecarlosolvers_b200/libmcs_b200.so, sm_100a), %s; python profiles/sass_evidence.py\n" % rnd)
        f.write("# tcgen05.mma -> UTCHMMA, TMA (cp.async.bulk.tensor) -> UTMALDG, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR,\n"
                "# cp.async -> LDGSTS, redux.sync -> REDUX, match.any -> MATCH, barrier.cluster -> UCGABAR_ARV / UCGABAR_WAIT\n"
                "# (B200_PROFILING.md, 'What proves a Blackwell-native kernel')\n\n")
        for k, c in per.items():
            if not c:
                continue
            f.write(k + "\n    " + ", ".join("%s x%d" % kv for kv in sorted(c.items())) + "\n")
            for ln in lines[k]:
                f.write("  " + ln + "\n")
    print(path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
