"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads /
stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA would be the legacy mma.sync path (none)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "wavthruvec_pytorch_b200" / "libv2w.so"
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "FFMA", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True).stdout.splitlines()
    per = OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it)
            cur = cur.replace("(int)", "")
            # tapgemm_tc_kernel<KC, kPair, kSplit>: the second bool is the split-operand flavour
            m3 = re.match(r"(.*tapgemm_tc_kernel<\d+, \(bool\)[01], )\(bool\)([01])>", cur)
            if m3:
                cur = m3.group(1) + ("split" if m3.group(2) == "1" else "plain") + ">"
            # mrf_stacked_kernel<C, kPair, kBlk, kX0B>, mrf_stage_kernel<C, kPair, kSub, kBlk, kX0B>: name the last two
            m4 = re.match(r"(.*(?:mrf_stacked_kernel<\d+, \(bool\)[01]|mrf_stage_kernel<\d+, \(bool\)[01], \d+), )\(bool\)([01]), \(bool\)([01])>", cur)
            if m4:
                cur = (m4.group(1) + ("any-program" if m4.group(2) == "1" else "stage-program") + ", " +
                       ("x0-blocked" if m4.group(3) == "1" else "x0-rows") + ">")
            cur = cur.replace("(bool)1", "pair").replace("(bool)0", "single")
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("v2w::", "")
            per[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            per[cur][op] += 1
            per[cur]["_total"] += 1
    arch = re.search(r"arch = (\S+)", sass)
    print(f"# cuobjdump -sass {LIB.name}  ({arch.group(1) if arch else '?'}); instructions per kernel")
    print(f"{'kernel':72s} {'total':>7s} " + " ".join(f"{w:>8s}" for w in WANT))
    tot = Counter()
    for k, c in per.items():
        print(f"{k[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[w]:8d}" for w in WANT))
        tot.update(c)
    print(f"{'TOTAL':72s} {tot['_total']:7d} " + " ".join(f"{tot[w]:8d}" for w in WANT))


if __name__ == "__main__":
    main()
