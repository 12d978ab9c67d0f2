"""Hottest SASS instructions (warp-stall samples) of one launch of an ncu report, with the source line
when the build carried -lineinfo.  usage: ncu_hot.py report.ncu-rep <launch index> [top N]"""
import csv
import subprocess
import sys


def main():
    rep, idx = sys.argv[1], int(sys.argv[2])
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines())]
    print(rows[0][1][:120])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
        s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
        print(r[ix["# Samples"]].rjust(6), r[ix["Address"]][-5:], r[ix["Source"]].strip()[:64].ljust(64), s)


if __name__ == "__main__":
    main()
