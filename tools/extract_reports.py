"""Summarise ncu reports into profiles/<tag>_ncu_summary.txt: python tools/extract_reports.py <tag> <command line that was profiled> -- rep1 rep2 ...
For every kernel of every report: duration, DRAM bytes, pipe utilisation, stall reasons (the metric list of extract_profile.py)."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from extract_profile import ROOT, WANT, rows_of, to_bytes  # noqa: E402


def main():
    tag = sys.argv[1]
    sep = sys.argv.index("--")
    cmdline = " ".join(sys.argv[2:sep])
    lines = []
    for rep in sys.argv[sep + 1:]:
        if not os.path.exists(rep):
            continue
        hdr, units, data = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        lines.append("== %s (ncu --set full --clock-control none, B200, %s)" % (os.path.basename(rep), cmdline))
        for r in data:
            lines.append("   %-90s %s " % ("Kernel Name", r[idx["Kernel Name"]]))
            for key in ("Grid Size", "Block Size"):
                if key in idx:
                    lines.append("   %-90s %s " % (key, r[idx[key]]))
            for i, h in enumerate(hdr):
                if any(re.search(w, h) for w in WANT) or h == "launch__shared_mem_per_block_dynamic":
                    lines.append("   %-90s %s %s" % (h, r[i], units[i]))
            tot = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            lines.append("   %-90s %.0f byte" % ("dram bytes read + written (this launch)", tot))
    out = os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt")
    open(out, "a").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
