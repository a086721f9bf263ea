"""Extract the judged metrics from ncu reports into profiles/ (text summary + traffic.json)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
        r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^launch__registers_per_thread$", r"^launch__grid_size$", r"^launch__block_size$",
        r"^sm__pipe_tensor_subpipe_dmma_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
        r"^sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^lts__t_sector_hit_rate\.pct$",
        r"^sm__cycles_active\.(avg|max|min)$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^smsp__average_warps_issue_stalled_(wait|math_pipe_throttle|long_scoreboard|barrier|short_scoreboard)_per_issue_active\.ratio$",
        r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$"]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    traffic = {}
    lines = []
    for name, key in (("prof_gram", "gram_dram_bytes"), ("prof_coldot", "xupdate_dram_bytes"),
                      ("prof_tvfused", "tv_fused_dram_bytes"), ("prof_onepass", "onepass_dram_bytes")):
        rep = os.path.join(ROOT, "gpurun_out", name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        hdr, units, data = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        total = 0.0
        for r in data:
            lines.append("== %s: %s" % (name, r[idx["Kernel Name"]]))
            for i, h in enumerate(hdr):
                if any(re.search(w, h) for w in WANT):
                    lines.append("   %-90s %s %s" % (h, r[i], units[i]))
            total += to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        # x-update = two launches (both captured); gram = one launch
        traffic[key] = total if key == "xupdate_dram_bytes" else total / max(len(data), 1)
    open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("\n".join(lines))
    print(traffic)


if __name__ == "__main__":
    main()
