"""Turn ncu captures brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rN_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/rN_kernel.md
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("<unnamed>::", "").replace("void ", "")[:60]
        t = float(r[mv].replace(",", ""))
        seq.append((name, t))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% |\n")
        # one step: from one assembly (phase-0) launch to the next
        idx = [i for i, (n, _) in enumerate(seq) if n.startswith("nlp_phase0_kernel") or n.startswith("nlp_dyn_kernel")]
        if len(idx) > 4:
            a, b = idx[3], idx[4]
            step = seq[a:b]
            st = sum(t for _, t in step)
            f.write(f"\nOne step of the device-resident hot path ({len(step)} launches, {st / 1e6:.3f} ms under ncu):\n\n| kernel | ms | share of step |\n|---|---|---|\n")
            for n, t in step:
                f.write(f"| `{n}` | {t / 1e6:.3f} | {100 * t / st:.1f}% |\n")
    print(open(dst).read())


def kernel(src, dst, note=""):
    """src: a .ncu-rep, or the CSV that `ncu -i rep --page raw --csv` printed on the GPU box (a .ncu-rep of a dozen
    --set full captures exceeds what gpurun brings back)."""
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full ({src})\n\n{note}\n\n" if note else f"# ncu --set full ({src})\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## launch {r[0]}: `{name[:110]}`  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read())


def dram(src, dst, launch_id, points, note=""):
    """profiles/sdf_tc_kernel_dram.json (read by bench.py for roofline.traffic): DRAM bytes of ONE captured launch and the
    number of SDF points it processed."""
    import json
    rows = list(csv.reader(open(src).read().splitlines()))
    hdr, units = rows[0], rows[1]
    r = next(r for r in rows[2:] if r[0] == str(launch_id))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(k):
        i = hdr.index(k)
        return float(r[i].replace(",", "")) * scale[units[i]]
    d = {"kernel": r[hdr.index("Kernel Name")][:60], "points": int(points), "dram_bytes_read": val("dram__bytes_read.sum"),
         "dram_bytes_write": val("dram__bytes_write.sum"), "source": note or src}
    json.dump(d, open(dst, "w"), indent=1)
    print(d)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "dram": dram}[sys.argv[1]](*sys.argv[2:])
