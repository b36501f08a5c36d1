"""Summarise an .ncu-rep into a small markdown table (run here, no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof_tc.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.max", "SM cycles"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor active (x4 subpipes)"),
        ("sm__inst_executed_pipe_tc.sum", "tcgen05 instr"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__cycles_active.avg", "SMSP active cycles")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):            # metrics of the Triage sections carry a "UNIT.Section." prefix
        if "." in h and h.split(".", 2)[-1] not in col and h.count(".") >= 2:
            col.setdefault(h.split(".", 2)[-1], i)
    print(f"# ncu summary of `{path}` (cold-cache, serialised replays: compare shares, not absolutes)\n")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        print(f"## {name}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"| {label} (`{key}`) | {r[col[key]]} | {units[col[key]]} |")
        cyc = r[col["sm__cycles_elapsed.max"]] if "sm__cycles_elapsed.max" in col else ""
        t = r[col.get("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", 0)] if "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg" in col else ""
        try:
            if float(t) > 0:
                print(f"| tensor-pipe active fraction (hmma_cycles/4 / SM cycles) | {float(t) / 4 / float(cyc):.3f} | |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main(sys.argv[1])
