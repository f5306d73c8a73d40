"""Summarise an ncu report (.ncu-rep) into the few numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/roi_fwd.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    for rep in sys.argv[1:]:
        raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
        if len(raw) < 3:
            print(f"## {rep}: no data")
            continue
        hdr, units = raw[0], raw[1]
        for row in raw[2:]:
            rec = dict(zip(hdr, row))
            print(f"## {rep} -- {rec.get('Kernel Name', '?')[:90]}\n")
            print("| metric | value | unit |\n|---|---|---|")
            for h, u, v in zip(hdr, units, row):
                if h in KEYS:
                    print(f"| {h} | {v} | {u} |")
            print()
        src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv"]))))
        if len(src) > 3:
            h = src[1]
            data = src[2:]
            ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
            tot = sum(int(r[ia]) for r in data) or 1
            tots = sum(int(r[isamp]) for r in data) or 1
            op, ops = collections.Counter(), collections.Counter()
            for r in data:
                t = r[isrc].split()
                o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
                op[o] += int(r[ia])
                ops[o] += int(r[isamp])
            print(f"SASS: {len(data)} instructions, {tot} warp-instructions executed, {tots} stall samples\n")
            print("| opcode | % executed | % samples |\n|---|---|---|")
            for k, v in op.most_common(12):
                print(f"| {k} | {100 * v / tot:.1f} | {100 * ops[k] / tots:.1f} |")
            stall = collections.Counter()
            for i, name in enumerate(h):
                if name.startswith("stall_") and "Not Issued" not in name:
                    for r in data:
                        try:
                            stall[name] += int(r[i])
                        except ValueError:
                            pass
            print("\n| stall reason | samples |\n|---|---|")
            for k, v in stall.most_common(8):
                print(f"| {k} | {v} |")
            print()


if __name__ == "__main__":
    main()
