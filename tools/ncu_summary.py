"""Summarise an `ncu --set full --import-source on` report into a small text file for profiles/ (run where ncu is installed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof_summary.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:95s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    src = ncu_csv(rep, "source")
    if len(src) < 3:
        return
    h = src[1]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data = [r for r in src[2:] if len(r) >= len(h) and r[ix["Instructions Executed"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = collections.Counter()
    for r in data:
        for k in stalls:
            agg[k[6:]] += int(r[ix[k]] or 0)
    print(f"\nwarp-state samples: {tot}")
    for k, v in agg.most_common(8):
        print(f"  {k:22s} {v:9d} {100.0 * v / max(tot, 1):5.1f} %")
    buckets = collections.OrderedDict()
    for r in data:
        e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        d = buckets.setdefault(e, [0, 0])
        d[0] += 1
        d[1] += s
    print("\ncode regions by execution count (warp-level executions of each SASS instruction):")
    for e, (n, s) in sorted(buckets.items(), key=lambda kv: -kv[1][1])[:8]:
        print(f"  executed {e:10d} x : {n:5d} SASS instructions, {100.0 * s / max(tot, 1):5.1f} % of the samples")
    print("\nhottest SASS instructions:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
        st = {k[6:]: int(r[ix[k]] or 0) for k in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"  {r[ix['Source']][:64]:64s} {r[ix['# Samples']]:>7s}  " + " ".join(f"{k}={v}" for k, v in top if v))


if __name__ == "__main__":
    main()
