"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full --import-source on`) into the
text file committed under profiles/.   usage: python tools/ncu_summary.py rep.ncu-rep [out.txt] [note]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# ncu summary of {rep.split('/')[-1]}", file=out)
    if note:
        print(f"# {note}", file=out)
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print(f"\n## kernel: {d.get('Kernel Name', '?')}   grid={d.get('Grid Size', '?')} block={d.get('Block Size', '?')}", file=out)
        for h, u, v in zip(hdr, units, row):
            if h in KEYS:
                print(f"{h:75s} {v:>18s} {u}", file=out)
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    if len(src) > 3:
        h = src[1]
        data = src[2:]
        i_s, i_src, i_ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        tot = sum(int(r[i_s]) for r in data if len(r) > i_s and r[i_s].isdigit())
        agg = collections.Counter()
        for r in data:
            for c in stall_cols:
                if len(r) > c and r[c].isdigit():
                    agg[h[c]] += int(r[c])
        print(f"\n## warp-stall sampling (first kernel): {tot} samples, {len(data)} SASS instructions", file=out)
        print("   " + ", ".join(f"{k}={v}" for k, v in agg.most_common(8)), file=out)
        print("   top instructions by samples:", file=out)
        top = sorted((r for r in data if len(r) > i_s and r[i_s].isdigit()), key=lambda r: -int(r[i_s]))[:14]
        for r in top:
            st = sorted(((h[c], int(r[c])) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0), key=lambda kv: -kv[1])[:2]
            print(f"   {int(r[i_s]):6d}  x{r[i_ex]:>8s}  {r[i_src].strip()[:64]:64s} {st}", file=out)


if __name__ == "__main__":
    main()
