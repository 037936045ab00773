"""Summarise an `ncu --set full` capture (.ncu-rep, first kernel in it) into the metrics DESIGN.md quotes, and - for the
body kernel - refresh profiles/body_kernel_traffic.json (DRAM bytes per launch, stamped with the hash of the kernel's
sources so that bench.py refuses it once the kernel changes).
    python tools/ncu_metrics.py <report.ncu-rep> "<title line>" [--traffic <batch>] > profiles/<name>_metrics.txt"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ["dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed"]
print(title)
got = {}
for name in WANT:
    if name in hdr:
        i = hdr.index(name)
        got[name] = (vals[i], units[i])
        print(f"{name:75s} {units[i]:14s} {vals[i]}")
if "--traffic" in sys.argv:
    import bench
    def to_bytes(name):
        v, u = got[name]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    out = {"kernel": "body2_umma_kernel<false>", "batch": int(sys.argv[sys.argv.index("--traffic") + 1]),
           "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
           "src_hash": bench.source_hash(), "source": f"profiles/{os.path.basename(rep)} (ncu --set full --clock-control none, one launch)"}
    with open(os.path.join(ROOT, "profiles", "body_kernel_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
