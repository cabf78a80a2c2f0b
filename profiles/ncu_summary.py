"""Key metrics of ncu --set full captures (gpurun_out/*.ncu-rep, read here with `ncu -i`) -> a text summary for profiles/.
    python profiles/ncu_summary.py out.txt rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers/thread"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__cycles_active.avg", "SM active cycles (avg)"), ("sm__cycles_elapsed.avg", "elapsed cycles (avg)")]

out = open(sys.argv[1], "w")
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out.write(f"== {rep.split('/')[-1]}\n   kernel: {vals[col['Kernel Name']]}\n")
    for key, label in WANT:
        if key in col:
            out.write(f"   {label:42s} {vals[col[key]]} {units[col[key]]}\n")
    out.write("\n")
out.close()
print(open(sys.argv[1]).read())
