"""Utilisation summary of ncu reports: python scripts/ncu_util_table.py report.ncu-rep [...]
Prints duration, tensor-pipe, shared-memory bank, L2 and DRAM utilisation of the (first) launch in every report."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts (tensor reads) %"),
        ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem bank reads %"),
        ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem bank writes %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from SMs"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX (incl. shared) %")]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(path, "no data")
        continue
    h, units, v = rows[0], rows[1], rows[-1]
    d = dict(zip(h, v))
    u = dict(zip(h, units))
    print(f"== {path}: {d.get('Kernel Name', '')[:90]}")
    for k, label in WANT:
        if k in d:
            print(f"   {label:34s} {d[k]:>14s} {u.get(k, '')}")
