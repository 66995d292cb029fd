"""profiles/r1_ncu_tc_kernels_full_summary.csv from an `ncu --set full` report of one bench step.

usage: python tools/ncu_summary.py gpurun_out/prof_r1c.ncu-rep > profiles/r1_ncu_tc_kernels_full_summary.csv
Reads the report with `ncu -i <rep> --page raw --csv` (two header rows: names, units), keeps the columns the
roofline discussion in DESIGN.md uses, and labels the 11 tensor-core launches of an IAN_simple step in launch order.
"""
import csv, io, subprocess, sys

LAYERS = ["enc_conv1(conv1_tc)", "enc_conv2", "enc_conv3", "enc_conv4", "enc_fc1(splitK)", "enc_head(splitK)",
          "l_dec_fc2(splitK)", "dec_conv1(stream-K)", "dec_conv2", "dec_conv3", "dec_out(decout_tc)"]
KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
names, units, data = rows[0], rows[1], rows[2:]
idx = [names.index(k) for k in KEEP]
out = csv.writer(sys.stdout, lineterminator="\n")
out.writerow(["layer"] + ["%s [%s]" % (names[i], units[i]) for i in idx])
for k, r in enumerate(data):
    out.writerow([LAYERS[k] if len(data) == len(LAYERS) else "launch%d" % k] + [r[i] for i in idx])
