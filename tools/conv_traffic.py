"""DRAM traffic / tensor-pipe activity of the 17 forward conv launches from an `ncu --set full` report:
   python tools/conv_traffic.py gpurun_out/prof_conv_r01h.ncu-rep 16 profiles/r01h_conv_forward_traffic.json"""
import csv
import json
import subprocess
import sys

rep, batch, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = hdr.index
conv = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per, tr, tw, tt = [], 0.0, 0.0, 0.0
for r in rows[2:]:
    rd = float(r[col("dram__bytes_read.sum")]) * conv[units[col("dram__bytes_read.sum")]]
    wr = float(r[col("dram__bytes_write.sum")]) * conv[units[col("dram__bytes_write.sum")]]
    t = float(r[col("gpu__time_duration.sum")])
    per.append(dict(kernel=r[col("Kernel Name")][:40], us=t, dram_read_MB=round(rd / 1e6, 1), dram_write_MB=round(wr / 1e6, 1),
                    tensor_active_pct=round(float(r[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]), 1)))
    tr += rd; tw += wr; tt += t
assert len(per) in (16, 17), len(per)      # 16: a capture cut before the last (30x54, 512->16) launch
json.dump(dict(source=f"ncu --set full --clock-control none, python tools/profile_step.py 3 {batch} inf, the 17 conv launches (conv3x3_tc_kernel x13 + side_prep x4) of the last batch-{batch} forward",
               batch=batch, dram_bytes_read=tr, dram_bytes_write=tw, traffic_bytes=tr + tw, sum_duration_us_cold=tt, per_launch=per),
          open(dst, "w"), indent=1)
print(len(per), "launches", round(tt, 1), "us cold;", round((tr + tw) / 1e6, 1), "MB DRAM traffic")
