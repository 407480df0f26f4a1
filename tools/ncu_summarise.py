#!/usr/bin/env python
"""Summarises an `ncu --page raw --csv` dump of the conv launches of one 4K frame:
writes profiles/<name>.md (per-launch table) and profiles/conv_traffic.json (DRAM bytes per launch,
read by bench.py for roofline.traffic).  usage: ncu_summarise.py <raw.csv> <name>"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = ["enc_conv0", "enc_conv1", "enc_conv2", "enc_conv3", "enc_conv4", "enc_conv5a", "enc_conv5b", "dec_conv4a",
          "dec_conv4b", "dec_conv3a", "dec_conv3b", "dec_conv2a", "dec_conv2b", "dec_conv1a", "dec_conv1b", "dec_conv0"]

def main():
  rows = list(csv.reader(open(sys.argv[1])))
  hdr, units, data = rows[0], rows[1], rows[2:]
  ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
  scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
  ip = hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in hdr else None
  idr = hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in hdr else None
  out = ["# ncu --set full, conv3x3_tc_kernel: the 16 launches of one 4K frame (cold-cache, serialised)", "",
         "`tensor pipe %` = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed, `dram %` = gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "",
         "| launch | layer | grid | dram read MB | dram write MB | time us | tensor pipe % | dram % |", "|---|---|---|---:|---:|---:|---:|---:|"]
  tot = 0.0
  for i, r in enumerate(data[:16]):
    rd = float(r[ir]) * scale[units[ir]]; wr = float(r[iw]) * scale[units[iw]]
    tot += rd + wr
    out.append("| %d | %s | %s | %.1f | %.1f | %.1f | %s | %s |" % (i, LAYERS[i] if i < 16 else "?", r[hdr.index("Grid Size")], rd / 1e6, wr / 1e6, float(r[it]),
               ("%.1f" % float(r[ip])) if ip is not None else "-", ("%.1f" % float(r[idr])) if idr is not None else "-"))
  out += ["", "Total DRAM traffic of the 16 launches: %.3f GB (algorithmic minimum 858.2 B/px x 8.29 Mpx = 7.12 GB; "
          "the difference is L2 residency between producer and consumer launches)." % (tot / 1e9)]
  name = sys.argv[2]
  open(os.path.join(ROOT, "profiles", name + ".md"), "w").write("\n".join(out) + "\n")
  json.dump({"dram_bytes_per_launch_avg": tot / 16, "dram_bytes_per_frame": tot, "launches": 16, "source": name + ".md"},
            open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"))
  print("\n".join(out))

if __name__ == "__main__":
  main()
