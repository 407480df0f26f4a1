#!/usr/bin/env python
"""Summarises an `ncu --page raw --csv` dump of the launches of one 4K frame: writes profiles/<name>.md (one row per
launch: time, tensor pipe, DRAM, shared-memory data pipe split into tensor-core operand reads and LSU traffic) and
profiles/conv_traffic.json (DRAM bytes per conv launch, read by bench.py for roofline.traffic).
usage: ncu_summarise.py <raw.csv> <name> [layer names of the conv launches, comma separated]"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# conv launches of the base UNet frame with fused pairs (device parameter fusePairs, default on)
LAYERS = ["enc_conv0+enc_conv1 (pair)", "enc_conv2", "enc_conv3", "enc_conv4", "enc_conv5a", "enc_conv5b", "dec_conv4a", "dec_conv4b",
          "dec_conv3a", "dec_conv3b", "dec_conv2a", "dec_conv2b", "dec_conv1a", "dec_conv1b+dec_conv0+output (pair)"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3}


def main():
  rows = list(csv.reader(open(sys.argv[1])))
  hdr, units, data = rows[0], rows[1], rows[2:]
  name = sys.argv[2]
  layers = sys.argv[3].split(",") if len(sys.argv) > 3 else LAYERS

  def col(r, key, default=None):
    if key not in hdr:
      return default
    i = hdr.index(key)
    try:
      return float(r[i]) * SCALE.get(units[i], 1.0)
    except ValueError:
      return default

  out = ["# ncu --set full --clock-control none: the launches of one 4K RT hdr+alb+nrm frame (cold-cache, serialised)", "",
         "tensor % = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed; dram % = gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed;",
         "smem tc % / smem lsu % = l1tex__data_pipe_{tc,lsu}_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed (the shared-memory data pipe:",
         "tensor-core operand reads vs ld/st.shared of the epilogues); time = gpu__time_duration.sum.", "",
         "| # | kernel | layer | grid x block | regs | time us | tensor % | dram % | smem tc % | smem lsu % | dram read MB | dram write MB |",
         "|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
  # the capture window may start with the tail of the previous frame: conv launches before the first elementwise
  # kernel (autoexposure / input process) are the LAST layers, the ones after it the frame's layers in order
  names = [r[hdr.index("Kernel Name")].split("(")[0] for r in data]
  is_conv = [n.startswith("conv3x3") for n in names]
  lead = 0
  while lead < len(data) and is_conv[lead]:
    lead += 1
  tot, nconv, ci = 0.0, 0, 0
  for i, r in enumerate(data):
    kn = names[i].replace("unnamed>::", "")
    conv = is_conv[i]
    layer = ""
    if conv:
      if i < lead:
        layer = layers[len(layers) - lead + i] + " [previous frame]"
      else:
        layer = layers[ci] if ci < len(layers) else "?"
        ci += 1
    rd, wr = col(r, "dram__bytes_read.sum", 0.0), col(r, "dram__bytes_write.sum", 0.0)
    if conv:
      tot += rd + wr; nconv += 1
    f = lambda v: "-" if v is None else "%.1f" % v
    out.append("| %d | %s | %s | %s x %s | %s | %s | %s | %s | %s | %s | %.1f | %.1f |" % (
      i, kn, layer, r[hdr.index("Grid Size")], r[hdr.index("Block Size")], f(col(r, "launch__registers_per_thread")),
      f(col(r, "gpu__time_duration.sum")), f(col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")),
      f(col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
      f(col(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")),
      f(col(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")), rd / 1e6, wr / 1e6))
  out += ["", "Total DRAM traffic of the %d conv launches: %.3f GB." % (nconv, tot / 1e9)]
  open(os.path.join(ROOT, "profiles", name + ".md"), "w").write("\n".join(out) + "\n")
  if nconv:
    json.dump({"dram_bytes_per_launch_avg": tot / nconv, "dram_bytes_per_frame": tot, "launches": nconv, "source": name + ".md"},
              open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"))
  print("\n".join(out))


if __name__ == "__main__":
  main()
