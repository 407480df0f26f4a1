#!/bin/bash
# The reference's own oidnBenchmark and oidnTest (unmodified sources) running on oidn_b200 through the reference's
# public API: baseline/_b200 is built by tools/build_integration_module.sh. It is listed in .gpurunignore (109 MB):
# comment that line out before sending this script to the GPU box.
export LD_LIBRARY_PATH=$PWD/baseline/_b200/lib
export OIDN_B200_WEIGHTS_DIR=$PWD/baseline/_b200/weights
mkdir -p gpurun_out
cp -f oidn_b200/liboidn_b200.so baseline/_b200/lib/   # the library as built now, not as it was when the module was linked
{
baseline/_b200/bin/oidnBenchmark --ld
timeout 30 baseline/_b200/bin/oidnBenchmark -d cuda -r "RT\.hdr_alb_nrm\.(1920x1080|3840x2160)" -q high
timeout 20 baseline/_b200/bin/oidnBenchmark -d cuda -r "RTLightmap\.hdr\.4096x4096"
timeout ${1:-40} baseline/_b200/bin/oidnTest --device cuda
echo "oidnTest exit=$?"
} > gpurun_out/integration_run.log 2>&1
tail -40 gpurun_out/integration_run.log
