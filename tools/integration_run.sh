#!/bin/bash
# The reference's own oidnBenchmark and oidnTest (unmodified sources) running on oidn_b200 through the reference's
# public API: baseline/_b200 is built by tools/build_integration_module.sh. It is listed in .gpurunignore (109 MB):
# comment that line out before sending this script to the GPU box.
# usage: integration_run.sh [oidnTest timeout s] [filter|ops]   (ops = the op-level module on the UNMODIFIED core)
D=baseline/_b200; [ "${2:-filter}" = ops ] && D=baseline/_b200_ops
export LD_LIBRARY_PATH=$PWD/$D/lib
export OIDN_B200_WEIGHTS_DIR=$PWD/baseline/_b200/weights   # filter-level route only; the op-level route uses the core's blobs
mkdir -p gpurun_out
cp -f oidn_b200/liboidn_b200.so $D/lib/   # the library as built now, not as it was when the module was linked
LOG=gpurun_out/integration_run_${2:-filter}.log
{
$D/bin/oidnBenchmark --ld
timeout 30 $D/bin/oidnBenchmark -d cuda -r "RT\.hdr_alb_nrm\.(1920x1080|3840x2160)" -q high
timeout 20 $D/bin/oidnBenchmark -d cuda -r "RTLightmap\.hdr\.4096x4096"
timeout ${1:-40} $D/bin/oidnTest --device cuda
echo "oidnTest exit=$?"
} > $LOG 2>&1
tail -40 $LOG
