#!/bin/bash
# The reference's own CUDA device (devices/cuda: CUTLASS 2.x Sm80 mma.sync kernels, compiled for sm_100) on this
# box, through the reference's own oidnBenchmark: the kernel-to-beat number of SURVEY.md section 8(d).
# baseline/_ref is built by tools/build_reference_cuda.sh (cmake on a /tmp copy of /root/reference with synthetic
# weights/*.tza of the same architectures; binaries only, git-ignored).
export LD_LIBRARY_PATH=$PWD/baseline/_ref/lib
B=baseline/_ref/bin/oidnBenchmark
mkdir -p gpurun_out
{
$B --ld
$B -d cuda -r "RT\.hdr_alb_nrm\.(1920x1080|3840x2160|1280x720)" -q high
$B -d cuda -r "RT\.hdr_alb_nrm\.3840x2160" -q balanced
$B -d cuda -r "RT\.hdr_alb_nrm\.3840x2160" -q high -t half
$B -d cuda -r "RT\.hdr_calb_cnrm\.3840x2160" -q high
$B -d cuda -r "RT\.hdr_calb_cnrm\.1920x1080" -q high -s 7680 4320
$B -d cuda -r "RTLightmap\.hdr\.4096x4096" -q high
$B -d cuda -r "RT\.hdr_alb_nrm\.3840x2160" -q high --buffer hostcopy
} > gpurun_out/ref_cuda_bench.log 2>&1
cat gpurun_out/ref_cuda_bench.log
