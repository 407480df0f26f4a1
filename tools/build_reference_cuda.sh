#!/bin/bash
# Builds the UNMODIFIED reference (core + C API + its own CUDA device + apps) from a writable /tmp copy of
# /root/reference (its CMake project configure_file()s into the source dir, which is read-only) and copies the
# BINARIES into baseline/_ref (git-ignored, travels to the GPU box). The reference's weights/*.tza are Git-LFS
# pointers in this checkout, so synthetic TZA files of the same architectures (oidn_b200.weights.model_tza: the
# byte sizes equal the LFS sizes; passthrough=True: the network returns its input plus a small perturbation, so the
# output sanity window of apps/oidnTest.cpp holds) are placed in the copy's weights/ and compiled in as the built-in blobs.
# The CPU device stays off: it needs ISPC + oneTBB, which this image does not have.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=/tmp/oidn_ref; BLD=/tmp/oidn_build
rm -rf $SRC $BLD; mkdir -p $SRC $BLD
cp -r /root/reference/. $SRC/ && rm -rf $SRC/.git && chmod -R u+w $SRC
cd "$ROOT" && python - <<'PY'
from oidn_b200 import weights
names = """rt_alb rt_alb_large rt_hdr rt_hdr_small rt_hdr_alb rt_hdr_alb_small rt_hdr_alb_nrm rt_hdr_alb_nrm_small rt_hdr_calb_cnrm
rt_hdr_calb_cnrm_small rt_hdr_calb_cnrm_large rt_ldr rt_ldr_small rt_ldr_alb rt_ldr_alb_small rt_ldr_alb_nrm rt_ldr_alb_nrm_small
rt_ldr_calb_cnrm rt_ldr_calb_cnrm_small rt_nrm rt_nrm_large rtlightmap_hdr rtlightmap_dir""".split()
for n in names:
  kind = "small" if n.endswith("_small") else ("large" if n.endswith("_large") else "base")
  ic = 9 if ("alb_nrm" in n or "calb_cnrm" in n) else (6 if n.endswith("_alb") or "_alb_small" in n else 3)
  open("/tmp/oidn_ref/weights/%s.tza" % n, "wb").write(weights.model_tza(kind, ic, seed=0, passthrough=True))
PY
cd $BLD
cmake -G Ninja $SRC -DCMAKE_BUILD_TYPE=Release -DOIDN_DEVICE_CPU=OFF -DOIDN_DEVICE_CUDA=ON -DOIDN_DEVICE_CUDA_API=RuntimeStatic \
      -DOIDN_FILTER_RT=ON -DOIDN_FILTER_RTLIGHTMAP=ON > cmake.log 2>&1
ninja -j8 > ninja.log 2>&1
mkdir -p "$ROOT/baseline/_ref/lib" "$ROOT/baseline/_ref/bin"
cp -a libOpenImageDenoise.so* libOpenImageDenoise_core.so* libOpenImageDenoise_device_cuda.so* "$ROOT/baseline/_ref/lib/"
cp oidnBenchmark oidnTest oidnDenoise "$ROOT/baseline/_ref/bin/"
echo "reference build installed in $ROOT/baseline/_ref"
