#!/bin/bash
# Builds integration/b200_module.cpp against the reference tree and installs a runnable copy in baseline/_b200
# (git-ignored binaries): the reference's core + C API + apps with ONE word changed in the core (`virtual` on
# Device::newFilter, core/device.h), and THIS backend as lib OpenImageDenoise_device_cuda (the module the core
# loads for DeviceType::CUDA). The reference's own oidnBenchmark / oidnTest / oidnDenoise then run on oidn_b200:
#   LD_LIBRARY_PATH=baseline/_b200/lib OIDN_B200_WEIGHTS_DIR=baseline/_b200/weights baseline/_b200/bin/oidnBenchmark -d cuda
# Built-in weights: the reference compiles its blobs into the core; this backend reads <weightsDir>/<name>.tza, so
# the same synthetic TZA files that were compiled into the core are installed next to the binaries.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=/tmp/oidn_ref_b200; BLD=/tmp/oidn_build_b200; OUT="$ROOT/baseline/_b200"
[ -d /tmp/oidn_ref/weights ] || bash "$ROOT/tools/build_reference_cuda.sh"
rm -rf $SRC $BLD; mkdir -p $BLD
cp -r /tmp/oidn_ref $SRC
grep -q '^    Ref<Filter> newFilter(const std::string& type);' $SRC/core/device.h
sed -i 's/^    Ref<Filter> newFilter(const std::string& type);/    virtual Ref<Filter> newFilter(const std::string\& type);/' $SRC/core/device.h
cd $BLD
# the reference's own CUDA device is configured too (a device must be enabled) but only core, API and apps are built
cmake -G Ninja $SRC -DCMAKE_BUILD_TYPE=Release -DOIDN_DEVICE_CPU=OFF -DOIDN_DEVICE_CUDA=ON -DOIDN_DEVICE_CUDA_API=RuntimeStatic \
      -DOIDN_FILTER_RT=ON -DOIDN_FILTER_RTLIGHTMAP=ON > cmake.log 2>&1
ninja -j8 OpenImageDenoise OpenImageDenoise_core oidnBenchmark oidnTest oidnDenoise > ninja.log 2>&1
make -s -C "$ROOT/oidn_b200/csrc"
CXX=${CXX:-g++}
$CXX -std=c++17 -O2 -fPIC -fvisibility=hidden -fvisibility-inlines-hidden -Wall -Wno-unknown-pragmas \
     -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS \
     -I/usr/local/cuda/targets/x86_64-linux/include -I"$ROOT/include" -isystem $SRC -isystem $SRC/external -isystem $BLD \
     -shared -Wl,-soname,libOpenImageDenoise_device_cuda.so.2.4.1 -Wl,-z,now \
     -o $BLD/libOpenImageDenoise_device_cuda.so.2.4.1 "$ROOT/integration/b200_module.cpp" \
     $BLD/libOpenImageDenoise_core.so.2.4.1 -L"$ROOT/oidn_b200" -loidn_b200 \
     -L/usr/local/cuda/targets/x86_64-linux/lib -lcudart_static -lrt -lpthread -ldl \
     -Wl,-rpath,'$ORIGIN'
rm -rf "$OUT"; mkdir -p "$OUT/lib" "$OUT/bin" "$OUT/weights"
cp -a $BLD/libOpenImageDenoise.so* $BLD/libOpenImageDenoise_core.so* "$OUT/lib/"
cp $BLD/libOpenImageDenoise_device_cuda.so.2.4.1 "$OUT/lib/"; ln -sf libOpenImageDenoise_device_cuda.so.2.4.1 "$OUT/lib/libOpenImageDenoise_device_cuda.so"
cp "$ROOT/oidn_b200/liboidn_b200.so" "$OUT/lib/"
cp $BLD/oidnBenchmark $BLD/oidnTest $BLD/oidnDenoise "$OUT/bin/"
cp $SRC/weights/*.tza "$OUT/weights/"
echo "integration build (filter-level module, core with the one-word change) installed in $OUT"

# Op-level variant (-DOIDN_B200_OP_LEVEL): NO core change, so it is linked against the UNMODIFIED reference build of
# tools/build_reference_cuda.sh (/tmp/oidn_build) and only replaces that build's CUDA device module. The reference's
# own filters, graph, arena planner and tile loop drive this backend's ops; built-in weights are the core's blobs.
OPS="$ROOT/baseline/_b200_ops"; REFBLD=/tmp/oidn_build
[ -f $REFBLD/libOpenImageDenoise_core.so.2.4.1 ] || bash "$ROOT/tools/build_reference_cuda.sh"
rm -rf "$OPS"; mkdir -p "$OPS/lib" "$OPS/bin"
$CXX -std=c++17 -O2 -fPIC -fvisibility=hidden -fvisibility-inlines-hidden -Wall -Wno-unknown-pragmas \
     -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS -DOIDN_B200_OP_LEVEL \
     -I/usr/local/cuda/targets/x86_64-linux/include -I"$ROOT/include" -isystem /tmp/oidn_ref -isystem /tmp/oidn_ref/external -isystem $REFBLD \
     -shared -Wl,-soname,libOpenImageDenoise_device_cuda.so.2.4.1 -Wl,-z,now \
     -o "$OPS/lib/libOpenImageDenoise_device_cuda.so.2.4.1" "$ROOT/integration/b200_module.cpp" \
     $REFBLD/libOpenImageDenoise_core.so.2.4.1 -L"$ROOT/oidn_b200" -loidn_b200 \
     -L/usr/local/cuda/targets/x86_64-linux/lib -lcudart_static -lrt -lpthread -ldl \
     -Wl,-rpath,'$ORIGIN'
ln -sf libOpenImageDenoise_device_cuda.so.2.4.1 "$OPS/lib/libOpenImageDenoise_device_cuda.so"
for f in libOpenImageDenoise.so libOpenImageDenoise.so.2 libOpenImageDenoise.so.2.4.1 libOpenImageDenoise_core.so libOpenImageDenoise_core.so.2.4.1; do
  cp -al "$ROOT/baseline/_ref/lib/$f" "$OPS/lib/" 2>/dev/null || cp -a "$ROOT/baseline/_ref/lib/$f" "$OPS/lib/"
done
for f in oidnBenchmark oidnTest oidnDenoise; do
  cp -al "$ROOT/baseline/_ref/bin/$f" "$OPS/bin/" 2>/dev/null || cp -a "$ROOT/baseline/_ref/bin/$f" "$OPS/bin/"
done
cp "$ROOT/oidn_b200/liboidn_b200.so" "$OPS/lib/"
echo "op-level module (unmodified reference core) installed in $OPS"
