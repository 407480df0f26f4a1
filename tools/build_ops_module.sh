#!/bin/bash
# Rebuilds only the op-level device module (integration/b200_module.cpp -DOIDN_B200_OP_LEVEL) against the unmodified
# reference build of tools/build_reference_cuda.sh (/tmp/oidn_build) and refreshes baseline/_b200_ops.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OPS="$ROOT/baseline/_b200_ops"; REFBLD=/tmp/oidn_build
[ -f $REFBLD/libOpenImageDenoise_core.so.2.4.1 ] || bash "$ROOT/tools/build_reference_cuda.sh"
make -s -C "$ROOT/oidn_b200/csrc"
mkdir -p "$OPS/lib" "$OPS/bin"
${CXX:-g++} -std=c++17 -O2 -fPIC -fvisibility=hidden -fvisibility-inlines-hidden -Wall -Wno-unknown-pragmas \
     -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS -DOIDN_B200_OP_LEVEL \
     -I/usr/local/cuda/targets/x86_64-linux/include -I"$ROOT/include" -isystem /tmp/oidn_ref -isystem /tmp/oidn_ref/external -isystem $REFBLD \
     -shared -Wl,-soname,libOpenImageDenoise_device_cuda.so.2.4.1 -Wl,-z,now \
     -o "$OPS/lib/libOpenImageDenoise_device_cuda.so.2.4.1" "$ROOT/integration/b200_module.cpp" \
     $REFBLD/libOpenImageDenoise_core.so.2.4.1 -L"$ROOT/oidn_b200" -loidn_b200 \
     -L/usr/local/cuda/targets/x86_64-linux/lib -lcudart_static -lrt -lpthread -ldl \
     -Wl,-rpath,'$ORIGIN'
ln -sf libOpenImageDenoise_device_cuda.so.2.4.1 "$OPS/lib/libOpenImageDenoise_device_cuda.so"
cp "$ROOT/oidn_b200/liboidn_b200.so" "$OPS/lib/"
echo "op-level module rebuilt in $OPS"
