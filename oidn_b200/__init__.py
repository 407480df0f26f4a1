"""oidn_b200 — a B200-native (sm_100a) backend for Open Image Denoise's UNet denoising path.

The product is the C-ABI shared library `liboidn_b200.so` (include/oidn_b200.h,
include/oidn_b200_kernels.h): hand-written CUDA kernels plus a C++ host layer that mirrors the
reference's Engine/Op plugin surface and its RT / RTLightmap filters. This package is the thin
Python host used by the tests and the benchmark: `oidn_b200.api` wraps the filter-level C ABI with
the vocabulary of the reference's C++ wrapper (include/OpenImageDenoise/oidn.hpp: DeviceRef,
FilterRef, BufferRef), `oidn_b200.weights` writes/reads TZA weight blobs and `oidn_b200.synth`
generates oidnBenchmark-style inputs. There is no CPU fallback: importing `oidn_b200.api` without a
built library raises.
"""
from .weights import model_tza, make_weights, write_tza, read_tza, unet_layers, flops_per_pixel  # noqa: F401

__version__ = "0.1.0"
