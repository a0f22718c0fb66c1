#!/bin/bash
# Builds bindings/_C*.so: the pybind11 module with the reference's three native entry points (vision.cpp:11-15) on top of
# libmrcnn_b200.so.  Host compiler only (the CUDA code is in the library); the .so is git-ignored and ships to the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
LIBDIR="$HERE/../maskrcnn_b200"
PY="${PYTHON:-python}"
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
read -r TORCH_INC1 TORCH_INC2 TORCH_LIB PY_INC EXT ABI <<<"$($PY - <<'PY'
import sysconfig, torch, os
from torch.utils.cpp_extension import include_paths, library_paths
inc = [p for p in include_paths() if 'cuda' not in p.split(os.sep)[-2:]]
print(inc[0], inc[1], library_paths()[0], sysconfig.get_paths()['include'],
      sysconfig.get_config_var('EXT_SUFFIX'), int(torch._C._GLIBCXX_USE_CXX11_ABI))
PY
)"
TARGET="$HERE/_C$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$HERE/vision_b200.cpp" ] && [ "$TARGET" -nt "$HERE/../include/mrcnn_b200.h" ]; then
  echo "bindings/_C up to date"; exit 0
fi
g++ -O2 -std=c++17 -fPIC -shared -w \
    -DTORCH_EXTENSION_NAME=_C -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI \
    -I"$TORCH_INC1" -I"$TORCH_INC2" -I"$PY_INC" -I"$CUDA_INC" \
    "$HERE/vision_b200.cpp" \
    -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python \
    -L"$LIBDIR" -lmrcnn_b200 -Wl,-rpath,"$TORCH_LIB" -Wl,-rpath,'$ORIGIN/../maskrcnn_b200' \
    -o "$TARGET"
echo "built $TARGET"
