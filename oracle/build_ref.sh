#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own CPU extension (vision.cpp + cpu/nms_cpu.cpp +
# cpu/crop_cpu.cpp) from where the sources lie under /root/reference into oracle/_ref/ref_C*.so.
# Sources are NOT copied; only the built .so lands in the (git-ignored) oracle/_ref/.
# -O2 is what the reference's own `setup.py install` would inherit from Python's CFLAGS.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/c++ext/maskrcnn/csrc"
[ -d "$REF" ] || { echo "reference sources not present at $REF; keeping prebuilt oracle/_ref (if any)"; exit 0; }
PY="${PYTHON:-python}"
OUT="$HERE/_ref"; mkdir -p "$OUT"
read -r TORCH_INC1 TORCH_INC2 TORCH_LIB PY_INC EXT ABI <<<"$($PY - <<'PY'
import sysconfig, torch, os
from torch.utils.cpp_extension import include_paths, library_paths
inc = [p for p in include_paths() if 'cuda' not in p.split(os.sep)[-2:]]
print(inc[0], inc[1], library_paths()[0], sysconfig.get_paths()['include'],
      sysconfig.get_config_var('EXT_SUFFIX'), int(torch._C._GLIBCXX_USE_CXX11_ABI))
PY
)"
TARGET="$OUT/ref_C$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/cpu/crop_cpu.cpp" ] && [ "$TARGET" -nt "$HERE/ref_compat.h" ]; then
  echo "oracle/_ref up to date"; exit 0
fi
g++ -O2 -std=c++17 -fPIC -shared -w \
    -DTORCH_EXTENSION_NAME=ref_C -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI \
    -include "$HERE/ref_compat.h" \
    -I"$REF" -I"$TORCH_INC1" -I"$TORCH_INC2" -I"$PY_INC" \
    "$REF/vision.cpp" "$REF/cpu/nms_cpu.cpp" "$REF/cpu/crop_cpu.cpp" \
    -L"$TORCH_LIB" -ltorch -ltorch_cpu -lc10 -ltorch_python -Wl,-rpath,"$TORCH_LIB" \
    -o "$TARGET"
echo "built $TARGET"
