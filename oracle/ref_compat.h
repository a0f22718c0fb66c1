/* TEST INFRASTRUCTURE ONLY — pre-included (g++ -include) when the UNMODIFIED reference
 * sources under /root/reference/c++ext/maskrcnn/csrc are compiled into oracle/_ref/.
 *
 * The reference was written for PyTorch 1.0.  Exactly one construct no longer compiles on
 * torch 2.x: cpu/nms_cpu.cpp:75 hands a DeprecatedTypeProperties (`dets.type()`) to
 * AT_DISPATCH_FLOATING_TYPES, which now wants a c10::ScalarType.  Instead of patching the
 * source we re-define the dispatch macro so that it accepts either.  No arithmetic is touched.
 */
#pragma once
#include <torch/extension.h>

namespace ref_compat {
inline c10::ScalarType to_scalar_type(c10::ScalarType t) { return t; }
inline c10::ScalarType to_scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)                                   \
    do {                                                                              \
        switch (ref_compat::to_scalar_type(TYPE)) {                                   \
            case c10::ScalarType::Float: { using scalar_t = float; __VA_ARGS__(); break; }   \
            case c10::ScalarType::Double: { using scalar_t = double; __VA_ARGS__(); break; } \
            default: AT_ERROR(NAME, " not implemented for this dtype");               \
        }                                                                             \
    } while (0)
