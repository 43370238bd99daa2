// oracle/ref_shim.h -- TEST INFRASTRUCTURE.  Force-included (-include) when oracle/ref_cuda_build.py compiles the
// reference's own CUDA op from /root/reference/src/models/ops/src WITHOUT touching its sources.
//
// The reference was written against torch 1.11 and dispatches with `AT_DISPATCH_FLOATING_TYPES(value.type(), ...)`
// (cuda/ms_deform_attn_cuda.cu:64,134).  torch 2.11's dispatch macro calls ::detail::scalar_type(TYPE), which only
// has an overload for at::ScalarType; this header adds the one for the deprecated `.type()` object.  Nothing else.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
