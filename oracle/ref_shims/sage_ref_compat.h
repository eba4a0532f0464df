// Force-included (-include) in front of the UNMODIFIED reference sources when
// oracle/build_ref.py compiles them for the parity oracle.  The reference was written
// against libtorch 1.9, where AT_DISPATCH_FLOATING_TYPES accepted `tensor.type()`
// (a DeprecatedTypeProperties); torch 2.x removed that overload of ::detail::scalar_type.
// Re-adding it here lets the 13 `x.type()` dispatch sites compile without touching the
// reference files.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail
{
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
} // namespace detail
