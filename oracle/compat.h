// Forced-include shim used ONLY when compiling the unmodified reference C++
// sources (under /root/reference/vinum_cpp/src) against the Arrow 24 headers
// that ship with this image's pyarrow.  The reference was written for Arrow 3.0
// where these two names still existed (agg_funcs.h:32, array_iterators.h:28).
// Test infrastructure only -- never part of the product path.
#pragma once
#include <string_view>
#include <arrow/util/bit_util.h>
namespace arrow {
namespace util { using string_view = std::string_view; }
namespace BitUtil = ::arrow::bit_util;
}  // namespace arrow
