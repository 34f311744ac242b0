// internal helpers of the C++ mirror
#pragma once
#include <tfusion_b200.h>

namespace tfusion {
namespace detail {
// a tiny context (8x8 image, 1-block scene) that owns the stream used by stand-alone memory / image calls
tfb_ctx* util_ctx();
void check(int rc, const char* what, const char* file, int line);
}  // namespace detail
}  // namespace tfusion
#define TF_CHECK(call) ::tfusion::detail::check((call), #call, __FILE__, __LINE__)
