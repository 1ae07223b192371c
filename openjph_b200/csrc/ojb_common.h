// ojb_common.h -- shared host-side types for the B200 HTJ2K path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <stdexcept>
#include <vector>

namespace ojb {

// Error convention mirrors the reference's OJPH_ERROR (src/core/others/ojph_message.cpp:156-171):
// a 32-bit code, a formatted message, and a C++ exception; the C-ABI layer (ojb_capi.cpp)
// catches it and returns a negative status + ojb_last_error().
struct Error : public std::runtime_error {
  uint32_t code;
  Error(uint32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(uint32_t code, const char* fmt, ...) {
  char buf[768];
  va_list ap; va_start(ap, fmt);
  int n = snprintf(buf, sizeof(buf), "ojph error 0x%08X: ", code);
  vsnprintf(buf + n, sizeof(buf) - (size_t)n, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

struct Rect {
  uint32_t x0 = 0, y0 = 0, w = 0, h = 0;
  uint32_t x1() const { return x0 + w; }
  uint32_t y1() const { return y0 + h; }
  bool empty() const { return w == 0 || h == 0; }
};

inline uint32_t div_ceil(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline uint32_t ilog2(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }

} // namespace ojb
