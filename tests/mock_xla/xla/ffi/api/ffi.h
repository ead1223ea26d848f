// MOCK of the subset of jaxlib's xla/ffi/api/ffi.h that merlot_reserve_b200/csrc/xla_ffi_shim.cc uses, restated from
// memory of the published API (jax >= 0.4.31).  Test infrastructure only: it lets tests/test_abi.py syntax-check the
// shim with g++ in an image that has no jaxlib.  It proves the shim is well-formed C++ against THIS restatement, not
// against the real header.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <optional>
#include <string>

namespace xla::ffi {
enum class ErrorCode { kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};
template <class T>
class ErrorOr {
 public:
  bool has_value() const { return v_.has_value(); }
  Error error() const { return Error(); }
  T* operator->() { return &*v_; }
  T& operator*() { return *v_; }
 private:
  std::optional<T> v_;
};
template <class T>
class Span {
 public:
  const T* begin() const { return p_; }
  size_t size() const { return n_; }
 private:
  const T* p_ = nullptr;
  size_t n_ = 0;
};
class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};
template <class T>
class Result {
 public:
  T* operator->() { return &t_; }
 private:
  T t_;
};
class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <class T> ErrorOr<T> get(size_t) const { return {}; }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <class T> ErrorOr<Result<T>> get(size_t) const { return {}; }
};
template <class T> struct PlatformStream {};
struct Binding {
  template <class T> Binding& Ctx() { return *this; }
  Binding& RemainingArgs() { return *this; }
  Binding& RemainingRets() { return *this; }
  template <class T> Binding& Attr(const char*) { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}  // namespace xla::ffi
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding) \
  extern "C" void* sym() { (void)(binding); return reinterpret_cast<void*>(&impl); }
