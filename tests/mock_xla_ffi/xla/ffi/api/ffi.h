// COMPILE-ONLY STAND-IN for jaxlib's "xla/ffi/api/ffi.h" (tests/mock_xla_ffi). jaxlib is not installable in the build image, so
// ffi/magpo_ffi.cc cannot be compiled against the real header here; this file declares just enough of the public XLA FFI C++ API
// surface that the shim uses (names and call shapes as documented for jax.ffi / XLA FFI, written from memory) for a
// `g++ -fsyntax-only` pass that checks the shim's own code: argument counts and types of every magpo_* call, struct filling, error
// paths. It proves nothing about the real header. Never shipped, never linked.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <utility>

namespace xla::ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T>
class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
 private:
  const T* p_;
  size_t n_;
};

class AnyBuffer {
 public:
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_, rank_); }
  size_t element_count() const { size_t n = 1; for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i]; return n; }
  size_t size_bytes() const { return bytes_; }
 private:
  void* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0, bytes_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
 private:
  T v_;
};

template <typename T>
class ErrorOr {
 public:
  bool has_value() const { return ok_; }
  T& value() { return v_; }
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
  Error error() const { return Error::InvalidArgument("missing operand"); }
 private:
  bool ok_ = true;
  T v_;
};

class RemainingArgs {
 public:
  size_t size() const { return n_; }
  template <typename T> ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
 private:
  size_t n_ = 0;
};
class RemainingRets {
 public:
  size_t size() const { return n_; }
  template <typename T> ErrorOr<Result<T>> get(size_t) const { return ErrorOr<Result<T>>(); }
 private:
  size_t n_ = 0;
};

template <typename T> struct PlatformStream {};

class Binding {  // the real builder tracks the handler's parameter types; the stand-in only accepts the same call chain
 public:
  template <typename T> Binding& Ctx() { return *this; }
  Binding& RemainingArgs() { return *this; }
  Binding& RemainingRets() { return *this; }
  template <typename T> Binding& Attr(std::string_view) { return *this; }
};
struct Ffi {
  static Binding Bind() { return {}; }
};

}  // namespace xla::ffi

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
// the real macro expands to `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)` dispatching through the binding to `impl`
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                 \
  static auto name##_binding_ = (binding);                                 \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { (void)&impl; return nullptr; }
