// MINIMAL STUB of the XLA FFI C++ API (xla/ffi/api/ffi.h), written for tests/test_abi.py only.
// jaxlib is not installable in this image, so forde-vision-language_b200/csrc/forde_xla_ffi.cc cannot meet the real
// header here.  This stub declares just the names that file uses -- Buffer / ResultBuffer / Error / PlatformStream /
// Ffi::Bind()...  / XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the shapes the public API documents, so that a syntax-only
// compile checks (a) the shim parses, includes what it uses and calls the C ABI with the header's current
// signatures, and (b) every handler's parameter list matches its binding (Binding::To static_asserts invocability).
// It is NOT a substitute for compiling against jaxlib; nothing links against it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { S8, S32, F32, F64, BF16 };
enum class ErrorCode { kOk, kInternal, kInvalidArgument };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <DataType dt> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<S8> { using type = int8_t; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<BF16> { using type = uint16_t; };

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t len = 0;
  size_t size() const { return len; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType dt>
class Buffer {
 public:
  using Native = typename NativeTypeOf<dt>::type;
  void* untyped_data() const { return data_; }
  Native* typed_data() const { return static_cast<Native*>(data_); }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return 0; }

 private:
  void* data_ = nullptr;
  Span<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename T> struct PlatformStream {};

namespace stub_detail {
template <typename T> struct CtxTag {};
template <typename T> struct Decoded { using type = T; };
template <typename T> struct Decoded<CtxTag<PlatformStream<T>>> { using type = T; };
}  // namespace stub_detail

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., stub_detail::CtxTag<T>> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename Fn>
  int To(Fn) const {
    static_assert(std::is_invocable_r<Error, Fn, typename stub_detail::Decoded<Ts>::type...>::value,
                  "handler signature does not match its binding");
    return 0;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  static const int name##_binding_check = (binding).To(impl); \
  extern "C" void* name(void*) { return nullptr; }
