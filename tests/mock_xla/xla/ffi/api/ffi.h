// TEST INFRASTRUCTURE ONLY -- a minimal, inert stand-in for jaxlib's
// `xla/ffi/api/ffi.h` (jax / jaxlib are not installable in this image).
//
// It models just enough of the typed-FFI binding surface for
// swirl_fem_b200/csrc/xla_ffi_shim.cc to be COMPILED and TYPE-CHECKED on the
// CPU (tests/test_xla_ffi_shim.py): `Ffi::Bind().Ctx<>().Arg<>().Ret<>()
// .Attr<>()` records the decoded C++ type of every bound operand, and
// `XLA_FFI_DEFINE_HANDLER_SYMBOL` static-asserts that the handler function is
// invocable with exactly those types and returns `ffi::Error` -- the contract
// the real header enforces through its own template machinery.  Nothing here
// executes a handler; the exported symbol returns nullptr.
#ifndef TESTS_MOCK_XLA_FFI_API_FFI_H_
#define TESTS_MOCK_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

extern "C" {
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_CallFrame XLA_FFI_CallFrame;
}

namespace xla {
namespace ffi {

enum class DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64,
                      BF16 };
enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal,
                       kUnimplemented };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message)
      : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* data, size_t size) : data_(data), size_(size) {}
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }

 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

class AnyBuffer {
 public:
  using Dimensions = Span<int64_t>;
  DataType element_type() const { return dtype_; }
  void* untyped_data() const { return data_; }
  Dimensions dimensions() const { return Dimensions(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
    return n;
  }
  size_t size_bytes() const { return element_count() * 8; }

 private:
  DataType dtype_ = DataType::F64;
  void* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <DataType dtype> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::F64> { using type = double; };

template <DataType dtype>
class Buffer {
 public:
  using Native = typename NativeTypeOf<dtype>::type;
  Native* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i];
    return n;
  }

 private:
  Native* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};

template <typename T>
struct PlatformStream {};

namespace internal {
// what a bound operand decodes to in the handler's parameter list
template <typename T> struct CtxDecoded { using type = T; };
template <typename T> struct CtxDecoded<PlatformStream<T>> { using type = T; };
}  // namespace internal

template <typename Fn, typename... Ts>
struct Handler {
  static_assert(std::is_invocable_r_v<Error, Fn, Ts...>,
                "handler signature does not match the Ffi::Bind() operand list");
  Fn fn;
};

template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., typename internal::CtxDecoded<T>::type> Ctx() && { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() && { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() && { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(std::string) && {
    static_assert(std::is_arithmetic_v<T>, "mock: scalar attributes only");
    return {};
  }
  template <typename Fn>
  Handler<Fn, Ts...> To(Fn fn) && { return Handler<Fn, Ts...>{fn}; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// The real macro instantiates the handler and exports a C symbol taking the
// call frame; so does this one (the static_assert above fires at this point).
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                    \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame* frame) {                  \
    static auto handler = (binding).To(impl);                                 \
    (void)handler;                                                            \
    (void)frame;                                                              \
    return nullptr;                                                           \
  }

#endif  // TESTS_MOCK_XLA_FFI_API_FFI_H_
