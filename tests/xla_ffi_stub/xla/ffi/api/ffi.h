// MINIMAL LOCAL STAND-IN for the binding types of XLA's "xla/ffi/api/ffi.h" -- TEST INFRASTRUCTURE, NOT THE REAL HEADER.
//
// jaxlib (which ships the real header under jax.ffi.include_dir()) is not installed in this image and cannot be, so
// fbs_b200/csrc/xla_ffi_shim.cc could otherwise never meet a compiler.  This file declares just the public names the shim
// uses -- xla::ffi::{DataType constants, Span, Buffer, Result / ResultBuffer, Error, ErrorCode, ScratchAllocator,
// PlatformStream, Ffi::Bind() with Ctx / Arg / Attr / Ret} and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the shapes they have in
// the public XLA FFI API (written from the documented interface, no XLA source is copied), so that
// tests/test_xla_ffi_shim.py can (1) parse and type-check the shim every round and (2) statically verify that every
// handler's implementation is invocable with exactly the argument list its Bind() chain decodes, which is the check the real
// header performs.  The handlers defined through this stand-in return an error when called: it is a compile-only check.
#ifndef FBS_TESTS_XLA_FFI_STUB_H_
#define FBS_TESTS_XLA_FFI_STUB_H_

#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>

extern "C" {
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_CallFrame XLA_FFI_CallFrame;
}

namespace xla {
namespace ffi {

enum class DataType : uint8_t { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType PRED = DataType::PRED;
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType U8 = DataType::U8;
inline constexpr DataType U32 = DataType::U32;
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType BF16 = DataType::BF16;

namespace internal {
template <DataType> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<DataType::PRED> { using type = bool; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeTypeOf<DataType::U32> { using type = uint32_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::BF16> { using type = uint16_t; };
}  // namespace internal

template <typename T>
class Span {
 public:
  constexpr Span(T* data, size_t size) : data_(data), size_(size) {}
  constexpr size_t size() const { return size_; }
  constexpr T& operator[](size_t i) const { return data_[i]; }
  constexpr T& front() const { return data_[0]; }
  constexpr T& back() const { return data_[size_ - 1]; }
  constexpr T* begin() const { return data_; }
  constexpr T* end() const { return data_ + size_; }

 private:
  T* data_;
  size_t size_;
};

template <DataType dtype>
class Buffer {
 public:
  using NativeType = typename internal::NativeTypeOf<dtype>::type;
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_, rank_); }
  NativeType* typed_data() const { return static_cast<NativeType*>(data_); }
  void* untyped_data() const { return data_; }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= static_cast<size_t>(dims_[i]);
    return n;
  }

 private:
  void* data_ = nullptr;
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
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode : uint8_t { kOk, kCancelled, kUnknown, kInvalidArgument, kNotFound, kUnimplemented, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t /*size*/, size_t /*alignment*/ = 1) { return std::nullopt; }
};

template <typename T>
struct PlatformStream {};

namespace internal {
// what a bound context decodes to in the implementation's argument list
template <typename T> struct CtxDecoded { using type = T; };
template <typename T> struct CtxDecoded<PlatformStream<T>> { using type = T; };
template <typename... Ts> struct TypeList {};
}  // namespace internal

template <typename... Decoded>
class Binding {
 public:
  template <typename T> Binding<Decoded..., typename internal::CtxDecoded<T>::type> Ctx() && { return {}; }
  template <typename T> Binding<Decoded..., T> Arg() && { return {}; }
  template <typename T> Binding<Decoded..., T> Attr(const char* /*name*/) && { return {}; }
  template <typename T> Binding<Decoded..., Result<T>> Ret() && { return {}; }
  template <typename Fn>
  static constexpr bool Invocable() { return std::is_invocable_r_v<Error, Fn, Decoded...>; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// The real macro defines `extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*)` that decodes the call frame and invokes
// `impl`.  The stand-in keeps the symbol and turns the decode step into a static signature check.
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                            \
  static_assert(decltype(binding)::template Invocable<decltype(&impl)>(),                                               \
                #impl " is not invocable with the arguments its Bind() chain decodes");                                 \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame* /*call_frame*/) {                                                 \
    return reinterpret_cast<XLA_FFI_Error*>(sizeof(&impl)); /* stand-in: never a valid success value */                 \
  }

#endif  // FBS_TESTS_XLA_FFI_STUB_H_
