// MOCK of the subset of xla/ffi/api/ffi.h that stac_mjx_b200/csrc/stacb_xla_ffi.cc uses -- TEST INFRASTRUCTURE ONLY.
// jaxlib (and with it the real header) is not installed in the authoring image, so the handler translation unit cannot be built
// there.  This mock has the same SHAPE as the typed FFI API (Buffer / ResultBuffer / Error / Ffi::Bind().Ctx().Attr().Arg().Ret() /
// XLA_FFI_DEFINE_HANDLER_SYMBOL) and type-checks, at compile time, that every handler is invocable with exactly the argument list its
// binding declares; compiling the TU against it also checks every call of the C ABI against include/stacb.h.  It proves nothing
// about run-time behaviour inside XLA.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32, S32, U8 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

struct Dims {
  std::vector<int64_t> d;
  size_t size() const { return d.size(); }
  int64_t operator[](size_t i) const { return d[i]; }
};

template <DataType dt>
struct Buffer {
  using T = typename NativeOf<dt>::type;
  T *ptr = nullptr;
  Dims dims;
  T *typed_data() const { return ptr; }
  const Dims &dimensions() const { return dims; }
  size_t size_bytes() const { return 0; }
};

template <typename T>
struct Result {
  T value;
  T *operator->() { return &value; }
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kInvalidArgument, kInternal };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};

// binding builder: accumulates the C++ argument types the handler must accept
template <typename... Ts>
struct Binding {
  template <typename T> struct CtxArg { using type = T; };
  template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
  template <typename C> Binding<Ts..., typename CtxArg<C>::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char *) const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename Fn> static constexpr bool accepts() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                          \
  static_assert(decltype(binding)::template accepts<decltype(&impl)>(), #impl " does not match its FFI binding");    \
  extern "C" void *symbol() { return reinterpret_cast<void *>(&impl); }
