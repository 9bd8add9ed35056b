// TEST INFRASTRUCTURE shim: minimal absl::Span so the reference's scoring
// sources compile unmodified without abseil (not available offline).
#ifndef SHIM_ABSL_TYPES_SPAN_H_
#define SHIM_ABSL_TYPES_SPAN_H_
#include <array>
#include <cstddef>
#include <type_traits>
#include <vector>
namespace absl {
template <typename T>
class Span {
 public:
  using value_type = std::remove_cv_t<T>;
  using const_iterator = T const*;
  using iterator = T*;
  static constexpr std::size_t npos = static_cast<std::size_t>(-1);
  constexpr Span() noexcept = default;
  constexpr Span(T* ptr, std::size_t len) noexcept : mPtr(ptr), mLen(len) {}
  template <typename A>
  Span(std::vector<value_type, A> const& v) noexcept : mPtr(v.data()), mLen(v.size()) {}  // NOLINT
  template <typename A, typename U = T, typename = std::enable_if_t<!std::is_const_v<U>>>
  Span(std::vector<value_type, A>& v) noexcept : mPtr(v.data()), mLen(v.size()) {}  // NOLINT
  template <std::size_t N>
  constexpr Span(std::array<value_type, N> const& a) noexcept : mPtr(a.data()), mLen(N) {}  // NOLINT
  template <typename U, typename = std::enable_if_t<std::is_same_v<U const, T>>>
  constexpr Span(Span<U> other) noexcept : mPtr(other.data()), mLen(other.size()) {}  // NOLINT
  [[nodiscard]] constexpr auto data() const noexcept -> T* { return mPtr; }
  [[nodiscard]] constexpr auto size() const noexcept -> std::size_t { return mLen; }
  [[nodiscard]] constexpr auto length() const noexcept -> std::size_t { return mLen; }
  [[nodiscard]] constexpr auto empty() const noexcept -> bool { return mLen == 0; }
  [[nodiscard]] constexpr auto operator[](std::size_t i) const noexcept -> T& { return mPtr[i]; }
  [[nodiscard]] constexpr auto front() const noexcept -> T& { return mPtr[0]; }
  [[nodiscard]] constexpr auto back() const noexcept -> T& { return mPtr[mLen - 1]; }
  [[nodiscard]] constexpr auto begin() const noexcept -> T* { return mPtr; }
  [[nodiscard]] constexpr auto end() const noexcept -> T* { return mPtr + mLen; }
  [[nodiscard]] constexpr auto cbegin() const noexcept -> T const* { return mPtr; }
  [[nodiscard]] constexpr auto cend() const noexcept -> T const* { return mPtr + mLen; }
  [[nodiscard]] constexpr auto subspan(std::size_t pos = 0, std::size_t len = npos) const -> Span {
    if (pos > mLen) pos = mLen;
    std::size_t const rem = mLen - pos;
    return Span(mPtr + pos, len < rem ? len : rem);
  }
 private:
  T* mPtr = nullptr;
  std::size_t mLen = 0;
};
template <typename C>
auto MakeConstSpan(C const& c) noexcept -> Span<typename C::value_type const> {
  return Span<typename C::value_type const>(c.data(), c.size());
}
template <typename T>
auto MakeConstSpan(T const* p, std::size_t n) noexcept -> Span<T const> { return Span<T const>(p, n); }
template <typename C>
auto MakeSpan(C& c) noexcept -> Span<typename C::value_type> {
  return Span<typename C::value_type>(c.data(), c.size());
}
}  // namespace absl
#endif
