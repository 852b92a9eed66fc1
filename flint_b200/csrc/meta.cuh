/* meta.cuh -- the handful of type traits the kernel headers use, written out so that the same headers compile under nvcc and
 * under NVRTC (run-time registration of user systems, rtc.cu), which has no standard library headers. */
#ifndef FLINT_B200_META_CUH
#define FLINT_B200_META_CUH

namespace flint {
namespace meta {

template <class T, T V> struct integral_constant { static constexpr T value = V; };
template <bool B> using bool_constant = integral_constant<bool, B>;
using true_type = bool_constant<true>;
using false_type = bool_constant<false>;
template <class...> using void_t = void;
template <class T> struct remove_reference { using type = T; };
template <class T> struct remove_reference<T&> { using type = T; };
template <class T> struct remove_reference<T&&> { using type = T; };
template <class T> struct remove_cv { using type = T; };
template <class T> struct remove_cv<const T> { using type = T; };
template <class T> struct remove_cv<volatile T> { using type = T; };
template <class T> struct remove_cv<const volatile T> { using type = T; };
template <class T> using remove_cvref_t = typename remove_cv<typename remove_reference<T>::type>::type;
template <class T> T&& declval() noexcept;       /* unevaluated contexts only */

}  // namespace meta
}  // namespace flint
#endif
