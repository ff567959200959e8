// STUB of rmagine/types/Bundle.hpp (tests/stubs/README.md): a Bundle inherits every attribute struct it is given
#pragma once
#include <type_traits>

namespace rmagine {
template <typename... Tp> struct Bundle : public Tp... {
    template <typename T> static constexpr bool has() { return (std::is_same_v<T, Tp> || ...); }
};
}  // namespace rmagine
