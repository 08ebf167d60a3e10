#pragma once
namespace taichi {
namespace bit {
constexpr bool is_power_of_two(int x) { return x > 0 && (x & (x - 1)) == 0; }
}  // namespace bit
}  // namespace taichi
