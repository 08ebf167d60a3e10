#pragma once
namespace taichi {
namespace bit {
constexpr bool is_power_of_two(int x) { return x > 0 && (x & (x - 1)) == 0; }
constexpr bool is_power_of_two(unsigned long x) { return x > 0 && (x & (x - 1)) == 0; }
inline int log2int(unsigned long x) { int r = 0; while (x > 1) { x >>= 1; r++; } return r; }
}  // namespace bit
}  // namespace taichi
