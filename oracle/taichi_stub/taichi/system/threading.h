#pragma once
#include <taichi/util.h>
namespace taichi {
struct Spinlock {  // one byte, as GridState's power-of-two size assert requires (src/mpm_fwd.h:113-117)
  uint8 flag = 0;
  void lock() { while (__atomic_test_and_set(&flag, __ATOMIC_ACQUIRE)) {} }   // src/rigid_transfer.cpp:62,73
  void unlock() { __atomic_clear(&flag, __ATOMIC_RELEASE); }
};
}  // namespace taichi
