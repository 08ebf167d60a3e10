#pragma once
#include <taichi/util.h>
namespace taichi {
struct Spinlock {  // one byte, as GridState's power-of-two size assert requires (src/mpm_fwd.h:113-117)
  uint8 flag = 0;
};
}  // namespace taichi
