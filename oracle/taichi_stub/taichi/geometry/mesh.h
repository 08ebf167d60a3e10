#pragma once
#include <taichi/util.h>
