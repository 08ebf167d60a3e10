#pragma once
#include <taichi/stub_more.h>
