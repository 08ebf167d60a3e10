#pragma once
#include "../simt_cub.h"
