// resolves <cuda_runtime.h> for the emulator build (tests/simt): everything lives in simt.h
#pragma once
#include "../simt.h"
