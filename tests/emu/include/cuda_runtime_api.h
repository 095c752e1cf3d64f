// stand-in for <cuda_runtime_api.h> in emulator builds (TEST INFRASTRUCTURE ONLY, see cuda_runtime.h)
#pragma once
#include "cuda_runtime.h"
