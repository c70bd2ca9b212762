// TEST INFRASTRUCTURE ONLY (tests/emu): a stand-in for <cuda_runtime.h> that lets g++ compile a .cu translation unit of
// the product and EXECUTE its kernels on host threads (tests/emu/emu.h).  Never part of the product or of any timed path.
#pragma once
#include "emu.h"
