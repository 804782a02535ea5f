// Stand-in for <cuda_runtime.h> when a kernel source is compiled by g++ for the warp emulator (tests only):
// everything the kernels need comes from warp_emu.hpp.
#pragma once
#include "../warp_emu.hpp"
