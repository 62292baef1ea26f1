// coop.h -- grid-wide barrier of a cooperative launch: cooperative groups on the device, the host-thread emulation of the
// test hook otherwise (cta_emu.h also supplies the CUDA vocabulary the kernels are written in).
#pragma once
#ifndef LBM_HOSTCHECK
#include <cooperative_groups.h>
#define LBM_GRID_SYNC() cooperative_groups::this_grid().sync()
#else
#include "cta_emu.h"
#endif
