// stand-in for the CPU mock build: NVTX ranges are no-ops
#pragma once
inline int nvtxRangePushA(const char *) { return 0; }
inline int nvtxRangePop() { return 0; }
