// nccl.h -- TEST-ONLY stand-in (tests/mock/README.md): the types comm.hpp names; no communicator exists in the mock
#pragma once
#include <cstddef>
struct ncclComm;
typedef ncclComm *ncclComm_t;
typedef int ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclChar = 0, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
