// TEST INFRASTRUCTURE ONLY (tests/emu): the few NCCL type names fg_api.cu mentions, so that the emulated build
// does not need the CUDA headers the real <nccl.h> pulls in. The emulated library never creates a communicator
// (multi-rank CPU tests exchange through gloo in host memory); the symbols are resolved with dlopen at run time.
#pragma once
#include <stddef.h>
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3 } ncclResult_t;
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat32 = 7 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
