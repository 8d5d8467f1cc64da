// dlpack_abi.h -- the DLPack v0.8 C structs (unversioned DLManagedTensor), restated
// from the published ABI so the library has no third-party include.  These are the
// structs torch.from_dlpack consumes from a PyCapsule named "dltensor".
#pragma once
#include <stdint.h>
extern "C" {
typedef enum { kXrDLCPU = 1, kXrDLCUDA = 2 } XrDLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } XrDLDevice;
typedef enum { kXrDLInt = 0, kXrDLUInt = 1, kXrDLFloat = 2 } XrDLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } XrDLDataType;
typedef struct {
    void *data; XrDLDevice device; int32_t ndim; XrDLDataType dtype;
    int64_t *shape; int64_t *strides; uint64_t byte_offset;
} XrDLTensor;
typedef struct XrDLManagedTensor {
    XrDLTensor dl_tensor; void *manager_ctx; void (*deleter)(struct XrDLManagedTensor *self);
} XrDLManagedTensor;
}
