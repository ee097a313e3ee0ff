// Minimal restatement of the DLPack in-memory tensor ABI (dmlc/dlpack, "DLManagedTensor" flavour) --
// only the structs libscfeat needs to hand a device buffer to tf.experimental.dlpack.from_dlpack /
// torch.from_dlpack.  Field order and widths follow the published specification.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;

typedef struct {
    DLDeviceType device_type;
    int32_t device_id;
} DLDevice;

typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U } DLDataTypeCode;

typedef struct {
    uint8_t code;
    uint8_t bits;
    uint16_t lanes;
} DLDataType;

typedef struct {
    void* data;
    DLDevice device;
    int32_t ndim;
    DLDataType dtype;
    int64_t* shape;
    int64_t* strides;
    uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void* manager_ctx;
    void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
