"""Probe: does cuTensorMapEncodeTiled accept OVERLAPPING strides (stride[1] < dim[0]*elemsize)?"""
import torch
from cuda.bindings import driver as drv
x = torch.zeros(1 << 20, dtype=torch.bfloat16, device="cuda")
def tryit(dims, strides, box, es):
    r = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, len(dims), x.data_ptr(),
        [drv.cuuint64_t(d) for d in dims], [drv.cuuint64_t(s) for s in strides], [drv.cuuint32_t(b) for b in box],
        [drv.cuuint32_t(e) for e in es], drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE,
        drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    print(dims, strides, box, "->", r[0])
# inner 64 elements, next dim stride 32 bytes (16 elements): overlapping windows
tryit([64, 112, 224, 4], [32, 232 * 16, 232 * 16 * 224], [64, 112, 1, 1], [1, 1, 1, 1])
tryit([64, 112, 224, 4], [32, 232 * 16, 232 * 16 * 224], [64, 112, 2, 1], [1, 1, 2, 1])
tryit([64, 112, 224, 4], [16, 232 * 16, 232 * 16 * 224], [64, 112, 1, 1], [1, 1, 1, 1])
