#!/usr/bin/env python3
"""Probe: does cuTensorMapEncodeTiled accept a dimension whose stride is smaller than the extent of the previous one
(overlapping sliding windows)?  GPU box only."""
import torch
from cuda.bindings import driver as drv

torch.zeros(1).cuda()
buf = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
DT = drv.CUtensorMapDataType


def enc(dims, strides, box, estr, swz=drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, dt=DT.CU_TENSOR_MAP_DATA_TYPE_TFLOAT32):
    r = drv.cuTensorMapEncodeTiled(dt, len(dims), buf.data_ptr(), [drv.cuuint64_t(d) for d in dims],
                                   [drv.cuuint64_t(s) for s in strides], [drv.cuuint32_t(b) for b in box],
                                   [drv.cuuint32_t(e) for e in estr], drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                   drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    return r[0]


Wp, H, N, Q = 232, 224, 256, 112
print("overlap s2 :", enc([32, Q, H, N], [32, Wp * 16, H * Wp * 16], [32, 112, 1, 1], [1, 1, 1, 1]))
print("overlap s1 :", enc([32, 32, 32, 64], [16, 40 * 16, 32 * 40 * 16], [32, 32, 4, 1], [1, 1, 1, 1]))
print("overlap s1 box128:", enc([32, 224, 224, 8], [16, 232 * 16, 224 * 232 * 16], [32, 128, 1, 1], [1, 1, 1, 1]))
print("estride h  :", enc([32, Q, H, N], [32, Wp * 16, H * Wp * 16], [32, 56, 4, 1], [1, 1, 2, 1]))
print("atom32     :", enc([32, Q, H, N], [32, Wp * 16, H * Wp * 16], [32, 32, 1, 1], [1, 1, 1, 1],
                          swz=drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
