// Platform layer: the product is compiled by nvcc for sm_100a.  The only other configuration, ZK_EMU, compiles
// the same kernel sources with g++ against tests/emu/cuda_emu.hpp so that kernel logic can be exercised on host
// threads by the CPU-only test suite.  ZK_EMU is never defined for libzkcnn_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#ifdef ZK_EMU
#include "../../tests/emu/cuda_emu.hpp"
#define ZK_HD
#define ZK_DEVCONST static const
#define ZK_ON_DEVICE 0
#define ZK_NOINLINE __attribute__((noinline))
// kernel launch: run the body on the emulator
#define ZK_LAUNCH(kernel, grid, block, smem, stream, ...) \
    zkemu::launch_k((grid), (block), (smem), kernel, __VA_ARGS__)
#define ZK_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(zkemu::dyn_smem())
typedef void *zk_stream_t;
#elif defined(ZK_HOST_ONLY)
// host-side users of the field / curve headers (zkcnn_b200/host): no CUDA dependency, portable arithmetic only
#define ZK_HD
#define ZK_DEVCONST static const
#define ZK_ON_DEVICE 0
#define ZK_NOINLINE __attribute__((noinline))
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#else
#include <cuda_runtime.h>
#define ZK_HD __host__ __device__
#define ZK_DEVCONST static __device__ __constant__ const
#ifdef __CUDA_ARCH__
#define ZK_ON_DEVICE 1
#else
#define ZK_ON_DEVICE 0
#endif
#define ZK_NOINLINE __noinline__
#define ZK_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define ZK_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char zk_dyn_smem_[]; \
    type *name = reinterpret_cast<type *>(zk_dyn_smem_)
typedef cudaStream_t zk_stream_t;
#endif

// A constant array visible from both host and device code (device copy lives in constant memory).
#define ZK_CONST_ARRAY(name, n, ...)                  \
    static const uint32_t h_##name[n] = {__VA_ARGS__}; \
    ZK_DEVCONST uint32_t d_##name[n] = {__VA_ARGS__}
#if ZK_ON_DEVICE
#define ZK_C(name) d_##name
#else
#define ZK_C(name) h_##name
#endif

#define ZK_SM_COUNT 148  // B200: 148 SMs; grids are sized in multiples of this
