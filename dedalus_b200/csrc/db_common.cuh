// Common definitions for the dedalus_b200 CUDA kernels (sm_100a).
#pragma once
#ifdef DB_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#define DB_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define DB_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char db_smem_raw[]; \
    type* name = reinterpret_cast<type*>(db_smem_raw)
#endif

#ifdef DB_EMU
#define DB_LDCS(p) (*(p))
#else
#define DB_LDCS(p) __ldcs(p)           // streaming (evict-first) load: one-pass data must not displace reused lines
#endif

#include "../../include/dedalus_b200.h"

// error plumbing (thread-compatible: one stream per rank, last error per process)
void db_set_error(const char* fmt, ...);
int db_check_launch(const char* what);

#define DB_MAX_SMEM (227 * 1024)
