// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
// Minimal single-process emulation of the CUDA execution model so that the kernels in
// dedalus_b200/csrc/*.cu can be compiled with g++ (-x c++ -DDB_EMU) and their indexing / barrier logic
// checked on the GPU-less build container.  One CUDA thread = one ucontext fiber; __syncthreads() yields to
// a round-robin block scheduler.  Blocks run one after another.  The emulated library is built as
// tests/emu/libdedalus_b200_emu.so and loaded only by tests (tests/emu/emu_lib.py).
#pragma once
#include <ucontext.h>
#include <csetjmp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__
#define __align__(x)

struct uint3e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

struct EmuState {
    uint3e tid, bid; dim3 bdim, gdim;
    unsigned char* smem;
};
extern thread_local EmuState emu_cur;           // state of the running fiber
#define threadIdx (emu_cur.tid)
#define blockIdx (emu_cur.bid)
#define blockDim (emu_cur.bdim)
#define gridDim (emu_cur.gdim)

void emu_syncthreads();
void emu_yield();          // one scheduling step (used by the emulated mbarrier wait and warp shuffles)
#define __syncthreads() emu_syncthreads()
#define __syncwarp(...) ((void)0)
// warp shuffles: every lane of the warp publishes its value, yields once (all live fibers advance one step per
// scheduler round), reads the source lane's slot, and yields again before any slot can be overwritten
double emu_shfl_exchange(double v, int src_lane);
static inline double __shfl_sync(unsigned, double v, int src) { return emu_shfl_exchange(v, src & 31); }
static inline double __shfl_down_sync(unsigned, double v, int off)
{
    const int lane = (int)(emu_cur.tid.x & 31);
    return emu_shfl_exchange(v, lane + off < 32 ? lane + off : lane);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

#define DB_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define DB_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_cur.smem)
#define DB_SET_SMEM_ATTR(kern) ((void)0)
