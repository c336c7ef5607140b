// Error plumbing, version / device probes and small utility kernels.
#include "db_common.cuh"
#include <cstdarg>

static char g_last_error[512] = "";

void db_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int db_check_launch(const char* what)
{
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        db_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(err));
        return 2;
    }
    return 0;
}

extern "C" const char* db_last_error(void) { return g_last_error; }
extern "C" int db_version(void) { return 100; }

extern "C" int db_device_arch(void)
{
#ifdef DB_EMU
    return 0;
#else
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { db_set_error("no CUDA device"); return -1; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { db_set_error("cudaGetDeviceProperties failed"); return -1; }
    return prop.major * 10 + prop.minor;
#endif
}

// |x| max reduction: per-block partial maxima combined with atomicMax on the ordered-int representation
__global__ void k_absmax(const double* __restrict__ x, int64_t count, unsigned long long* __restrict__ out)
{
    DB_SMEM(double, red);
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double v = fabs(x[i]);
        m = (v > m) ? v : m;
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { double o = red[threadIdx.x + s]; if (o > red[threadIdx.x]) red[threadIdx.x] = o; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // non-negative doubles order like their bit patterns
        unsigned long long bits;
        double v = red[0];
        memcpy(&bits, &v, sizeof(bits));
#ifdef DB_EMU
        if (bits > *out) *out = bits;
#else
        atomicMax(out, bits);
#endif
    }
}

extern "C" int db_absmax(const double* x, int64_t count, double* out, void* stream)
{
    // `out` must be zero-initialised by the caller (0.0 has an all-zero bit pattern)
    if (count <= 0) return 0;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    DB_LAUNCH(k_absmax, dim3((unsigned)blocks), dim3(256), 256 * sizeof(double), stream, x, count, reinterpret_cast<unsigned long long*>(out));
    return db_check_launch("absmax");
}
