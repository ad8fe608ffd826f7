// Error plumbing, ABI version and small stand-alone entry points of libsvb_vocoder.so.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace svb {
static thread_local char g_err[1024] = "no error";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

// w[r][:] = g[r] * v[r][:] / ||v[r][:]||_2   -- one CTA per row
__global__ void weight_norm_fold_kernel(const float *__restrict__ v, const float *__restrict__ g, long long inner,
                                        float *__restrict__ w) {
    __shared__ float red[32];
    const long long r = blockIdx.x;
    const float *vr = v + r * inner;
    float s = 0.f;
    for (long long i = threadIdx.x; i < inner; i += blockDim.x) s = fmaf(vr[i], vr[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) red[0] = g[r] / sqrtf(s);
    }
    __syncthreads();
    const float sc = red[0];
    for (long long i = threadIdx.x; i < inner; i += blockDim.x) w[r * inner + i] = vr[i] * sc;
}
}  // namespace svb

using namespace svb;

extern "C" const char *svb_last_error(void) { return get_error(); }
extern "C" int svb_abi_version(void) { return SVB_ABI_VERSION; }

extern "C" int svb_fold_weight_norm_host(const float *v_host, const float *g_host, int64_t d0, int64_t inner,
                                         float *w_host, int device) {
    SVB_CHECK(v_host && g_host && w_host && d0 > 0 && inner > 0, SVB_ERR_INVALID, "fold_weight_norm: bad argument");
    SVB_CUDA(cudaSetDevice(device));
    float *dv = nullptr, *dg = nullptr, *dw = nullptr;
    const size_t n = (size_t)d0 * inner;
    SVB_CUDA(cudaMalloc((void **)&dv, n * 4));
    SVB_CUDA(cudaMalloc((void **)&dg, d0 * 4));
    SVB_CUDA(cudaMalloc((void **)&dw, n * 4));
    cudaError_t e = cudaMemcpy(dv, v_host, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dg, g_host, d0 * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        weight_norm_fold_kernel<<<(unsigned)d0, 256>>>(dv, dg, inner, dw);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(w_host, dw, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(dv), cudaFree(dg), cudaFree(dw);
    if (e != cudaSuccess) {
        set_error("fold_weight_norm: %s", cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    return SVB_OK;
}

extern "C" int svb_fold_weight_norm_dev(const float *v_dev, const float *g_dev, int64_t d0, int64_t inner, float *w_dev,
                                        void *stream) {
    SVB_CHECK(v_dev && g_dev && w_dev && d0 > 0 && inner > 0, SVB_ERR_INVALID, "fold_weight_norm_dev: bad argument");
    weight_norm_fold_kernel<<<(unsigned)d0, 256, 0, as_stream(stream)>>>(v_dev, g_dev, inner, w_dev);
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}
