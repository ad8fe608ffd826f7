// Shared helpers for libsvb_vocoder.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/svb_vocoder.h"

namespace svb {

// ---- error plumbing (nothing throws across the C ABI) -------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define SVB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            svb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SVB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SVB_CHECK(cond, code, ...)                                                             \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            svb::set_error(__VA_ARGS__);                                                       \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

#define SVB_TRY(expr)                                                                          \
    do {                                                                                       \
        int _s = (expr);                                                                       \
        if (_s != SVB_OK) return _s;                                                           \
    } while (0)

// ---- activation layout: "C4T" -----------------------------------------------------------------
// Generator activations live in HBM channel-quad interleaved, time-major inside a quad:
//     buf[b][c / 4][PAD + t][c % 4]           (fp32, one "row" = 4 channels = 16 bytes)
// with PAD zero rows before t = 0 and zeros from t = T up to Tp - PAD .. Tp.  So
//   * a conv tap at dilation d is a row shift (the zero padding of Conv1d is the physical padding),
//   * for a fixed channel quad consecutive time steps are consecutive 16-byte rows: 128-bit
//     coalesced loads along time, and a [rows x 4ch] slab is ONE contiguous span, i.e. a single
//     cp.async.bulk per quad,
//   * the slab is exactly the K-major, no-swizzle tcgen05 core-matrix layout (8 rows x 16 B),
//     so any row shift is a legal 16-byte-aligned UMMA descriptor start address.
constexpr int kPad = 64;           // zero rows each side (>= largest ResBlock halo, 60)
constexpr int kTileT = 256;        // time-tile granularity of the allocation

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int c4t_rows(int T) { return round_up(T, kTileT) + 2 * kPad; }
__host__ __device__ inline size_t c4t_floats(int B, int C, int T) {
    return (size_t)B * (size_t)(C / 4) * (size_t)c4t_rows(T) * 4;
}

struct Act {            // a C4T activation tensor
    float *p = nullptr;
    int B = 0, C = 0, T = 0, Tp = 0;
    __host__ __device__ inline size_t quad_stride() const { return (size_t)Tp * 4; }
    __host__ __device__ inline size_t batch_stride() const { return (size_t)(C / 4) * Tp * 4; }
};

__device__ __forceinline__ float lrelu(float x, float slope) { return x >= 0.f ? x : x * slope; }
__device__ __forceinline__ float4 lrelu4(float4 v, float s) {
    return make_float4(lrelu(v.x, s), lrelu(v.y, s), lrelu(v.z, s), lrelu(v.w, s));
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace svb
