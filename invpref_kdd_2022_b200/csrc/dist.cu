// Building blocks of the multi-GPU paths (the reference is single-GPU: no counterpart there).
//
//  adam_dense_kernel       : torch.optim.Adam on a flat tensor whose gradient was materialised because it
//                            had to be reduced across ranks first (NCCL all-reduce / all-to-all).
//  gather_rows_kernel      : packs the rows a peer asked for into a contiguous send buffer.
//  scatter_add_rows_kernel : adds one peer's partial row gradients into the owner's shard; indices are
//                            unique within a call, peers are applied in rank order => deterministic.
#include "common.cuh"
#include "kernels.h"

namespace invpref {

namespace {

__global__ void __launch_bounds__(256) adam_dense_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                         float* __restrict__ v, const float* __restrict__ grad,
                                                         int64_t n, AdamScalars s, int vec_ok) {
    const int64_t n4 = vec_ok ? (n >> 2) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t q = t0; q < n4; q += stride) {
        float p[4], mm[4], vv[4], g[4];
        ldv_stream<4>(theta + q * 4, p);
        ldv_stream<4>(m + q * 4, mm);
        ldv_stream<4>(v + q * 4, vv);
        ldv_stream<4>(grad + q * 4, g);
#pragma unroll
        for (int x = 0; x < 4; ++x) adam_update(p[x], mm[x], vv[x], g[x], s);
        stv<4>(theta + q * 4, p);
        stv_stream<4>(m + q * 4, mm);
        stv_stream<4>(v + q * 4, vv);
    }
    for (int64_t q = n4 * 4 + t0; q < n; q += stride) {
        float p = theta[q], mm = m[q], vv = v[q];
        adam_update(p, mm, vv, grad[q], s);
        theta[q] = p; m[q] = mm; v[q] = vv;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table,
                                                          const int64_t* __restrict__ rows, int64_t n, int dim,
                                                          float* __restrict__ out) {
    const int per_row = dim / VEC;
    const int64_t total = n * per_row;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = q / per_row;
        const int c = (int)(q - j * per_row) * VEC;
        float r[VEC];
        ldv<VEC>(table + rows[j] * dim + c, r);
        stv<VEC>(out + j * dim + c, r);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src,
                                                               const int64_t* __restrict__ rows, int64_t n, int dim,
                                                               float* __restrict__ table) {
    const int per_row = dim / VEC;
    const int64_t total = n * per_row;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = q / per_row;
        const int c = (int)(q - j * per_row) * VEC;
        float a[VEC], b[VEC];
        ldv<VEC>(src + j * dim + c, a);
        float* dst = table + rows[j] * dim + c;
        ldv<VEC>(dst, b);
#pragma unroll
        for (int x = 0; x < VEC; ++x) b[x] += a[x];
        stv<VEC>(dst, b);
    }
}

inline int grid_1d(int64_t work, int max_blocks = 148 * 16) {
    int64_t need = (work + 255) / 256;
    if (need < 1) need = 1;
    return (int)(need < max_blocks ? need : max_blocks);
}

}  // namespace

int launch_adam_dense(float* theta, float* m, float* v, const float* grad, int64_t n, const AdamScalars& s,
                      cudaStream_t stream) {
    const int vec_ok = (((uintptr_t)theta | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad) % 16) == 0;
    adam_dense_kernel<<<grid_1d(vec_ok ? n / 4 + 4 : n), 256, 0, stream>>>(theta, m, v, grad, n, s, vec_ok);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_gather_rows(const float* table, const int64_t* rows, int64_t n, int dim, float* out, cudaStream_t stream) {
    const bool v4 = dim % 4 == 0 && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (v4) gather_rows_kernel<4><<<grid_1d(n * (dim / 4)), 256, 0, stream>>>(table, rows, n, dim, out);
    else gather_rows_kernel<1><<<grid_1d(n * dim), 256, 0, stream>>>(table, rows, n, dim, out);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_scatter_add_rows(const float* src, const int64_t* rows, int64_t n, int dim, float* table,
                            cudaStream_t stream) {
    const bool v4 = dim % 4 == 0 && ((uintptr_t)table % 16 == 0) && ((uintptr_t)src % 16 == 0);
    if (v4) scatter_add_rows_kernel<4><<<grid_1d(n * (dim / 4)), 256, 0, stream>>>(src, rows, n, dim, table);
    else scatter_add_rows_kernel<1><<<grid_1d(n * dim), 256, 0, stream>>>(src, rows, n, dim, table);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
