// Device-side helpers of the implicit evaluator (reference evaluate.py:94-135).
//
// The reference builds, per test batch, python lists of (row, item) pairs from per-user python sets to mask the
// train positives (rating = -1024), highlight the item pool (+1024) and, after top-k, looks every predicted
// item up in the user's ground-truth set.  Here the per-user item lists live on the device once, as CSR over
// user ids (ascending items within a user), and two small kernels do the masking and the look-ups.
#include "common.cuh"
#include "kernels.h"

namespace invpref {

namespace {

// one warp per batch row: rating[r, items(users[r])] = value  (add == 0)  or  += value  (add != 0)
__global__ void __launch_bounds__(256) mask_scores_kernel(float* __restrict__ rating, int64_t b, int64_t n_items,
                                                          const int64_t* __restrict__ users,
                                                          const int64_t* __restrict__ off,
                                                          const int64_t* __restrict__ items, float value, int add) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < b; r += warps) {
        const int64_t u = users[r];
        const int64_t lo = off[u], hi = off[u + 1];
        float* row = rating + r * n_items;
        for (int64_t p = lo + lane; p < hi; p += 32) {
            const int64_t it = items[p];
            if (it >= 0 && it < n_items) {      // lists hold unique items: no two lanes touch the same element
                if (add) row[it] += value;
                else row[it] = value;
            }
        }
    }
}

// one thread per (row, rank): hits[r, j] = top[r, j] in items(users[r]) (binary search in the ascending list)
__global__ void __launch_bounds__(256) hits_from_csr_kernel(const int64_t* __restrict__ top, int64_t b, int k,
                                                            const int64_t* __restrict__ users,
                                                            const int64_t* __restrict__ off,
                                                            const int64_t* __restrict__ items,
                                                            uint8_t* __restrict__ hits, int64_t* __restrict__ n_list) {
    const int64_t total = b * k;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = q / k;
        const int64_t u = users[r];
        int64_t lo = off[u], hi = off[u + 1];
        if (q - r * k == 0 && n_list != nullptr) n_list[r] = hi - lo;
        const int64_t want = top[q];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (items[mid] < want) lo = mid + 1; else hi = mid;
        }
        hits[q] = (lo < off[u + 1] && items[lo] == want) ? 1 : 0;
    }
}

}  // namespace

int launch_mask_scores(float* rating, int64_t b, int64_t n_items, const int64_t* users, const int64_t* off,
                       const int64_t* items, float value, int add, cudaStream_t stream) {
    int64_t need = (b + 7) / 8;
    int grid = (int)(need < 1 ? 1 : (need < 148 * 8 ? need : 148 * 8));
    mask_scores_kernel<<<grid, 256, 0, stream>>>(rating, b, n_items, users, off, items, value, add);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_hits_from_csr(const int64_t* top, int64_t b, int k, const int64_t* users, const int64_t* off,
                         const int64_t* items, uint8_t* hits, int64_t* n_list, cudaStream_t stream) {
    int64_t need = (b * k + 255) / 256;
    int grid = (int)(need < 1 ? 1 : (need < 148 * 8 ? need : 148 * 8));
    hits_from_csr_kernel<<<grid, 256, 0, stream>>>(top, b, k, users, off, items, hits, n_list);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
