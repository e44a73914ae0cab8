// Access-pattern probe for the fused user pass: what DRAM bandwidth does a B200 deliver for the
// pass's memory pattern when there is NO arithmetic and NO dependency between rows?
//
//   stream : read-modify-write of six [U, 64] fp32 tables front to back (what sweep_kernel does; the copy peak's twin)
//   rows   : the user pass's pattern -- for S sorted row ids (about a third of the U rows, as a 4 M-interaction batch
//            over 10 M users touches), read the row of six tables, write it back, write two rows of a sequential
//            "stash", and gather two random rows of two [I, 64] item tables per 1.27 segments; 16 lanes per row,
//            float4 per lane, ROWS_IN_FLIGHT independent rows per group
//   gather : read-only random gather of four rows per sample (the re-assignment kernel's pattern)
//   items  : the item pass's pattern -- for every touched item row (sorted), read-modify-write the row of six [I, 64]
//            tables, and per interaction gather two rows of the sequential-by-user-segment stash (random order) and
//            read a 32-byte g-pack record
//
// Prints one JSON line with GB/s of each.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rowrmw_probe
// tools/rowrmw_probe.cu ; run on the GPU box (needs ~22 GB).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int D4 = 16;   // float4 per 256-byte row

__global__ void __launch_bounds__(256) stream_kernel(float4* t0, float4* t1, float4* t2, float4* t3, float4* t4,
                                                      float4* t5, int64_t n4) {
    float4* tabs[6] = {t0, t1, t2, t3, t4, t5};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) v[t] = __ldcs(tabs[t] + i);
#pragma unroll
        for (int t = 0; t < 6; ++t) { v[t].x += 1.f; __stcs(tabs[t] + i, v[t]); }
    }
}

template <int INFLIGHT>
__global__ void __launch_bounds__(256) rows_kernel(float4* t0, float4* t1, float4* t2, float4* t3, float4* t4,
                                                    float4* t5, const float4* __restrict__ i0,
                                                    const float4* __restrict__ i1, float4* __restrict__ stash,
                                                    const int32_t* __restrict__ rows, const int32_t* __restrict__ items,
                                                    int64_t S, float* sink) {
    float4* tabs[6] = {t0, t1, t2, t3, t4, t5};
    const int lane = threadIdx.x & 15;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    float acc = 0.f;
    for (int64_t s0 = g * INFLIGHT; s0 < S; s0 += ng * INFLIGHT) {
        float4 v[INFLIGHT][6], it[INFLIGHT][2];
        int64_t r[INFLIGHT];
#pragma unroll
        for (int q = 0; q < INFLIGHT; ++q) {
            const int64_t s = s0 + q < S ? s0 + q : S - 1;
            r[q] = rows[s];
            const int64_t ir = items[s];
#pragma unroll
            for (int t = 0; t < 6; ++t) v[q][t] = tabs[t][r[q] * D4 + lane];
            it[q][0] = i0[ir * D4 + lane];
            it[q][1] = i1[ir * D4 + lane];
        }
#pragma unroll
        for (int q = 0; q < INFLIGHT; ++q) {
            if (s0 + q < S) {
                acc += it[q][0].x + it[q][1].y;
#pragma unroll
                for (int t = 0; t < 6; ++t) { v[q][t].x += 1.f; tabs[t][r[q] * D4 + lane] = v[q][t]; }
                stash[((s0 + q) * 2 + 0) * D4 + lane] = v[q][0];
                stash[((s0 + q) * 2 + 1) * D4 + lane] = v[q][1];
            }
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

// One segment in flight per group (the fused pass's depth) + prefetch.global.L2 of the rows of the group's segment DIST
// rounds ahead: lane l pulls one 128-byte line (6 table rows + 2 item rows, two lines each).
template <int DIST>
__global__ void __launch_bounds__(256) rows_pf_kernel(float4* t0, float4* t1, float4* t2, float4* t3, float4* t4,
                                                       float4* t5, const float4* __restrict__ i0,
                                                       const float4* __restrict__ i1, float4* __restrict__ stash,
                                                       const int32_t* __restrict__ rows,
                                                       const int32_t* __restrict__ items, int64_t S, float* sink) {
    float4* tabs[6] = {t0, t1, t2, t3, t4, t5};
    const int lane = threadIdx.x & 15;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    float acc = 0.f;
    for (int64_t s = g; s < S; s += ng) {
        const int64_t sp = s + DIST * ng;
        if (sp < S) {
            const int t = lane >> 1, half = lane & 1;
            const float4* base = t < 6 ? tabs[t] : (t == 6 ? i0 : i1);
            const int64_t rr = t < 6 ? rows[sp] : items[sp];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + rr * D4 + half * 8));
        }
        const int64_t r = rows[s], ir = items[s];
        float4 v[6], it0, it1;
#pragma unroll
        for (int t = 0; t < 6; ++t) v[t] = tabs[t][r * D4 + lane];
        it0 = i0[ir * D4 + lane];
        it1 = i1[ir * D4 + lane];
        acc += it0.x + it1.y;
#pragma unroll
        for (int t = 0; t < 6; ++t) { v[t].x += 1.f; tabs[t][r * D4 + lane] = v[t]; }
        stash[(s * 2 + 0) * D4 + lane] = v[0];
        stash[(s * 2 + 1) * D4 + lane] = v[1];
    }
    if (acc == 12345.678f) *sink = acc;
}

__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ t0, const float4* __restrict__ t1,
                                                      const float4* __restrict__ i0, const float4* __restrict__ i1,
                                                      const int32_t* __restrict__ rows, const int32_t* __restrict__ items,
                                                      int64_t N, float* sink) {
    const int lane = threadIdx.x & 15;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    float acc = 0.f;
    for (int64_t s0 = g * 4; s0 < N; s0 += ng * 4) {
        float4 v[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t s = s0 + q < N ? s0 + q : N - 1;
            const int64_t r = rows[s], ir = items[s];
            v[q][0] = t0[r * D4 + lane]; v[q][1] = t1[r * D4 + lane];
            v[q][2] = i0[ir * D4 + lane]; v[q][3] = i1[ir * D4 + lane];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += v[q][0].x + v[q][1].y + v[q][2].z + v[q][3].w;
    }
    if (acc == 12345.678f) *sink = acc;
}

__global__ void __launch_bounds__(256) items_kernel(float4* t0, float4* t1, float4* t2, float4* t3, float4* t4,
                                                     float4* t5, const float4* __restrict__ stash,
                                                     const float4* __restrict__ gpack, const int32_t* __restrict__ seg_row,
                                                     const int32_t* __restrict__ pseg, int64_t S, int64_t B,
                                                     float* sink) {
    float4* tabs[6] = {t0, t1, t2, t3, t4, t5};
    const int lane = threadIdx.x & 15;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4;
    const int64_t ng = ((int64_t)gridDim.x * blockDim.x) >> 4;
    float acc = 0.f;
    // (a) per interaction, in item-sorted order: two stash rows (by user segment: random) + the g-pack record.  Walked
    //     four interactions at a time by whichever group comes next, NOT segment by segment: a hot item's 40 000
    //     interactions would otherwise serialise on one group (the real pass pre-reduces long segments in chunks)
    for (int64_t k = g * 4; k < B; k += ng * 4) {
        float4 a[4][2], gp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t kk = k + q < B ? k + q : B - 1;
            const int64_t ps = pseg[kk];
            a[q][0] = stash[(ps * 2 + 0) * D4 + lane];
            a[q][1] = stash[(ps * 2 + 1) * D4 + lane];
            gp[q] = gpack[kk * 2 + (lane & 1)];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += a[q][0].x + a[q][1].y + gp[q].z;
    }
    // (b) per touched item row: read-modify-write of the row of the six tables
    for (int64_t s = g; s < S; s += ng) {
        const int64_t r = seg_row[s];
        float4 v[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) v[t] = tabs[t][r * D4 + lane];
#pragma unroll
        for (int t = 0; t < 6; ++t) { v[t].x += acc; tabs[t][r * D4 + lane] = v[t]; }
    }
    if (acc == 12345.678f) *sink = acc;
}

int main() {
    const int64_t U = 10000000, I = 1000000, B = 1 << 22;
    std::mt19937_64 gen(20220814);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    std::vector<int32_t> u(B), it(B);
    for (int64_t k = 0; k < B; ++k) {
        u[k] = (int32_t)std::min<double>(U - 1, std::floor(U * std::pow(uni(gen), 1.5)));    // SURVEY.md 8d generators
        it[k] = (int32_t)std::min<double>(I - 1, std::floor(I * std::pow(uni(gen), 3.0)));
    }
    std::vector<int32_t> rows(u);
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    const int64_t S = (int64_t)rows.size();
    std::vector<int32_t> seg_items(it.begin(), it.begin() + S);       // one random item row pair per segment ...
    float4* tabs[6];
    for (int t = 0; t < 6; ++t) { CK(cudaMalloc(&tabs[t], U * 256)); CK(cudaMemset(tabs[t], 0, U * 256)); }
    float4 *i0, *i1, *stash;
    CK(cudaMalloc(&i0, I * 256)); CK(cudaMalloc(&i1, I * 256)); CK(cudaMalloc(&stash, S * 512));
    CK(cudaMemset(i0, 0, I * 256)); CK(cudaMemset(i1, 0, I * 256));
    int32_t *d_rows, *d_items, *d_u, *d_it;
    CK(cudaMalloc(&d_rows, S * 4)); CK(cudaMalloc(&d_items, S * 4)); CK(cudaMalloc(&d_u, B * 4)); CK(cudaMalloc(&d_it, B * 4));
    CK(cudaMemcpy(d_rows, rows.data(), S * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_items, seg_items.data(), S * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_u, u.data(), B * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_it, it.data(), B * 4, cudaMemcpyHostToDevice));
    float* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto&& launch, int reps) -> float {
        launch(); launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        return ms / reps;
    };
    const int grid = 148 * 8;
    const float ms_stream = timeit([&] { stream_kernel<<<grid, 256>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], U * D4); }, 5);
    const double stream_gb = 6.0 * U * 256 * 2 / 1e9;
    // rows pattern: 6 rows read + 6 written + 2 stash rows written + 2 item rows read per segment, ids 8 B
    const double rows_gb = (double)S * (6 * 256 + 6 * 256 + 2 * 256 + 2 * 256 + 8) / 1e9;
    const float ms_r2 = timeit([&] { rows_kernel<2><<<grid, 256>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 10);
    const float ms_r4 = timeit([&] { rows_kernel<4><<<grid, 256>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 10);
    const float ms_r4b = timeit([&] { rows_kernel<4><<<148 * 16, 256>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 10);
    // the same pattern at the FUSED KERNEL'S occupancy: 2 CTAs of 256 threads per SM (100 KB of dynamic shared memory
    // each, unused), one wave of 296 CTAs, 1 / 2 / 4 independent segments in flight per 16-lane group
    cudaFuncSetAttribute(rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const float ms_o1 = timeit([&] { rows_kernel<1><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    const float ms_o2 = timeit([&] { rows_kernel<2><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    const float ms_o4 = timeit([&] { rows_kernel<4><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    printf("{\"user_pass_pattern_at_16_warps_per_sm\": {\"inflight1\": {\"ms\": %.4f, \"GBs\": %.1f}, \"inflight2\": {\"ms\": %.4f, \"GBs\": %.1f}, \"inflight4\": {\"ms\": %.4f, \"GBs\": %.1f}}}\n",
           ms_o1, rows_gb / ms_o1 * 1e3, ms_o2, rows_gb / ms_o2 * 1e3, ms_o4, rows_gb / ms_o4 * 1e3);
    cudaFuncSetAttribute(rows_pf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rows_pf_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(rows_pf_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const float ms_p2 = timeit([&] { rows_pf_kernel<2><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    const float ms_p4 = timeit([&] { rows_pf_kernel<4><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    const float ms_p8 = timeit([&] { rows_pf_kernel<8><<<296, 256, 100 * 1024>>>(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], tabs[5], i0, i1, stash, d_rows, d_items, S, sink); }, 5);
    printf("{\"user_pass_pattern_at_16_warps_per_sm_1_in_flight_plus_L2_prefetch\": {\"dist2\": {\"ms\": %.4f, \"GBs\": %.1f}, \"dist4\": {\"ms\": %.4f, \"GBs\": %.1f}, \"dist8\": {\"ms\": %.4f, \"GBs\": %.1f}}}\n",
           ms_p2, rows_gb / ms_p2 * 1e3, ms_p4, rows_gb / ms_p4 * 1e3, ms_p8, rows_gb / ms_p8 * 1e3);
    const float ms_g = timeit([&] { gather_kernel<<<148 * 16, 256>>>(tabs[0], tabs[1], i0, i1, d_u, d_it, B, sink); }, 10);
    const double gather_gb = (double)B * (4 * 256 + 8) / 1e9;
    // ---- item pass pattern: item segments of the same batch, partner = user segment of each interaction ----
    std::vector<int32_t> order(B);
    for (int64_t k = 0; k < B; ++k) order[k] = (int32_t)k;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return it[x] < it[y]; });
    std::vector<int32_t> i_row, i_off, i_pseg(B);
    for (int64_t k = 0; k < B; ++k) {
        const int32_t n = order[k];
        if (k == 0 || it[n] != it[order[k - 1]]) { i_row.push_back(it[n]); i_off.push_back((int32_t)k); }
        i_pseg[k] = (int32_t)(std::lower_bound(rows.begin(), rows.end(), u[n]) - rows.begin());
    }
    i_off.push_back((int32_t)B);
    const int64_t SI = (int64_t)i_row.size();
    float4* it_tabs[6];
    for (int t = 0; t < 6; ++t) { CK(cudaMalloc(&it_tabs[t], I * 256)); CK(cudaMemset(it_tabs[t], 0, I * 256)); }
    float4* gpack; CK(cudaMalloc(&gpack, B * 32)); CK(cudaMemset(gpack, 0, B * 32));
    int32_t *d_irow, *d_ioff, *d_ipseg;
    CK(cudaMalloc(&d_irow, SI * 4)); CK(cudaMalloc(&d_ioff, (SI + 1) * 4)); CK(cudaMalloc(&d_ipseg, B * 4));
    CK(cudaMemcpy(d_irow, i_row.data(), SI * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ioff, i_off.data(), (SI + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ipseg, i_pseg.data(), B * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(stash, 0, S * 512));
    const float ms_it = timeit([&] { items_kernel<<<148 * 16, 256>>>(it_tabs[0], it_tabs[1], it_tabs[2], it_tabs[3], it_tabs[4], it_tabs[5], stash, gpack, d_irow, d_ipseg, SI, B, sink); }, 10);
    const double items_gb = ((double)SI * 12 * 256 + (double)B * (2 * 256 + 32 + 4)) / 1e9;
    CK(cudaGetLastError());
    printf("{\"item_pass_pattern\": {\"item_segments\": %lld, \"GB\": %.3f, \"ms\": %.4f, \"GBs\": %.1f}}\n",
           (long long)SI, items_gb, ms_it, items_gb / ms_it * 1e3);
    printf("{\"probe\": \"row access patterns, no arithmetic\", \"segments\": %lld, \"stream_rmw_6_tables\": {\"GB\": %.3f, \"ms\": %.4f, \"GBs\": %.1f}, "
           "\"user_pass_pattern\": {\"GB\": %.3f, \"inflight2\": {\"ms\": %.4f, \"GBs\": %.1f}, \"inflight4\": {\"ms\": %.4f, \"GBs\": %.1f}, "
           "\"inflight4_grid16\": {\"ms\": %.4f, \"GBs\": %.1f}}, \"random_gather_4_rows\": {\"GB\": %.3f, \"ms\": %.4f, \"GBs\": %.1f}}\n",
           (long long)S, stream_gb, ms_stream, stream_gb / ms_stream * 1e3, rows_gb, ms_r2, rows_gb / ms_r2 * 1e3, ms_r4,
           rows_gb / ms_r4 * 1e3, ms_r4b, rows_gb / ms_r4b * 1e3, gather_gb, ms_g, gather_gb / ms_g * 1e3);
    return 0;
}
