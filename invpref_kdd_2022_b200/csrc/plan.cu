// Sort-segment plan: stable sort of a batch by row id, unique rows, segment offsets, chunking of
// long segments, touched-row bitmap.  Replaces the duplicate handling of ATen's
// embedding_dense_backward (autograd of the reference's models.py:449-455).
//
// The sort and the prefix sums are hand-written (sort.cuh: stable LSD radix sort over the bits a row id can have,
// reduce-then-scan prefix sums; no library call), so `perm` is bit-equal to torch.sort(ids, stable=True).  The plan
// depends only on the (fixed) batch slicing of utils.py:12-19 and is built once per batch by the trainer, outside the
// per-step hot loop (the e2e path builds it on a loader stream, one step ahead).
#include "common.cuh"
#include "sort.cuh"

namespace invpref {

namespace {

__global__ void prep_keys_kernel(const int64_t* __restrict__ ids, int64_t B, int64_t rows, int32_t* __restrict__ keys,
                                 int32_t* __restrict__ counters) {
    int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n == 0) { counters[0] = 0; counters[1] = 0; }
    if (n >= B) return;
    int64_t id = ids[n];
    if (id < 0 || id >= rows) {
        counters[2] = 1;   // id out of range: flagged, clamped so that nothing reads out of bounds
        id = id < 0 ? 0 : rows - 1;
    }
    keys[n] = (int32_t)id;
}

// invpref_check_ids: bit 0 / 1 / 2 of *flag = some user / item / env id outside its table
__global__ void __launch_bounds__(256) check_ids_kernel(const int64_t* __restrict__ users,
                                                        const int64_t* __restrict__ items,
                                                        const int64_t* __restrict__ envs, int64_t B, int64_t n_users,
                                                        int64_t n_items, int64_t n_envs, int32_t* __restrict__ flag) {
    int bad = 0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < B; n += (int64_t)gridDim.x * blockDim.x) {
        if (users != nullptr) { const int64_t u = users[n]; if (u < 0 || u >= n_users) bad |= 1; }
        if (items != nullptr) { const int64_t i = items[n]; if (i < 0 || i >= n_items) bad |= 2; }
        if (envs != nullptr) { const int64_t e = envs[n]; if (e < 0 || e >= n_envs) bad |= 4; }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad != 0 && (threadIdx.x & 31) == 0) atomicOr(flag, bad);
}

// other_seg_of (nullable): seg_of of the OTHER side's finished plan -> pseg[k] = segment, on the other side, of the
// partner row of sorted position k (read by the item pass to index the stash of user rows; the user side, built
// first, has no reader for it and passes nullptr).
__global__ void write_segments_kernel(const int32_t* __restrict__ sorted, const int32_t* __restrict__ segid,
                                      const int32_t* __restrict__ perm, const int64_t* __restrict__ other_ids,
                                      int64_t other_rows, const int32_t* __restrict__ other_seg_of, int64_t B,
                                      PlanSide p) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool in = k < B;
    const int lane = threadIdx.x & 31;
    const int32_t key = in ? sorted[k] : -1;
    const int32_t s = in ? segid[k] - 1 : 0;
    const bool head = in && ((k == 0) || (sorted[k - 1] != key));
    // touched-row bitmap: keys are sorted, so the lanes of one 32-row word are consecutive -> OR them with a
    // segmented shuffle reduction and issue one atomic per (warp, word) instead of one per segment head
    {
        const int word = in ? (key >> 5) : -1;
        unsigned bits = head ? (1u << (key & 31)) : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned ob = __shfl_down_sync(0xffffffffu, bits, o);
            const int ow = __shfl_down_sync(0xffffffffu, word, o);
            if (lane + o < 32 && ow == word) bits |= ob;
        }
        const int pw = __shfl_up_sync(0xffffffffu, word, 1);
        if (in && (lane == 0 || pw != word) && bits != 0u) atomicOr(&p.touched[word], bits);
    }
    if (!in) return;
    const int32_t orig = perm[k];
    p.seg_of[orig] = s;
    if (other_seg_of != nullptr) p.pseg[k] = other_seg_of[orig];
    if (head) {
        p.seg_row[s] = key;
        p.seg_off[s] = (int32_t)k;
    }
    if (k == B - 1) {
        p.counters[0] = s + 1;
        p.seg_off[s + 1] = (int32_t)B;
    }
    if (other_ids != nullptr) {
        int64_t o = other_ids[orig];
        if (o < 0 || o >= other_rows) {
            p.counters[2] = 1;   // flagged like an out-of-range key (prep_keys_kernel), clamped
            o = o < 0 ? 0 : other_rows - 1;
        }
        p.partner[k] = (int32_t)o;
    }
}

__global__ void chunk_counts_kernel(PlanSide p, int chunk) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s > p.max_seg) return;
    int32_t n_seg = p.counters[0];
    int32_t c = 0;
    if (s < n_seg) {
        int32_t len = p.seg_off[s + 1] - p.seg_off[s];
        if (len > 2 * chunk) c = (len + chunk - 1) / chunk;
    }
    p.seg_chunk[s] = c;
}

// Cost prefix of the segments before s: strictly increasing in s (a long segment, pre-reduced by the chunks
// kernels, counts len - ceil(len / chunk) * chunk / 2 > 0 of its interactions).
__device__ __forceinline__ int64_t seg_cost(const PlanSide& p, int64_t s, int half_chunk) {
    return (int64_t)PLAN_CSEG * s + p.seg_off[s] - (int64_t)half_chunk * p.seg_chunk[s];
}

// One thread per segment s (one launch instead of two):
//  * descriptors: seg_desc[s], the chunk descriptors of a long segment, the hot list;
//  * work ranges: range_start[r] = smallest s in [0, n_seg] with seg_cost(s) * R >= r * seg_cost(n_seg), r = 0..R.
//    Thread s writes the r's in (q(s-1), q(s)], q(s) = seg_cost(s) * R / total: every r is written exactly once.  q(s-1)
//    comes from the neighbouring lane (one 64-bit division per thread instead of two).
__global__ void write_chunks_ranges_kernel(PlanSide p, int chunk, int has_partner, int n_ranges) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int32_t n_seg = p.counters[0];
    const int half_chunk = chunk / 2;
    if (s == 0) {
        p.counters[1] = p.seg_chunk[n_seg];
        p.counters[3] = n_ranges;
        p.range_start[0] = 0;
    }
    {
        const int64_t total = seg_cost(p, n_seg, half_chunk);
        const bool on = s <= n_seg;
        const int64_t q = on ? seg_cost(p, s, half_chunk) * n_ranges / total : 0;
        int64_t q_prev = __shfl_up_sync(0xffffffffu, q, 1);
        if ((threadIdx.x & 31) == 0 && on && s >= 1) q_prev = seg_cost(p, s - 1, half_chunk) * n_ranges / total;
        if (on && s >= 1)
            for (int64_t r = q_prev + 1; r <= q; ++r) p.range_start[r] = (int32_t)s;
    }
    if (s >= n_seg) return;
    int32_t c0 = p.seg_chunk[s], c1 = p.seg_chunk[s + 1];
    int32_t beg = p.seg_off[s], end = p.seg_off[s + 1];
    {   // {row, begin, end, perm[begin]}, {partner[begin], perm[begin+1], partner[begin+1], -}
        const bool two = beg + 1 < end;
        int4* d = reinterpret_cast<int4*>(p.seg_desc) + 2 * s;
        d[0] = make_int4(p.seg_row[s], beg, end, p.perm[beg]);
        d[1] = make_int4(has_partner ? p.partner[beg] : 0, two ? p.perm[beg + 1] : 0,
                         (two && has_partner) ? p.partner[beg + 1] : 0, 0);
    }
    if (c1 == c0) return;
    if (c1 - c0 > HOT_CHUNKS) p.hot_list[atomicAdd(&p.counters[4], 1)] = (int32_t)s;   // order is irrelevant
    for (int32_t c = c0; c < c1; ++c) {
        int32_t b = beg + (c - c0) * chunk;
        int32_t e = b + chunk < end ? b + chunk : end;
        reinterpret_cast<int4*>(p.chunk_desc)[c] = make_int4((int32_t)s, b, e, 0);
    }
}

__global__ void empty_plan_kernel(PlanSide p) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        p.counters[0] = 0; p.counters[1] = 0; p.counters[3] = 0;
        p.seg_off[0] = 0; p.seg_chunk[0] = 0; p.range_start[0] = 0;
    }
}

__global__ void widen_kernel(const int32_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n, const int32_t* limit,
                             int64_t extra) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t lim = limit ? (int64_t)(*limit) + extra : n;
    if (k < n && k < lim) dst[k] = src[k];
}

__global__ void widen_scalar_kernel(const int32_t* src, int64_t* dst) { *dst = *src; }

struct SortTmp {
    int32_t *keys_in, *keys_out, *keys_tmp, *vals_tmp, *scan;
    char *sort_scratch, *scan_scratch;
};

size_t sort_scratch_bytes_for(int64_t B, int64_t max_seg_plus1) {
    return psort::rs_scratch_bytes(B) + psort::sc_scratch_bytes(B > max_seg_plus1 ? B : max_seg_plus1);
}

SortTmp carve_sort_tmp(char* base, int64_t B) {
    SortTmp t;
    size_t arr = align_up((size_t)B * 4);
    t.keys_in = (int32_t*)base;
    t.keys_out = (int32_t*)(base + arr);
    t.keys_tmp = (int32_t*)(base + 2 * arr);
    t.vals_tmp = (int32_t*)(base + 3 * arr);
    t.scan = (int32_t*)(base + 4 * arr);
    t.sort_scratch = base + 5 * arr;
    t.scan_scratch = t.sort_scratch + psort::rs_scratch_bytes(B);
    return t;
}

inline int bits_for(int64_t rows) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < rows) ++b;
    return b;
}

inline unsigned grid_for(int64_t n, int block = 256) { return (unsigned)((n + block - 1) / block); }

}  // namespace

size_t sort_tmp_bytes_for(int64_t B, int64_t max_rows) {
    int64_t S = plan_max_seg(B, max_rows);
    return 5 * align_up((size_t)B * 4) + align_up(sort_scratch_bytes_for(B, S + 1));
}

// Builds one side of a plan.  All work is enqueued on `stream`; nothing is read back.
int build_plan_side(const int64_t* ids, const int64_t* other_ids, int64_t other_rows, const int32_t* other_seg_of,
                    PlanSide p, char* tmp, size_t tmp_bytes, cudaStream_t stream) {
    int64_t B = p.B;
    cudaMemsetAsync(p.touched, 0, (size_t)((p.rows + 31) / 32) * 4, stream);
    cudaMemsetAsync(p.counters, 0, 16 * 4, stream);
    if (B == 0) {
        empty_plan_kernel<<<1, 32, 0, stream>>>(p);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
    }
    if (tmp_bytes < sort_tmp_bytes_for(B, p.rows)) return INVPREF_ERR_WORKSPACE;
    SortTmp t = carve_sort_tmp(tmp, B);
    prep_keys_kernel<<<grid_for(B), 256, 0, stream>>>(ids, B, p.rows, t.keys_in, p.counters);
    int n_launch = 4;
    // perm = stable argsort of the row ids
    n_launch += psort::radix_sort_pairs(t.keys_in, nullptr, t.keys_out, p.perm, t.keys_tmp, t.vals_tmp, B,
                                        bits_for(p.rows), t.sort_scratch, stream);
    // scan[k] = 1 + segment of sorted position k (inclusive sum of the head flags)
    n_launch += psort::prefix_sum<1, true>(t.keys_out, t.scan, B, t.scan_scratch, stream);
    write_segments_kernel<<<grid_for(B), 256, 0, stream>>>(t.keys_out, t.scan, p.perm, other_ids, other_rows,
                                                           other_seg_of, B, p);
    chunk_counts_kernel<<<grid_for(p.max_seg + 1), 256, 0, stream>>>(p, chunk_for(B));
    n_launch += psort::prefix_sum<0, false>(p.seg_chunk, p.seg_chunk, p.max_seg + 1, t.scan_scratch, stream);
    write_chunks_ranges_kernel<<<grid_for(p.max_seg + 1), 256, 0, stream>>>(p, chunk_for(B), other_ids != nullptr,
                                                                           (int)plan_ranges(B));
    count_launch(n_launch);
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_check_ids(const int64_t* users, const int64_t* items, const int64_t* envs, int64_t B, int64_t n_users,
                     int64_t n_items, int64_t n_envs, int32_t* flag, cudaStream_t stream) {
    const int64_t need = (B + 255) / 256;
    const int grid = (int)(need < 1 ? 1 : (need < 148 * 8 ? need : 148 * 8));
    check_ids_kernel<<<grid, 256, 0, stream>>>(users, items, envs, B, n_users, n_items, n_envs, flag);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int build_segments_i64(const int64_t* ids, int64_t B, int64_t rows, int64_t* perm, int64_t* seg_row, int64_t* seg_off,
                       int64_t* n_seg, char* ws, size_t ws_bytes, cudaStream_t stream) {
    int64_t Bp = B > 0 ? B : 1;
    size_t side = plan_side_bytes(Bp, rows);
    size_t need = side + sort_tmp_bytes_for(Bp, rows);
    if (ws_bytes < need) return INVPREF_ERR_WORKSPACE;
    PlanSide p = carve_plan_side(ws, Bp, rows);
    p.B = B;
    int rc = build_plan_side(ids, nullptr, 0, nullptr, p, ws + side, ws_bytes - side, stream);
    if (rc != INVPREF_OK) return rc;
    if (B > 0) {
        widen_kernel<<<grid_for(B), 256, 0, stream>>>(p.perm, perm, B, nullptr, 0);
        widen_kernel<<<grid_for(p.max_seg), 256, 0, stream>>>(p.seg_row, seg_row, p.max_seg, p.counters, 0);
        widen_kernel<<<grid_for(p.max_seg + 1), 256, 0, stream>>>(p.seg_off, seg_off, p.max_seg + 1, p.counters, 1);
        count_launch(3);
    } else {
        widen_kernel<<<1, 32, 0, stream>>>(p.seg_off, seg_off, 1, nullptr, 0);
        count_launch();
    }
    widen_scalar_kernel<<<1, 1, 0, stream>>>(p.counters, n_seg);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
