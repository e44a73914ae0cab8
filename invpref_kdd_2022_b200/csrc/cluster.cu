// EM environment re-assignment (train.py:846-879, 912-957).
//
// The reference runs K full forwards (K x the gather traffic, K wasted classifier passes), cats the
// K distance columns, adds the tie-break eps and argmins.  Here each sample's four rows are gathered
// ONCE, the K env-aware scores are formed in registers against E (staged in shared memory), and the
// first-min argmin, the env histogram (train.py:949) and the diff count (train.py:933-934) come out of
// the same pass.
#include "common.cuh"
#include "kernels.h"

namespace invpref {

namespace {

template <int VEC, int NV, int KT>
__global__ void __launch_bounds__(BLOCK) cluster_kernel(ClusterArgs a) {
    extern __shared__ float smem[];
    const int D = a.D, K = a.K;
    float* sE = smem;   // [K*D]
    __shared__ unsigned long long sHist[INVPREF_MAX_ENVS + 1];   // [K] histogram, [8] diff
    const int tid = threadIdx.x, lane = tid & (GROUP - 1);
    const unsigned gmask = group_mask();
    for (int t = tid; t < K * D; t += BLOCK) sE[t] = a.E[t];
    if (tid <= INVPREF_MAX_ENVS) sHist[tid] = 0ull;
    __syncthreads();

    const int64_t ngroups = (int64_t)gridDim.x * GROUPS_PER_BLOCK;
    for (int64_t n = (int64_t)blockIdx.x * GROUPS_PER_BLOCK + (tid >> 4); n < a.B; n += ngroups) {
        const int64_t u = a.users[n], it = a.items[n];
        Row<VEC, NV> ra, rc, rue, rie;
        load_row<VEC, NV, true>(ra, a.Uinv, u, D, lane);
        load_row<VEC, NV>(rc, a.Iinv, it, D, lane);
        load_row<VEC, NV, true>(rue, a.Uenv, u, D, lane);
        load_row<VEC, NV>(rie, a.Ienv, it, D, lane);
        const float y = a.scores[n];
        float z1 = 0.f;
        float z2[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) z2[k] = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
            if (d0 < D) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int x = j * VEC + v;
                    z1 += ra.x[x] * rc.x[x];
                    const float t = rue.x[x] * rie.x[x];
#pragma unroll
                    for (int k = 0; k < KT; ++k)
                        if (k < K) z2[k] += t * sE[k * D + d0 + v];
                }
            }
        }
        z1 = group_sum(z1, gmask);
#pragma unroll
        for (int k = 0; k < KT; ++k) z2[k] = group_sum(z2[k], gmask);
        if (lane == 0) {
            const float* eps = (a.perm_idx != nullptr) ? a.eps_table + a.perm_idx[n] * K : nullptr;
            const float s_inv = a.implicit ? sigmoidf_(z1) : z1;
            float best = 0.f;
            int arg = 0;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                if (k < K) {
                    float d;
                    if (a.implicit) {
                        const float s = s_inv * sigmoidf_(z2[k]);
                        d = -(y * fmaxf(logf(s), -100.f) + (1.f - y) * fmaxf(logf(1.f - s), -100.f));
                    } else {
                        const float r = (z1 + z2[k]) - y;
                        d = r * r;
                    }
                    if (eps != nullptr) d = d + eps[k];
                    if (k == 0 || d < best) { best = d; arg = k; }   // torch.argmin: first minimum wins
                }
            }
            a.new_envs[n] = (int64_t)arg;
            if (a.hist != nullptr) atomicAdd(&sHist[arg], 1ull);
            if (a.diff != nullptr && a.old_envs[n] != (int64_t)arg) atomicAdd(&sHist[INVPREF_MAX_ENVS], 1ull);
        }
    }
    __syncthreads();
    if (tid < K && a.hist != nullptr && sHist[tid] != 0ull) atomicAdd(&a.hist[tid], sHist[tid]);
    if (tid == INVPREF_MAX_ENVS && a.diff != nullptr && sHist[tid] != 0ull) atomicAdd(a.diff, sHist[tid]);
}

__global__ void __launch_bounds__(256) env_hist_kernel(const int64_t* __restrict__ envs, int64_t N, int K,
                                                       unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sH[INVPREF_MAX_ENVS];
    if (threadIdx.x < INVPREF_MAX_ENVS) sH[threadIdx.x] = 0u;
    __syncthreads();
    // a CTA handles 2^16 samples so the 32-bit shared counters cannot overflow
    const int64_t per = 1 << 16;
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = lo + per < N ? lo + per : N;
    for (int64_t n = lo + threadIdx.x; n < hi; n += blockDim.x) {
        const int64_t e = envs[n];
        if (e >= 0 && e < K) atomicAdd(&sH[(int)e], 1u);
    }
    __syncthreads();
    if ((int)threadIdx.x < K && sH[threadIdx.x] != 0u) atomicAdd(&hist[threadIdx.x], (unsigned long long)sH[threadIdx.x]);
}

// class_weights[k] = min(cnt_k + 1, N - 1) / N in double, rounded to fp32 (train.py:950-955);
// sample_weights[n] = class_weights[envs[n]] (train.py:956).
__global__ void __launch_bounds__(256) stat_envs_kernel(const int64_t* __restrict__ envs, int64_t N, int K,
                                                        const int64_t* __restrict__ hist,
                                                        float* __restrict__ class_weights,
                                                        float* __restrict__ sample_weights) {
    __shared__ float sCW[INVPREF_MAX_ENVS];
    if ((int)threadIdx.x < K) {
        double c = (double)hist[threadIdx.x] + 1.0;
        double cap = (double)(N - 1);
        double rate = (c < cap ? c : cap) / (double)N;
        sCW[threadIdx.x] = (float)rate;
        if (blockIdx.x == 0 && class_weights != nullptr) class_weights[threadIdx.x] = (float)rate;
    }
    __syncthreads();
    if (sample_weights == nullptr) return;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = envs[n];
        sample_weights[n] = (e >= 0 && e < K) ? sCW[(int)e] : 0.f;
    }
}

}  // namespace

int launch_cluster(const Geometry& g, const ClusterArgs& a, cudaStream_t stream) {
    size_t smem = (size_t)g.K * g.D * sizeof(float);
    int64_t need = (a.B + GROUPS_PER_BLOCK - 1) / GROUPS_PER_BLOCK;
    int grid = (int)(need < 1 ? 1 : (need < 148 * 16 ? need : 148 * 16));
#define CALL(V, N, KT_) cluster_kernel<V, N, KT_><<<grid, BLOCK, smem, stream>>>(a)
    INVPREF_DISPATCH_GEOM(g, CALL);
#undef CALL
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_env_hist(const int64_t* envs, int64_t N, int K, unsigned long long* hist, cudaStream_t stream) {
    int64_t per = 1 << 16;
    int grid = (int)((N + per - 1) / per);
    if (grid < 1) grid = 1;
    env_hist_kernel<<<grid, 256, 0, stream>>>(envs, N, K, hist);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_stat_envs(const int64_t* envs, int64_t N, int K, const int64_t* hist, float* class_weights,
                     float* sample_weights, cudaStream_t stream) {
    int64_t need = (N + 255) / 256;
    int grid = (int)(need < 1 ? 1 : (need < 148 * 8 ? need : 148 * 8));
    stat_envs_kernel<<<grid, 256, 0, stream>>>(envs, N, K, hist, class_weights, sample_weights);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
