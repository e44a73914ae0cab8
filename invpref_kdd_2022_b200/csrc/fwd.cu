// Fused forward kernels.
//
// fwd_train_kernel  = models.py:448-467 / 307-326 (5 gathers, products, row sums, sigmoids),
//                     models.py:206-209 (Linear(D->K) + log-softmax, by warp shuffles),
//                     train.py:797-822 (MSE/BCE, NLL, re-weighting, L1/L2 norms of the gathered rows)
//                     and the per-interaction part of the backward: g_z1, g_z2,
//                     g_logits, plus the batch reductions dW, db, dE -- all in ONE pass over the batch.
// fwd_only_kernel   = the same forward for callers that want s_inv / s_env / logp.
// predict_kernel    = models.py:534-539.
#include "common.cuh"
#include "kernels.h"

namespace invpref {

namespace {

// per-CTA partial layout (floats): [0] sum l_inv*w  [1] sum l_ea*w  [2] sum nll*w
// [3] sum x^2 over the 4 gathered rows  [4] sum |x|  [5] sum E[e]^2  [6] sum |E[e]|  [7] unused
// [8..16) db[k]   [16..24) count of env k   [24 .. 24+K*D) dW   [24+K*D .. 24+2*K*D) dE
constexpr int P_DB = 8, P_CNT = 16, P_DW = 24;

template <int VEC, int NV, int KT>
__global__ void __launch_bounds__(BLOCK, 3) fwd_train_kernel(FwdTrainArgs a) {
    extern __shared__ float smem[];
    const int D = a.D, K = a.K, KD = a.K * a.D;
    float* sE = smem;                 // [K*D]
    float* sW = smem + KD;            // [K*D] classifier weights
    float* sRed = smem + 2 * KD;      // [2*K*D] CTA reduction buffer for dW, dE
    float* sDE = smem + 4 * KD;       // [GROUPS_PER_BLOCK][K*D] per-group dE accumulators
    const int tid = threadIdx.x;
    const int lane = tid & (GROUP - 1);
    const unsigned gmask = group_mask();
    for (int t = tid; t < KD; t += BLOCK) { sE[t] = a.E[t]; sW[t] = a.W[t]; }
    for (int t = tid; t < 2 * KD; t += BLOCK) sRed[t] = 0.f;
    for (int t = tid; t < GROUPS_PER_BLOCK * KD; t += BLOCK) sDE[t] = 0.f;
    __syncthreads();
    float* myDE = sDE + (tid >> 4) * KD;   // only this group touches it: no atomics, fixed order

    float bk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) bk[k] = (k < K) ? a.b[k] : 0.f;

    // dW stays in registers (every interaction updates all K rows); dE[e] (one row per interaction,
    // chosen at run time) lives in this group's shared-memory slice
    float dW[KT][NV * VEC];
    float db[KT], cnt[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        db[k] = 0.f; cnt[k] = 0.f;
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) dW[k][x] = 0.f;
    }
    float s_linv = 0.f, s_lea = 0.f, s_nll = 0.f, s_sq = 0.f, s_abs = 0.f;

    const int64_t ngroups = (int64_t)gridDim.x * GROUPS_PER_BLOCK;
    for (int64_t n = (int64_t)blockIdx.x * GROUPS_PER_BLOCK + (tid >> 4); n < a.B; n += ngroups) {
        const int64_t u = a.users[n], it = a.items[n];
        const int e = (int)a.envs[n];
        Row<VEC, NV> ra, rc, rue, rie;
        load_row<VEC, NV>(ra, a.Uinv, u, D, lane);
        load_row<VEC, NV>(rc, a.Iinv, it, D, lane);
        load_row<VEC, NV>(rue, a.Uenv, u, D, lane);
        load_row<VEC, NV>(rie, a.Ienv, it, D, lane);
        const float y = a.generic ? 0.f : a.scores[n];
        const float w = (a.weights != nullptr) ? a.weights[n] : 1.f;

        float p[NV * VEC], t[NV * VEC];
        float z1 = 0.f, z2 = 0.f, sq = 0.f, ab = 0.f;
        float lg[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) lg[k] = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
            if (d0 < D) {
                float ee[VEC], wk[VEC];
                ldv<VEC>(sE + e * D + d0, ee);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int x = j * VEC + v;
                    p[x] = ra.x[x] * rc.x[x];
                    t[x] = rue.x[x] * rie.x[x];
                    z1 += p[x];
                    z2 += t[x] * ee[v];
                    sq += ra.x[x] * ra.x[x] + rc.x[x] * rc.x[x] + rue.x[x] * rue.x[x] + rie.x[x] * rie.x[x];
                    ab += fabsf(ra.x[x]) + fabsf(rc.x[x]) + fabsf(rue.x[x]) + fabsf(rie.x[x]);
                }
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    if (k < K) {
                        ldv<VEC>(sW + k * D + d0, wk);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) lg[k] += wk[v] * p[j * VEC + v];
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { p[j * VEC + v] = 0.f; t[j * VEC + v] = 0.f; }
            }
        }
        z1 = group_sum(z1, gmask);
        z2 = group_sum(z2, gmask);
#pragma unroll
        for (int k = 0; k < KT; ++k) lg[k] = (k < K) ? group_sum(lg[k], gmask) + bk[k] : -INFINITY;
        s_sq += sq; s_abs += ab;

        // softmax over K (models.py:208): ex[k] = exp(l_k - max), soft = ex / sum, lse = max + log(sum)
        float mx = lg[0];
#pragma unroll
        for (int k = 1; k < KT; ++k) mx = fmaxf(mx, lg[k]);
        float ex[KT];
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) { ex[k] = (k < K) ? expf(lg[k] - mx) : 0.f; se += ex[k]; }
        const float inv_se = 1.f / se;

        float g_z1, g_z2;
        float gl[KT];
        if (!a.generic) {
            const float wr = a.use_rec_rw ? w : 1.f;
            const float wc = a.use_class_rw ? w : 1.f;
            float s_inv, s2, s_env, l_inv, l_ea;
            if (a.implicit) {
                s_inv = sigmoidf_(z1);
                s2 = sigmoidf_(z2);
                s_env = s_inv * s2;
                // nn.BCELoss: log clamped at -100; backward clamps x(1-x) at 1e-12
                l_inv = -(y * fmaxf(logf(s_inv), -100.f) + (1.f - y) * fmaxf(logf(1.f - s_inv), -100.f));
                l_ea = -(y * fmaxf(logf(s_env), -100.f) + (1.f - y) * fmaxf(logf(1.f - s_env), -100.f));
                const float r1 = (s_inv - y) / fmaxf(s_inv * (1.f - s_inv), 1e-12f);
                const float r2 = (s_env - y) / fmaxf(s_env * (1.f - s_env), 1e-12f);
                const float g_s1 = wr * a.invB * (a.c_inv * r1 + a.c_ea * r2 * s2);
                const float g_s2 = wr * a.invB * a.c_ea * r2 * s_inv;
                g_z1 = g_s1 * s_inv * (1.f - s_inv);
                g_z2 = g_s2 * s2 * (1.f - s2);
            } else {
                s_inv = z1;
                s_env = z1 + z2;
                const float d1 = s_inv - y, d2 = s_env - y;
                l_inv = d1 * d1;
                l_ea = d2 * d2;
                g_z1 = wr * a.invB * 2.f * (a.c_inv * d1 + a.c_ea * d2);
                g_z2 = wr * a.invB * 2.f * a.c_ea * d2;
            }
            const float coef = a.c_env * wc * a.invB;
            float le = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                gl[k] = coef * (ex[k] * inv_se - ((k == e) ? 1.f : 0.f));
                if (k == e) le = lg[k];
            }
            if (lane == 0) {
                s_linv += l_inv * wr;
                s_lea += l_ea * wr;
                s_nll += ((mx + logf(se)) - le) * wc;      // -log_softmax[e]
            }
        } else {
            // generic autograd backward: upstream grads of (s_inv, s_env, logp)
            const float us1 = a.up_s_inv ? a.up_s_inv[n] : 0.f;
            const float us2 = a.up_s_env ? a.up_s_env[n] : 0.f;
            if (a.implicit) {
                const float s_inv = sigmoidf_(z1), s2 = sigmoidf_(z2);
                g_z1 = (us1 + us2 * s2) * s_inv * (1.f - s_inv);
                g_z2 = us2 * s_inv * s2 * (1.f - s2);
            } else {
                g_z1 = us1 + us2;
                g_z2 = us2;
            }
            float usum = 0.f;
            float ul[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                ul[k] = (a.up_logp && k < K) ? a.up_logp[n * K + k] : 0.f;
                usum += ul[k];
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) gl[k] = (k < K) ? ul[k] - ex[k] * inv_se * usum : 0.f;
        }

        // batch reductions: dW = g_logits^T p (not reversed), db, dE[e] += g_z2 * ue*ie
#pragma unroll
        for (int k = 0; k < KT; ++k) {
#pragma unroll
            for (int x = 0; x < NV * VEC; ++x) dW[k][x] += gl[k] * p[x];
            if (lane == 0) {
                db[k] += gl[k];
                cnt[k] += (k == e) ? 1.f : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
            if (d0 < D) {
                float acc[VEC];
                ldv<VEC>(myDE + e * D + d0, acc);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] += g_z2 * t[j * VEC + v];
                stv<VEC>(myDE + e * D + d0, acc);
            }
        }

        if (lane == 0) {
            float* gp = a.gpack + n * a.GS;
            float out[12];
            out[0] = g_z1;
            out[1] = g_z2;
            out[2] = __int_as_float(e);
#pragma unroll
            const float neg_alpha = a.dyn ? a.dyn->neg_alpha : a.neg_alpha;
#pragma unroll
            for (int k = 0; k < 9; ++k) out[3 + k] = (k < KT) ? neg_alpha * gl[k < KT ? k : 0] : 0.f;
            *reinterpret_cast<float4*>(gp) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(gp + 4) = make_float4(out[4], out[5], out[6], out[7]);
            if (KT > 5 && a.GS > 8) *reinterpret_cast<float4*>(gp + 8) = make_float4(out[8], out[9], out[10], out[11]);
        }
    }

    // ---- CTA reduction, fixed order => deterministic ----
    // dW: the two groups of a warp first, then the 8 warps one after another through smem;
    // dE: the 16 per-group slices are summed in group order
    const int warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < KT; ++k)
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) dW[k][x] += __shfl_xor_sync(0xffffffffu, dW[k][x], 16);
    for (int wsel = 0; wsel < BLOCK / 32; ++wsel) {
        if (warp == wsel && (tid & 31) < GROUP) {
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                if (k < K) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const int d0 = dim_of<VEC>(lane, j);
                        if (d0 < D) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) sRed[k * D + d0 + v] += dW[k][j * VEC + v];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int t2 = tid; t2 < KD; t2 += BLOCK) {
        float s = 0.f;
        for (int gsel = 0; gsel < GROUPS_PER_BLOCK; ++gsel) s += sDE[gsel * KD + t2];
        sRed[KD + t2] = s;
    }
    // scalars: reduce inside the warp, then across warps in order
    __shared__ float sScal[BLOCK / 32][24];
    float sc[24];
    sc[0] = s_linv; sc[1] = s_lea; sc[2] = s_nll; sc[3] = s_sq; sc[4] = s_abs; sc[5] = 0.f; sc[6] = 0.f; sc[7] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[8 + k] = (k < KT) ? db[k < KT ? k : 0] : 0.f;
        sc[16 + k] = (k < KT) ? cnt[k < KT ? k : 0] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 24; ++q) sc[q] = warp_sum(sc[q]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int q = 0; q < 24; ++q) sScal[warp][q] = sc[q];
    }
    __syncthreads();
    float* out = a.partials + (int64_t)blockIdx.x * a.P;
    if (tid < 24) {
        float s = 0.f;
        for (int wsel = 0; wsel < BLOCK / 32; ++wsel) s += sScal[wsel][tid];
        out[tid] = s;
    }
    for (int t2 = tid; t2 < 2 * KD; t2 += BLOCK) out[P_DW + t2] = sRed[t2];
}

template <int VEC, int NV, int KT>
__global__ void __launch_bounds__(BLOCK) fwd_only_kernel(FwdOnlyArgs a) {
    extern __shared__ float smem[];
    const int D = a.D, K = a.K;
    float* sE = smem;
    float* sW = smem + K * D;
    const int tid = threadIdx.x, lane = tid & (GROUP - 1);
    const unsigned gmask = group_mask();
    for (int t = tid; t < K * D; t += BLOCK) { sE[t] = a.E[t]; sW[t] = a.W[t]; }
    __syncthreads();
    const bool want_cls = a.logp != nullptr;
    const int64_t ngroups = (int64_t)gridDim.x * GROUPS_PER_BLOCK;
    for (int64_t n = (int64_t)blockIdx.x * GROUPS_PER_BLOCK + (tid >> 4); n < a.B; n += ngroups) {
        const int64_t u = a.users[n], it = a.items[n];
        const int e = a.envs ? (int)a.envs[n] : 0;
        Row<VEC, NV> ra, rc, rue, rie;
        load_row<VEC, NV>(ra, a.Uinv, u, D, lane);
        load_row<VEC, NV>(rc, a.Iinv, it, D, lane);
        load_row<VEC, NV>(rue, a.Uenv, u, D, lane);
        load_row<VEC, NV>(rie, a.Ienv, it, D, lane);
        float z1 = 0.f, z2 = 0.f;
        float lg[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) lg[k] = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int x = j * VEC + v;
                const float ee = (d0 < D) ? sE[e * D + d0 + v] : 0.f;
                const float p = ra.x[x] * rc.x[x];
                z1 += p;
                z2 += rue.x[x] * rie.x[x] * ee;
                if (want_cls) {
#pragma unroll
                    for (int k = 0; k < KT; ++k) lg[k] += ((k < K && d0 < D) ? sW[k * D + d0 + v] : 0.f) * p;
                }
            }
        }
        z1 = group_sum(z1, gmask);
        z2 = group_sum(z2, gmask);
        float s_inv, s_env;
        if (a.implicit) {
            s_inv = sigmoidf_(z1);
            s_env = s_inv * sigmoidf_(z2);
        } else {
            s_inv = z1;
            s_env = z1 + z2;
        }
        if (want_cls) {
#pragma unroll
            for (int k = 0; k < KT; ++k) lg[k] = (k < K) ? group_sum(lg[k], gmask) + a.b[k] : -INFINITY;
            float mx = lg[0];
#pragma unroll
            for (int k = 1; k < KT; ++k) mx = fmaxf(mx, lg[k]);
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) se += (k < K) ? expf(lg[k] - mx) : 0.f;
            const float lse = mx + logf(se);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (k < K) a.logp[n * K + k] = lg[k] - lse;
            }
        }
        if (lane == 0) {
            if (a.s_inv) a.s_inv[n] = s_inv;
            if (a.s_env) a.s_env[n] = s_env;
        }
    }
}

template <int VEC, int NV>
__global__ void __launch_bounds__(BLOCK) predict_kernel(const float* __restrict__ Uinv, const float* __restrict__ Iinv,
                                                        const int64_t* __restrict__ users,
                                                        const int64_t* __restrict__ items, int64_t B, int D,
                                                        float* __restrict__ score) {
    const int tid = threadIdx.x, lane = tid & (GROUP - 1);
    const unsigned gmask = group_mask();
    const int64_t ngroups = (int64_t)gridDim.x * GROUPS_PER_BLOCK;
    for (int64_t n = (int64_t)blockIdx.x * GROUPS_PER_BLOCK + (tid >> 4); n < B; n += ngroups) {
        Row<VEC, NV> ra, rc;
        load_row<VEC, NV>(ra, Uinv, users[n], D, lane);
        load_row<VEC, NV>(rc, Iinv, items[n], D, lane);
        float z = 0.f;
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) z += ra.x[x] * rc.x[x];
        z = group_sum(z, gmask);
        if (lane == 0) score[n] = z;
    }
}

int grid_for_groups(int64_t B, int max_blocks) {
    int64_t need = (B + GROUPS_PER_BLOCK - 1) / GROUPS_PER_BLOCK;
    if (need < 1) need = 1;
    return (int)(need < max_blocks ? need : max_blocks);
}

}  // namespace

int fwd_train_grid(int64_t B) { return grid_for_groups(B, 148 * 4); }

int launch_fwd_train(const Geometry& g, const FwdTrainArgs& a, int grid, cudaStream_t stream) {
    size_t smem = (size_t)(4 + GROUPS_PER_BLOCK) * g.K * g.D * sizeof(float);
#define CALL(V, N, KT_)                                                                                         \
    do {                                                                                                        \
        INVPREF_SET_SMEM_ONCE((fwd_train_kernel<V, N, KT_>), smem);                                             \
        fwd_train_kernel<V, N, KT_><<<grid, BLOCK, smem, stream>>>(a);                                          \
    } while (0)
    INVPREF_DISPATCH_GEOM(g, CALL);
#undef CALL
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_fwd_only(const Geometry& g, const FwdOnlyArgs& a, cudaStream_t stream) {
    size_t smem = (size_t)2 * g.K * g.D * sizeof(float);
    int grid = grid_for_groups(a.B, 148 * 16);
#define CALL(V, N, KT_) fwd_only_kernel<V, N, KT_><<<grid, BLOCK, smem, stream>>>(a)
    INVPREF_DISPATCH_GEOM(g, CALL);
#undef CALL
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_predict(const Geometry& g, const float* Uinv, const float* Iinv, const int64_t* users, const int64_t* items,
                   int64_t B, float* score, cudaStream_t stream) {
    int grid = grid_for_groups(B, 148 * 16);
    int D = g.D;
#define CALL(V, N) predict_kernel<V, N><<<grid, BLOCK, 0, stream>>>(Uinv, Iinv, users, items, B, D, score)
    INVPREF_DISPATCH_VN(g, CALL);
#undef CALL
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
