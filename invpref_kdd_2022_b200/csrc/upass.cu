// Fused USER PASS: forward + losses + user-side segmented backward + Adam in one sweep over the batch in
// user-sorted order.
//
// Replaces fwd_train_kernel + bwd_rows_kernel(user side) of the unfused path (same reference code:
// models.py:448-467 / 307-326, 206-209; train.py:797-834; embedding_dense_backward; optim.Adam).
// One 16-lane group owns one user segment: the two user rows are loaded ONCE per segment (not once per
// interaction and again for Adam), the item rows once per interaction (not twice), and the per-interaction
// g-pack is only written (for the item pass), never read back here.  The classifier part of the user
// gradient is factored per segment:
//     Q_k[d] = sum_n g_logits[n,k] * c_n[d]        (c = item invariant row)
//     dUinv[u] += sum_n g_z1[n] c_n  +  (-alpha) * sum_k W[k,:] (.) Q_k          (gradient reversal)
//     dW[k,:]  += a_u (.) Q_k                                                     (not reversed)
// so W is applied once per segment instead of once per interaction.
//
//  upass_chunks_kernel : CHUNK-sized pieces of user segments longer than LONG_T (rare: hot users).
//  upass_rows_kernel   : every user segment; long ones only add up their chunk partials.
// Both write per-CTA partial sums (losses, norms, db, env counts, dW, dE) for tail_kernel; all sums run in
// a fixed order (no floating-point atomics).
#include "common.cuh"
#include "kernels.h"
#include "lossmath.cuh"

namespace invpref {

namespace {

constexpr int P_DB = 8, P_CNT = 16, P_DW = 24;

// Per-lane sums that persist over all segments a group handles.  The eleven-plus scalar sums (three losses,
// db[K], env counts[K]) are spread over the lanes of the group, one register each, instead of every lane
// carrying all of them: lane 0..2 -> losses, lane 3..3+K-1 -> db[k] (stat1); lane k -> count of env k (stat2).
struct Running {
    float stat1, stat2;
    float sq, ab;         // sum x^2 / |x| over the gathered rows (this lane's dims)
};

template <int VEC, int NV, int KT>
__device__ __forceinline__ void fused_range(const UserPassArgs& a, const LossCfg& cfg, const float* __restrict__ sE,
                                            const float* __restrict__ sW, float* __restrict__ myDE,
                                            const float* __restrict__ sB, const Row<VEC, NV>& ra, const Row<VEC, NV>& rue,
                                            int beg, int end, int lane, unsigned gmask, float (&acc0)[NV * VEC],
                                            float (&Q)[KT][NV * VEC], float (&acc_env)[NV * VEC], Running& st) {
    const int D = a.side.D, K = a.side.K;
    const int32_t* __restrict__ perm = a.side.plan.perm;
    const int32_t* __restrict__ partner = a.side.plan.partner;
    for (int k = beg; k < end; ++k) {
        const int n = perm[k];
        const int it = partner[k];
        const int e = (int)a.envs[n];
        const float y = a.scores[n];
        const float w = (a.weights != nullptr) ? a.weights[n] : 1.f;
        Row<VEC, NV> rc, rie;
        load_row<VEC, NV>(rc, a.side.partner_inv, it, D, lane);
        load_row<VEC, NV>(rie, a.side.partner_env, it, D, lane);
        float z1 = 0.f, z2 = 0.f, sq = 0.f, ab = 0.f;
        float lg[KT], t[NV * VEC], ee[NV * VEC];
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) lg[kk] = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
            if (d0 < D) {
                ldv<VEC>(sE + e * D + d0, &ee[j * VEC]);
                float p[VEC], wk[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int x = j * VEC + v;
                    p[v] = ra.x[x] * rc.x[x];
                    t[x] = rue.x[x] * rie.x[x];
                    z1 += p[v];
                    z2 += t[x] * ee[x];
                    sq += rc.x[x] * rc.x[x] + rie.x[x] * rie.x[x];
                    ab += fabsf(rc.x[x]) + fabsf(rie.x[x]);
                }
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) {
                    if (kk < K) {
                        ldv<VEC>(sW + kk * D + d0, wk);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) lg[kk] += wk[v] * p[v];
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { t[j * VEC + v] = 0.f; ee[j * VEC + v] = 0.f; }
            }
        }
        z1 = group_sum(z1, gmask);
        z2 = group_sum(z2, gmask);
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) lg[kk] = (kk < K) ? group_sum(lg[kk], gmask) + sB[kk] : -INFINITY;

        float g_z1, g_z2, gl[KT], lw[3];
        loss_grads<KT>(cfg, z1, z2, lg, y, w, e, g_z1, g_z2, gl, lw);
        st.sq += sq;
        st.ab += ab;
        {
            float v1 = (lane == 0) ? lw[0] : ((lane == 1) ? lw[1] : ((lane == 2) ? lw[2] : 0.f));
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) v1 = (lane == 3 + kk) ? gl[kk] : v1;
            st.stat1 += v1;
            st.stat2 += (lane == e) ? 1.f : 0.f;
        }
        // user-side gradient pieces and dE[e] += g_z2 * ue*ie (this group's shared-memory slice)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int d0 = dim_of<VEC>(lane, j);
            if (d0 < D) {
                float de[VEC];
                ldv<VEC>(myDE + e * D + d0, de);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const int x = j * VEC + v;
                    acc0[x] += g_z1 * rc.x[x];
                    acc_env[x] += g_z2 * rie.x[x] * ee[x];
                    de[v] += g_z2 * t[x];
#pragma unroll
                    for (int kk = 0; kk < KT; ++kk) Q[kk][x] += gl[kk] * rc.x[x];
                }
                stv<VEC>(myDE + e * D + d0, de);
            }
        }
        if (lane == 0) {   // g-pack for the item pass: g_z1, g_z2, env, -alpha * g_logits
            float* gp = a.gpack_out + (int64_t)n * a.side.GS;
            float out[12];
            out[0] = g_z1;
            out[1] = g_z2;
            out[2] = __int_as_float(e);
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) out[3 + kk] = (kk < KT) ? a.neg_alpha * gl[kk < KT ? kk : 0] : 0.f;
            *reinterpret_cast<float4*>(gp) = make_float4(out[0], out[1], out[2], out[3]);
            *reinterpret_cast<float4*>(gp + 4) = make_float4(out[4], out[5], out[6], out[7]);
            if (KT > 5 && a.side.GS > 8)
                *reinterpret_cast<float4*>(gp + 8) = make_float4(out[8], out[9], out[10], out[11]);
        }
    }
}

// gi = acc0 + (-alpha) sum_k W_k (.) Q_k ;  dW_k += a (.) Q_k  (this group's shared-memory slice)
template <int VEC, int NV, int KT>
__device__ __forceinline__ void finish_range(const UserPassArgs& a, const float* __restrict__ sW,
                                             float* __restrict__ myDW, const Row<VEC, NV>& ra, int lane,
                                             const float (&acc0)[NV * VEC], const float (&Q)[KT][NV * VEC],
                                             Row<VEC, NV>& gi) {
    const int D = a.side.D, K = a.side.K;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int d0 = dim_of<VEC>(lane, j);
        if (d0 < D) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) gi.x[j * VEC + v] = acc0[j * VEC + v];
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                if (kk < K) {
                    float wk[VEC], dw[VEC];
                    ldv<VEC>(sW + kk * D + d0, wk);
                    ldv<VEC>(myDW + kk * D + d0, dw);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const int x = j * VEC + v;
                        gi.x[x] += a.neg_alpha * wk[v] * Q[kk][x];
                        dw[v] += ra.x[x] * Q[kk][x];
                    }
                    stv<VEC>(myDW + kk * D + d0, dw);
                }
            }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) gi.x[j * VEC + v] = 0.f;
        }
    }
}

// Lazy mode: replay the zero-gradient Adam steps last+1 .. upto (inclusive) of one user row in registers.
template <int VEC, int NV>
__device__ __forceinline__ void replay_steps(const BwdSideArgs& sd, int last, int upto, Row<VEC, NV>& th_i,
                                             Row<VEC, NV>& th_e, Row<VEC, NV>& m_i, Row<VEC, NV>& m_e,
                                             Row<VEC, NV>& v_i, Row<VEC, NV>& v_e) {
    for (int j = last + 1; j <= upto; ++j) {
        const float2 sc = sd.sched[j];
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) {
            adam_zero_step(th_i.x[x], m_i.x[x], v_i.x[x], sd.adam, sc.x, sc.y);
            adam_zero_step(th_e.x[x], m_e.x[x], v_e.x[x], sd.adam, sc.x, sc.y);
        }
    }
}

struct Smem {
    float *sE, *sW, *sRed, *sDE, *sDW, *sB;
};

__device__ __forceinline__ Smem carve_smem(float* smem, int KD) {
    Smem s;
    s.sE = smem;
    s.sW = smem + KD;
    s.sRed = smem + 2 * KD;
    s.sDE = smem + 4 * KD;
    s.sDW = smem + (4 + GROUPS_PER_BLOCK) * KD;
    s.sB = smem + (4 + 2 * GROUPS_PER_BLOCK) * KD;
    return s;
}

__device__ __forceinline__ void stage(const UserPassArgs& a, const Smem& s, int KD, Running& st) {
    for (int t = threadIdx.x; t < KD; t += BLOCK) { s.sE[t] = a.side.E[t]; s.sW[t] = a.side.W[t]; }
    for (int t = threadIdx.x; t < 2 * GROUPS_PER_BLOCK * KD; t += BLOCK) s.sDE[t] = 0.f;   // sDE and sDW
    if (threadIdx.x < INVPREF_MAX_ENVS) s.sB[threadIdx.x] = ((int)threadIdx.x < a.side.K) ? a.b[threadIdx.x] : 0.f;
    __syncthreads();
    st.stat1 = st.stat2 = st.sq = st.ab = 0.f;
}

// CTA reduction in a fixed order, then this CTA's partial vector
__device__ __forceinline__ void write_partials(const UserPassArgs& a, const Smem& s, int KD, const Running& st,
                                               int cta) {
    __syncthreads();
    for (int t = threadIdx.x; t < KD; t += BLOCK) {
        float w = 0.f, e = 0.f;
        for (int g = 0; g < GROUPS_PER_BLOCK; ++g) { w += s.sDW[g * KD + t]; e += s.sDE[g * KD + t]; }
        s.sRed[t] = w;
        s.sRed[KD + t] = e;
    }
    __shared__ float sScal[BLOCK / 32][24];
    const int glane = threadIdx.x & (GROUP - 1);
    float sc[24];
    sc[0] = (glane == 0) ? st.stat1 : 0.f;
    sc[1] = (glane == 1) ? st.stat1 : 0.f;
    sc[2] = (glane == 2) ? st.stat1 : 0.f;
    sc[3] = st.sq; sc[4] = st.ab; sc[5] = sc[6] = sc[7] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[8 + k] = (glane == 3 + k) ? st.stat1 : 0.f;
        sc[16 + k] = (glane == k) ? st.stat2 : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 24; ++q) sc[q] = warp_sum(sc[q]);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int q = 0; q < 24; ++q) sScal[warp][q] = sc[q];
    }
    __syncthreads();
    float* out = a.partials + (int64_t)cta * a.P;
    if (threadIdx.x < 24) {
        float v = 0.f;
        for (int w = 0; w < BLOCK / 32; ++w) v += sScal[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
    for (int t = threadIdx.x; t < 2 * KD; t += BLOCK) out[P_DW + t] = s.sRed[t];
}

template <int VEC, int NV, int KT, bool LAZY>
__global__ void __launch_bounds__(BLOCK, 2) upass_chunks_kernel(UserPassArgs a, int cta_offset) {
    extern __shared__ float smem[];
    const int D = a.side.D, KD = a.side.K * a.side.D;
    const Smem s = carve_smem(smem, KD);
    Running st;
    stage(a, s, KD, st);
    const int lane = threadIdx.x & (GROUP - 1);
    const unsigned gmask = group_mask();
    float* myDE = s.sDE + (threadIdx.x >> 4) * KD;
    float* myDW = s.sDW + (threadIdx.x >> 4) * KD;
    const LossCfg cfg = {a.side.K, a.implicit, a.use_class_rw, a.use_rec_rw, a.c_inv, a.c_ea, a.c_env, a.invB};
    const int n_chunks = a.side.plan.counters[1];
    const int ngroups = gridDim.x * GROUPS_PER_BLOCK;
    for (int c = blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x >> 4); c < n_chunks; c += ngroups) {
        const int4 desc = reinterpret_cast<const int4*>(a.side.plan.chunk_desc)[c];
        const int64_t row = a.side.plan.seg_row[desc.x];
        Row<VEC, NV> ra, rue, gi, ge;
        load_row<VEC, NV>(ra, a.side.own_inv_in, row, D, lane);
        load_row<VEC, NV>(rue, a.side.own_env_in, row, D, lane);
        if (LAZY) {   // bring the row up to step-1 in registers (the rows kernel does the same and stores it)
            Row<VEC, NV> m_i, m_e, v_i, v_e;
            load_row<VEC, NV>(m_i, a.side.m_inv, row, D, lane);
            load_row<VEC, NV>(m_e, a.side.m_env, row, D, lane);
            load_row<VEC, NV>(v_i, a.side.v_inv, row, D, lane);
            load_row<VEC, NV>(v_e, a.side.v_env, row, D, lane);
            replay_steps<VEC, NV>(a.side, a.side.last_step[row], a.side.step - 1, ra, rue, m_i, m_e, v_i, v_e);
        }
        float acc0[NV * VEC], Q[KT][NV * VEC];
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) {
            acc0[x] = 0.f; ge.x[x] = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) Q[k][x] = 0.f;
        }
        fused_range<VEC, NV, KT>(a, cfg, s.sE, s.sW, myDE, s.sB, ra, rue, desc.y, desc.z, lane, gmask, acc0, Q, ge.x, st);
        finish_range<VEC, NV, KT>(a, s.sW, myDW, ra, lane, acc0, Q, gi);
        store_row<VEC, NV>(gi, a.side.chunk_part, (int64_t)c * 2, D, lane);
        store_row<VEC, NV>(ge, a.side.chunk_part, (int64_t)c * 2 + 1, D, lane);
    }
    write_partials(a, s, KD, st, cta_offset + blockIdx.x);
}

template <int VEC, int NV, int KT, int EPI, bool LAZY>
__global__ void __launch_bounds__(BLOCK, 2) upass_rows_kernel(UserPassArgs a) {
    extern __shared__ float smem[];
    const int D = a.side.D, KD = a.side.K * a.side.D;
    const Smem s = carve_smem(smem, KD);
    Running st;
    stage(a, s, KD, st);
    const int lane = threadIdx.x & (GROUP - 1);
    const unsigned gmask = group_mask();
    float* myDE = s.sDE + (threadIdx.x >> 4) * KD;
    float* myDW = s.sDW + (threadIdx.x >> 4) * KD;
    const LossCfg cfg = {a.side.K, a.implicit, a.use_class_rw, a.use_rec_rw, a.c_inv, a.c_ea, a.c_env, a.invB};
    const int n_seg = a.side.plan.counters[0];
    const int ngroups = gridDim.x * GROUPS_PER_BLOCK;
    const int s0 = blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x >> 4);
    // Software pipeline over this group's segments s0, s0+ng, ...: while segment j is processed, the rows
    // that segment j+1 touches first (its two user rows, their Adam state, the item rows and the per-sample
    // scalars of its first interaction) are requested into L2.  Every address a prefetch needs comes from a
    // register that was loaded one iteration earlier, so the (in-order) warp never waits on it.
    const int32_t* __restrict__ seg_row = a.side.plan.seg_row;
    const int32_t* __restrict__ seg_off = a.side.plan.seg_off;
    int row1 = 0, beg1 = 0, pid1 = 0, n1 = 0, row2 = 0, beg2 = 0;
    if (s0 + ngroups < n_seg) {
        row1 = seg_row[s0 + ngroups]; beg1 = seg_off[s0 + ngroups];
        pid1 = a.side.plan.partner[beg1]; n1 = a.side.plan.perm[beg1];
    }
    if (s0 + 2 * ngroups < n_seg) { row2 = seg_row[s0 + 2 * ngroups]; beg2 = seg_off[s0 + 2 * ngroups]; }
    for (int sgm = s0; sgm < n_seg; sgm += ngroups) {
        if (sgm + ngroups < n_seg) {
            prefetch_row(a.side.own_inv_in, row1, D, lane);
            prefetch_row(a.side.own_env_in, row1, D, lane);
            prefetch_row(a.side.partner_inv, pid1, D, lane);
            prefetch_row(a.side.partner_env, pid1, D, lane);
            if (EPI == EPI_ADAM) {
                prefetch_row(a.side.m_inv, row1, D, lane);
                prefetch_row(a.side.m_env, row1, D, lane);
                prefetch_row(a.side.v_inv, row1, D, lane);
                prefetch_row(a.side.v_env, row1, D, lane);
            }
            if (lane == 8) prefetch_l2(a.envs + n1);
            if (lane == 9) prefetch_l2(a.scores + n1);
            if (lane == 10 && a.weights != nullptr) prefetch_l2(a.weights + n1);
        }
        int pid2 = 0, n2 = 0, row3 = 0, beg3 = 0;
        if (sgm + 2 * ngroups < n_seg) { pid2 = a.side.plan.partner[beg2]; n2 = a.side.plan.perm[beg2]; }
        if (sgm + 3 * ngroups < n_seg) { row3 = seg_row[sgm + 3 * ngroups]; beg3 = seg_off[sgm + 3 * ngroups]; }
        const int64_t row = seg_row[sgm];
        const int beg = seg_off[sgm], end = seg_off[sgm + 1];
        const int c0 = a.side.plan.seg_chunk[sgm], c1 = a.side.plan.seg_chunk[sgm + 1];
        Row<VEC, NV> ra, rue, gi, ge;
        Row<VEC, NV> m_i, m_e, v_i, v_e;
        load_row<VEC, NV>(ra, a.side.own_inv_in, row, D, lane);
        load_row<VEC, NV>(rue, a.side.own_env_in, row, D, lane);
        if (LAZY) {
            // the row may be several steps behind: replay the skipped zero-gradient Adam steps in registers,
            // then stash the caught-up row (what every reader of this step must see) for the item pass
            load_row<VEC, NV, true>(m_i, a.side.m_inv, row, D, lane);
            load_row<VEC, NV, true>(m_e, a.side.m_env, row, D, lane);
            load_row<VEC, NV, true>(v_i, a.side.v_inv, row, D, lane);
            load_row<VEC, NV, true>(v_e, a.side.v_env, row, D, lane);
            replay_steps<VEC, NV>(a.side, a.side.last_step[row], a.side.step - 1, ra, rue, m_i, m_e, v_i, v_e);
            store_row<VEC, NV>(ra, a.side.stash, (int64_t)sgm * 2, D, lane);
            store_row<VEC, NV>(rue, a.side.stash, (int64_t)sgm * 2 + 1, D, lane);
        }
        if (c1 > c0) {   // long segment: its forward + reduction ran in upass_chunks_kernel
#pragma unroll
            for (int x = 0; x < NV * VEC; ++x) { gi.x[x] = 0.f; ge.x[x] = 0.f; }
            for (int c = c0; c < c1; ++c) {
                Row<VEC, NV> pi, pe;
                load_row<VEC, NV>(pi, a.side.chunk_part, (int64_t)c * 2, D, lane);
                load_row<VEC, NV>(pe, a.side.chunk_part, (int64_t)c * 2 + 1, D, lane);
#pragma unroll
                for (int x = 0; x < NV * VEC; ++x) { gi.x[x] += pi.x[x]; ge.x[x] += pe.x[x]; }
            }
        } else {
            float acc0[NV * VEC], Q[KT][NV * VEC];
#pragma unroll
            for (int x = 0; x < NV * VEC; ++x) {
                acc0[x] = 0.f; ge.x[x] = 0.f;
#pragma unroll
                for (int k = 0; k < KT; ++k) Q[k][x] = 0.f;
            }
            fused_range<VEC, NV, KT>(a, cfg, s.sE, s.sW, myDE, s.sB, ra, rue, beg, end, lane, gmask, acc0, Q, ge.x, st);
            finish_range<VEC, NV, KT>(a, s.sW, myDW, ra, lane, acc0, Q, gi);
        }
        // the user rows' own L1/L2 terms (models.py:469-482): every occurrence in the batch counts
        const float cnt = (float)(end - beg);
        float sq = 0.f, ab = 0.f;
#pragma unroll
        for (int x = 0; x < NV * VEC; ++x) {
            sq += ra.x[x] * ra.x[x] + rue.x[x] * rue.x[x];
            ab += fabsf(ra.x[x]) + fabsf(rue.x[x]);
            gi.x[x] += cnt * (a.side.reg2 * ra.x[x] + mul_sign(a.side.reg1, ra.x[x]));
            ge.x[x] += cnt * (a.side.reg2 * rue.x[x] + mul_sign(a.side.reg1, rue.x[x]));
        }
        st.sq += cnt * sq;
        st.ab += cnt * ab;
        if (a.side.grad_inv != nullptr) {
            store_row<VEC, NV>(gi, a.side.grad_inv, row, D, lane);
            store_row<VEC, NV>(ge, a.side.grad_env, row, D, lane);
        }
        if (EPI == EPI_ADAM) {
            if (!LAZY) {
                load_row<VEC, NV, true>(m_i, a.side.m_inv, row, D, lane);
                load_row<VEC, NV, true>(m_e, a.side.m_env, row, D, lane);
                load_row<VEC, NV, true>(v_i, a.side.v_inv, row, D, lane);
                load_row<VEC, NV, true>(v_e, a.side.v_env, row, D, lane);
            }
#pragma unroll
            for (int x = 0; x < NV * VEC; ++x) {
                adam_update(ra.x[x], m_i.x[x], v_i.x[x], gi.x[x], a.side.adam);
                adam_update(rue.x[x], m_e.x[x], v_e.x[x], ge.x[x], a.side.adam);
            }
            store_row<VEC, NV>(ra, a.side.own_inv_out, row, D, lane);
            store_row<VEC, NV>(rue, a.side.own_env_out, row, D, lane);
            store_row<VEC, NV, true>(m_i, a.side.m_inv, row, D, lane);
            store_row<VEC, NV, true>(m_e, a.side.m_env, row, D, lane);
            store_row<VEC, NV, true>(v_i, a.side.v_inv, row, D, lane);
            store_row<VEC, NV, true>(v_e, a.side.v_env, row, D, lane);
            if (LAZY && lane == 0) a.side.last_step[row] = a.side.step;
        }
        row1 = row2; beg1 = beg2; pid1 = pid2; n1 = n2;
        row2 = row3; beg2 = beg3;
    }
    write_partials(a, s, KD, st, blockIdx.x);
}

inline size_t upass_smem(const Geometry& g) {
    return ((size_t)(4 + 2 * GROUPS_PER_BLOCK) * g.K * g.D + INVPREF_MAX_ENVS) * sizeof(float);
}

}  // namespace

bool upass_supported(const Geometry& g) { return upass_smem(g) <= 96 * 1024; }

int upass_rows_grid(int64_t max_seg) {
    int64_t need = (max_seg + GROUPS_PER_BLOCK - 1) / GROUPS_PER_BLOCK;
    if (need < 1) need = 1;
    return (int)(need < 148 * 4 ? need : 148 * 4);
}

int launch_upass_chunks(const Geometry& g, const UserPassArgs& a, int cta_offset, cudaStream_t stream) {
    const size_t smem = upass_smem(g);
    const bool lazy = a.side.last_step != nullptr;
#define LAUNCH(KERNEL)                                                                                           \
    do {                                                                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        KERNEL<<<UPASS_CHUNK_CTAS, BLOCK, smem, stream>>>(a, cta_offset);                                        \
    } while (0)
#define CALL(V, N, KT_)                                                                                          \
    do {                                                                                                         \
        if (lazy) LAUNCH((upass_chunks_kernel<V, N, KT_, true>));                                                \
        else LAUNCH((upass_chunks_kernel<V, N, KT_, false>));                                                    \
    } while (0)
    INVPREF_DISPATCH_GEOM(g, CALL);
#undef CALL
#undef LAUNCH
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

int launch_upass_rows(const Geometry& g, const UserPassArgs& a, int epi, int grid, cudaStream_t stream) {
    const size_t smem = upass_smem(g);
    const bool lazy = a.side.last_step != nullptr;
    if (lazy && epi != EPI_ADAM) return INVPREF_ERR_BAD_ARG;
#define LAUNCH(KERNEL)                                                                                           \
    do {                                                                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        KERNEL<<<grid, BLOCK, smem, stream>>>(a);                                                                \
    } while (0)
#define CALL(V, N, KT_)                                                                                          \
    do {                                                                                                         \
        if (lazy) LAUNCH((upass_rows_kernel<V, N, KT_, EPI_ADAM, true>));                                        \
        else if (epi == EPI_ADAM) LAUNCH((upass_rows_kernel<V, N, KT_, EPI_ADAM, false>));                       \
        else LAUNCH((upass_rows_kernel<V, N, KT_, EPI_EXPORT, false>));                                          \
    } while (0)
    INVPREF_DISPATCH_GEOM(g, CALL);
#undef CALL
#undef LAUNCH
    count_launch();
    return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
}

}  // namespace invpref
