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

// One interaction, part 1: dot products against the user rows -> logits and scores (unreduced per-lane sums).
template <int VEC, int NV, int KT>
struct Inter {
    float z1, z2, sq, ab;
    float lg[KT], t[NV * VEC], ee[NV * VEC];
};

template <int VEC, int NV, int KT>
__device__ __forceinline__ void inter_dots(const UserPassArgs& a, const float* __restrict__ sE,
                                           const float* __restrict__ sW, const Row<VEC, NV>& ra,
                                           const Row<VEC, NV>& rue, const Row<VEC, NV>& rc, const Row<VEC, NV>& rie,
                                           int e, int lane, Inter<VEC, NV, KT>& q, int D, int K) {
    q.z1 = 0.f; q.z2 = 0.f; q.sq = 0.f; q.ab = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) q.lg[kk] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int d0 = dim_of<VEC>(lane, j);
        if (d0 < D) {
            ldv<VEC>(sE + e * D + d0, &q.ee[j * VEC]);
            float p[VEC], wk[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int x = j * VEC + v;
                p[v] = ra.x[x] * rc.x[x];
                q.t[x] = rue.x[x] * rie.x[x];
                q.z1 += p[v];
                q.z2 += q.t[x] * q.ee[x];
                q.sq += rc.x[x] * rc.x[x] + rie.x[x] * rie.x[x];
                q.ab += fabsf(rc.x[x]) + fabsf(rie.x[x]);
            }
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                if (kk < K) {
                    ldv<VEC>(sW + kk * D + d0, wk);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) q.lg[kk] += wk[v] * p[v];
                }
            }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { q.t[j * VEC + v] = 0.f; q.ee[j * VEC + v] = 0.f; }
        }
    }
}

// One interaction, part 2: group reductions, losses and their backward scalars, the user-side gradient
// pieces, dE, the running sums, and the g-pack of interaction n for the item pass.
template <int VEC, int NV, int KT>
__device__ __forceinline__ void inter_grads(const UserPassArgs& a, const LossCfg& cfg, float* __restrict__ myDE,
                                            const float* __restrict__ sB, const Row<VEC, NV>& rc,
                                            const Row<VEC, NV>& rie, Inter<VEC, NV, KT>& q, int n, int e, float y,
                                            float w, int lane, unsigned gmask, float (&acc0)[NV * VEC],
                                            float (&Q)[KT][NV * VEC], float (&acc_env)[NV * VEC], Running& st,
                                            int D, int K) {
    const float z1 = group_sum(q.z1, gmask);
    const float z2 = group_sum(q.z2, gmask);
    float lg[KT];
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) lg[kk] = (kk < K) ? group_sum(q.lg[kk], gmask) + sB[kk] : -INFINITY;

    float g_z1, g_z2, gl[KT], lw[3];
    loss_grads<KT>(cfg, z1, z2, lg, y, w, e, g_z1, g_z2, gl, lw);
    st.sq += q.sq;
    st.ab += q.ab;
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
                acc_env[x] += g_z2 * rie.x[x] * q.ee[x];
                de[v] += g_z2 * q.t[x];
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) Q[kk][x] += gl[kk] * rc.x[x];
            }
            stv<VEC>(myDE + e * D + d0, de);
        }
    }
    if (lane == 0) {   // g-pack for the item pass: g_z1, g_z2, env, -alpha * g_logits
        const int GS = (K <= 5) ? 8 : 12;   // make_geometry
        float* gp = a.gpack_out + (int64_t)n * GS;
        float out[12];
        out[0] = g_z1;
        out[1] = g_z2;
        out[2] = __int_as_float(e);
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) out[3 + kk] = (kk < KT) ? a.neg_alpha * gl[kk < KT ? kk : 0] : 0.f;
        *reinterpret_cast<float4*>(gp) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(gp + 4) = make_float4(out[4], out[5], out[6], out[7]);
        if (KT > 5 && GS > 8)
            *reinterpret_cast<float4*>(gp + 8) = make_float4(out[8], out[9], out[10], out[11]);
    }
}

// Interactions [beg, end) of one user segment, item rows loaded straight from global memory (chunks kernel and
// the unstaged rows kernel).
template <int VEC, int NV, int KT>
__device__ __forceinline__ void fused_range(const UserPassArgs& a, const LossCfg& cfg, const float* __restrict__ sE,
                                            const float* __restrict__ sW, float* __restrict__ myDE,
                                            const float* __restrict__ sB, const Row<VEC, NV>& ra, const Row<VEC, NV>& rue,
                                            int beg, int end, int lane, unsigned gmask, float (&acc0)[NV * VEC],
                                            float (&Q)[KT][NV * VEC], float (&acc_env)[NV * VEC], Running& st) {
    const int D = a.side.D;
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
        Inter<VEC, NV, KT> q;
        inter_dots<VEC, NV, KT>(a, sE, sW, ra, rue, rc, rie, e, lane, q, D, a.side.K);
        inter_grads<VEC, NV, KT>(a, cfg, myDE, sB, rc, rie, q, n, e, y, w, lane, gmask, acc0, Q, acc_env, st, D, a.side.K);
    }
}

// gi = acc0 + (-alpha) sum_k W_k (.) Q_k ;  dW_k += a (.) Q_k  (this group's shared-memory slice)
template <int VEC, int NV, int KT>
__device__ __forceinline__ void finish_range(const UserPassArgs& a, const float* __restrict__ sW,
                                             float* __restrict__ myDW, const Row<VEC, NV>& ra, int lane,
                                             const float (&acc0)[NV * VEC], const float (&Q)[KT][NV * VEC],
                                             Row<VEC, NV>& gi, int D, int K) {
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
        finish_range<VEC, NV, KT>(a, s.sW, myDW, ra, lane, acc0, Q, gi, D, a.side.K);
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
            finish_range<VEC, NV, KT>(a, s.sW, myDW, ra, lane, acc0, Q, gi, D, a.side.K);
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

// ---------------------------------------------------------------------------------------------------------
// Staged rows kernel (row slices of <= 16 bytes per lane, i.e. D <= 64).
//
// ncu on the register-only kernel above (round 1): 30 % of the warp samples wait on the long scoreboard -- the
// (m, v, theta) rows of the segment behind a chain of dependent index loads, and the item rows of each
// interaction -- with only 16 warps per SM to hide it.  Here EVERYTHING a group will need is copied global ->
// shared with cp.async one step before it is used, so that no register (and no register scoreboard: ptxas made
// the inner loop's first branch wait for every load issued at the top of the segment, 35 % of the samples in
// an intermediate version that kept index loads in registers) is tied up while data is in flight:
//   * group A_{m+1}, requested at the top of segment m: the eight rows of the next segment (theta, m, v of both
//     user tables + the two item rows of its first interaction; every lane copies exactly the slice it later
//     reads), the scalars of its first interaction (env, score, weight) and the row's last_step, and the
//     32-byte descriptor {row, begin, end, perm/partner[begin], perm/partner[begin+1]} of the segment after it;
//   * group G_i, requested while interaction i is computed: the item rows and scalars of interaction i+1 and
//     the indices of interaction i+2.
// Row slices are read back by the lane that copied them; the few words shared by the group (descriptors,
// scalars, indices) after cp.async.wait_group + __syncwarp.  The arithmetic is the register-only kernel's,
// value for value.
constexpr int UP_SLOTS = 8;    // per stage: th_i, th_e, m_i, m_e, v_i, v_e, first item inv, first item env
// per-group metadata (32-bit words)
constexpr int UM_DESC = 0;     // [4][8]  descriptor ring, segment ordinal & 3 (ordinal o is read at the top of
                               //         segments o-1 and o; o+3 is requested during segment o: four slots)
constexpr int UM_SEGS = 32;    // [2][4]  {env, score, weight, last_step} of a segment's first interaction, by stage
constexpr int UM_ITS = 40;     // [2][4]  {env, score, weight, -} of interaction i, slot i & 1
constexpr int UM_ITI = 48;     // [2][2]  {perm, partner} of interaction i, slot i & 1
constexpr int UM_WORDS = 56;

// EXACT: D = 16 * VEC * NV and K = KT are compile-time constants (bounds guards fold away, row offsets are shifts).
template <int VEC, int NV, int KT, int EPI, bool LAZY, bool EXACT>
__global__ void __launch_bounds__(BLOCK, 2) upass_rows_staged_kernel(UserPassArgs a, int long_len) {
    extern __shared__ __align__(16) float smem[];
    const int D = EXACT ? GROUP * VEC * NV : a.side.D, K = EXACT ? KT : a.side.K, KD = K * D;
    const Smem s = carve_smem(smem, KD);
    float* ring = smem + (((4 + 2 * GROUPS_PER_BLOCK) * KD + INVPREF_MAX_ENVS + 3) & ~3);   // [2][8][NV][BLOCK][VEC]
    int32_t* meta = reinterpret_cast<int32_t*>(ring + (size_t)2 * UP_SLOTS * NV * VEC * BLOCK) +
                    (threadIdx.x >> 4) * UM_WORDS;                                          // this group's words
    Running st;
    stage(a, s, KD, st);
    const int lane = threadIdx.x & (GROUP - 1);
    const unsigned gmask = group_mask();
    float* myDE = s.sDE + (threadIdx.x >> 4) * KD;
    float* myDW = s.sDW + (threadIdx.x >> 4) * KD;
    const LossCfg cfg = {K, a.implicit, a.use_class_rw, a.use_rec_rw, a.c_inv, a.c_ea, a.c_env, a.invB};
    const int n_seg = a.side.plan.counters[0];
    const int ng = gridDim.x * GROUPS_PER_BLOCK;
    const int s0 = blockIdx.x * GROUPS_PER_BLOCK + (threadIdx.x >> 4);
    const int32_t* __restrict__ seg_desc = a.side.plan.seg_desc;
    const int32_t* __restrict__ perm = a.side.plan.perm;
    const int32_t* __restrict__ partner = a.side.plan.partner;
    const bool has_w = a.weights != nullptr;

    // descriptor of segment `seg` (ordinal `ord` of this group) -> ring slot ord & 3; lanes 0 and 1, 16 bytes each
    auto request_desc = [&](int ord, int seg) {
        if (seg < n_seg && lane < 2)
            cp_async<16>(smem_addr(meta + UM_DESC + (ord & 3) * 8 + lane * 4), seg_desc + (int64_t)seg * 8 + lane * 4);
    };
    // env / score / weight of interaction n -> dst[0..2] (lanes 8, 9, 10)
    auto request_scalars = [&](int32_t* dst, int n) {
        if (lane == 8) cp_async<4>(smem_addr(dst + 0), reinterpret_cast<const int32_t*>(a.envs + n));   // low word
        if (lane == 9) cp_async<4>(smem_addr(dst + 1), a.scores + n);
        if (lane == 10) {
            if (has_w) cp_async<4>(smem_addr(dst + 2), a.weights + n);
            else dst[2] = __float_as_int(1.f);
        }
    };
    // rows + first-interaction scalars + last_step of the segment described by dsc[] (shared memory) -> stage stg
    auto request_segment = [&](int stg, const int32_t* dsc) {
        const int row = dsc[0], n0 = dsc[3], p0 = dsc[4];
        stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 0, a.side.own_inv_in, row, D, lane);
        stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 1, a.side.own_env_in, row, D, lane);
        if (EPI == EPI_ADAM) {
            stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 2, a.side.m_inv, row, D, lane);
            stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 3, a.side.m_env, row, D, lane);
            stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 4, a.side.v_inv, row, D, lane);
            stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 5, a.side.v_env, row, D, lane);
        }
        stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 6, a.side.partner_inv, p0, D, lane);
        stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 7, a.side.partner_env, p0, D, lane);
        request_scalars(meta + UM_SEGS + stg * 4, n0);
        if (LAZY && lane == 11) cp_async<4>(smem_addr(meta + UM_SEGS + stg * 4 + 3), a.side.last_step + row);
    };

    // prologue: descriptors of the first two segments (the only exposed latency), then group A_0
    request_desc(0, s0);
    request_desc(1, s0 + ng);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp(gmask);
    if (s0 < n_seg) request_segment(0, meta + UM_DESC);
    request_desc(2, s0 + 2 * ng);
    cp_async_commit();

    int ord = 0;
    for (int sgm = s0; sgm < n_seg; sgm += ng, ++ord) {
        const int stg = ord & 1;
        cp_async_wait<0>();      // A_ord: this segment's rows and scalars, the next segment's descriptor
        __syncwarp(gmask);
        if (sgm + ng < n_seg) request_segment(stg ^ 1, meta + UM_DESC + ((ord + 1) & 3) * 8);
        request_desc(ord + 3, sgm + 3 * ng);     // its slot held this group's previous segment
        cp_async_commit();       // A_{ord+1}

        const int4 dA = *reinterpret_cast<const int4*>(meta + UM_DESC + (ord & 3) * 8);       // row, begin, end, n0
        const int4 dB = *reinterpret_cast<const int4*>(meta + UM_DESC + (ord & 3) * 8 + 4);   // p0, n1, p1, -
        const int4 sS = *reinterpret_cast<const int4*>(meta + UM_SEGS + stg * 4);
        const int64_t row = dA.x;
        const int beg = dA.y, end = dA.z;
        int n = dA.w, n_a = dB.y, it_a = dB.z;     // this interaction's batch position; indices of the next one
        int e = sS.x;
        float y = __int_as_float(sS.y), w = __int_as_float(sS.z);
        const int last = sS.w;
        const bool is_long = end - beg > long_len;

        Row<VEC, NV> ra, rue, gi, ge;
        Row<VEC, NV> m_i, m_e, v_i, v_e;
        read_staged_row<VEC, NV>(ra, ring, stg * UP_SLOTS + 0, D, lane);
        read_staged_row<VEC, NV>(rue, ring, stg * UP_SLOTS + 1, D, lane);
        if (LAZY) {
            // the row may be several steps behind: replay the skipped zero-gradient Adam steps in registers,
            // then stash the caught-up row (what every reader of this step must see) for the item pass
            read_staged_row<VEC, NV>(m_i, ring, stg * UP_SLOTS + 2, D, lane);
            read_staged_row<VEC, NV>(m_e, ring, stg * UP_SLOTS + 3, D, lane);
            read_staged_row<VEC, NV>(v_i, ring, stg * UP_SLOTS + 4, D, lane);
            read_staged_row<VEC, NV>(v_e, ring, stg * UP_SLOTS + 5, D, lane);
            replay_steps<VEC, NV>(a.side, last, a.side.step - 1, ra, rue, m_i, m_e, v_i, v_e);
            store_row<VEC, NV>(ra, a.side.stash, (int64_t)sgm * 2, D, lane);
            store_row<VEC, NV>(rue, a.side.stash, (int64_t)sgm * 2 + 1, D, lane);
            // the caught-up moments go back to their (own) slots: no registers held over the interactions
            write_staged_row<VEC, NV>(m_i, ring, stg * UP_SLOTS + 2, D, lane);
            write_staged_row<VEC, NV>(m_e, ring, stg * UP_SLOTS + 3, D, lane);
            write_staged_row<VEC, NV>(v_i, ring, stg * UP_SLOTS + 4, D, lane);
            write_staged_row<VEC, NV>(v_e, ring, stg * UP_SLOTS + 5, D, lane);
        }
        if (is_long) {   // long segment: its forward + reduction ran in upass_chunks_kernel
            const int c0 = a.side.plan.seg_chunk[sgm], c1 = a.side.plan.seg_chunk[sgm + 1];
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
            for (int k = beg; k < end; ++k) {
                const int par = (k - beg) & 1;
                if (k > beg) {
                    cp_async_wait<0>();      // G_{i-1}: this interaction's rows and scalars, the next one's indices
                    __syncwarp(gmask);
                    const int4 sI = *reinterpret_cast<const int4*>(meta + UM_ITS + par * 4);
                    e = sI.x; y = __int_as_float(sI.y); w = __int_as_float(sI.z);
                    if (k + 1 < end) {
                        const int2 ix = *reinterpret_cast<const int2*>(meta + UM_ITI + (par ^ 1) * 2);
                        n_a = ix.x; it_a = ix.y;
                    }
                }
                Row<VEC, NV> rc, rie;
                read_staged_row<VEC, NV>(rc, ring, stg * UP_SLOTS + 6, D, lane);
                read_staged_row<VEC, NV>(rie, ring, stg * UP_SLOTS + 7, D, lane);
                Inter<VEC, NV, KT> q;
                inter_dots<VEC, NV, KT>(a, s.sE, s.sW, ra, rue, rc, rie, e, lane, q, D, K);
                if (k + 1 < end) {
                    // the item slots are in registers (the sums below depend on every element): request
                    // interaction i+1's rows into them, its scalars, and the indices of interaction i+2
                    asm volatile("" ::"f"(q.z1), "f"(q.z2) : "memory");
                    stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 6, a.side.partner_inv, it_a, D, lane);
                    stage_row_async<VEC, NV>(ring, stg * UP_SLOTS + 7, a.side.partner_env, it_a, D, lane);
                    request_scalars(meta + UM_ITS + (par ^ 1) * 4, n_a);
                    if (k + 2 < end) {
                        if (lane == 12) cp_async<4>(smem_addr(meta + UM_ITI + par * 2 + 0), perm + k + 2);
                        if (lane == 13) cp_async<4>(smem_addr(meta + UM_ITI + par * 2 + 1), partner + k + 2);
                    }
                    cp_async_commit();       // G_i
                }
                inter_grads<VEC, NV, KT>(a, cfg, myDE, s.sB, rc, rie, q, n, e, y, w, lane, gmask, acc0, Q, ge.x, st, D, K);
                n = n_a;
            }
            finish_range<VEC, NV, KT>(a, s.sW, myDW, ra, lane, acc0, Q, gi, D, K);
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
            read_staged_row<VEC, NV>(m_i, ring, stg * UP_SLOTS + 2, D, lane);
            read_staged_row<VEC, NV>(m_e, ring, stg * UP_SLOTS + 3, D, lane);
            read_staged_row<VEC, NV>(v_i, ring, stg * UP_SLOTS + 4, D, lane);
            read_staged_row<VEC, NV>(v_e, ring, stg * UP_SLOTS + 5, D, lane);
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
    }
    cp_async_wait<0>();
    write_partials(a, s, KD, st, blockIdx.x);
}

inline size_t upass_ring_bytes(const Geometry& g) {   // staged rows + per-group metadata
    return (size_t)2 * UP_SLOTS * g.NV * g.VEC * BLOCK * sizeof(float) + (size_t)GROUPS_PER_BLOCK * UM_WORDS * 4;
}
inline bool upass_staged(const Geometry& g) {
    static const bool enabled = [] {
        const char* e = getenv("INVPREF_STAGED");   // INVPREF_STAGED=0: register-only rows kernel (A/B runs)
        return !(e && e[0] == '0');
    }();
    return enabled && g.NV * g.VEC <= 4;
}

inline size_t upass_smem(const Geometry& g) {
    return ((size_t)(4 + 2 * GROUPS_PER_BLOCK) * g.K * g.D + INVPREF_MAX_ENVS) * sizeof(float);
}

}  // namespace

bool upass_supported(const Geometry& g) { return upass_smem(g) <= 96 * 1024; }

static bool use_staged(const Geometry& g) {
    return upass_staged(g) && ((upass_smem(g) + 15) & ~(size_t)15) + upass_ring_bytes(g) <= 227 * 1024;
}

int upass_rows_grid(const Geometry& g, int64_t max_seg) {
    int64_t need = (max_seg + GROUPS_PER_BLOCK - 1) / GROUPS_PER_BLOCK;
    if (need < 1) need = 1;
    // staged kernel: two CTAs are resident per SM (registers, shared memory) -> one wave
    const int64_t cap = use_staged(g) ? 148 * 2 : 148 * 4;
    return (int)(need < cap ? need : cap);
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
    const bool lazy = a.side.last_step != nullptr;
    if (lazy && epi != EPI_ADAM) return INVPREF_ERR_BAD_ARG;
    if (use_staged(g)) {
        const size_t smem = ((upass_smem(g) + 15) & ~(size_t)15) + upass_ring_bytes(g);
        const int long_len = 2 * chunk_for(a.side.plan.B);   // plan.cu: segments longer than this are chunked
#define LAUNCH(KERNEL)                                                                                           \
    do {                                                                                                         \
        cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
        KERNEL<<<grid, BLOCK, smem, stream>>>(a, long_len);                                                      \
    } while (0)
#define CALL_X(V, N, KT_, X)                                                                                     \
    do {                                                                                                         \
        if (lazy) LAUNCH((upass_rows_staged_kernel<V, N, KT_, EPI_ADAM, true, X>));                              \
        else if (epi == EPI_ADAM) LAUNCH((upass_rows_staged_kernel<V, N, KT_, EPI_ADAM, false, X>));             \
        else LAUNCH((upass_rows_staged_kernel<V, N, KT_, EPI_EXPORT, false, X>));                                \
    } while (0)
#define CALL(V, N, KT_) CALL_X(V, N, KT_, false)
#define CALL_EXACT(V, N, KT_) CALL_X(V, N, KT_, true)
        const int _k = g.KT;
        if (g.VEC == 4 && g.D == GROUP * 4 && g.K == g.KT) { INVPREF_DISPATCH_K(4, 1, _k, CALL_EXACT); }
        else if (g.VEC == 4) { INVPREF_DISPATCH_K(4, 1, _k, CALL); }
        else if (g.VEC == 2 && g.NV == 1) { INVPREF_DISPATCH_K(2, 1, _k, CALL); }
        else if (g.VEC == 2) { INVPREF_DISPATCH_K(2, 2, _k, CALL); }
        else { INVPREF_DISPATCH_K(1, 4, _k, CALL); }
#undef CALL
#undef CALL_EXACT
#undef CALL_X
#undef LAUNCH
        count_launch();
        return cudaGetLastError() == cudaSuccess ? INVPREF_OK : INVPREF_ERR_CUDA;
    }
    const size_t smem = upass_smem(g);
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
