// Per-interaction scalar math shared by the forward kernels: scores -> losses -> backward scalars
// (reference models.py:460-462 / 319-321, train.py:797-819, closed-form backward of SURVEY.md §3.4).
#pragma once

#include "common.cuh"

namespace invpref {

struct LossCfg {
    int K, implicit, use_class_rw, use_rec_rw;
    float c_inv, c_ea, c_env, invB;
};

// in : z1 = sum a*c, z2 = sum ue*ie*E[e], lg[k] = classifier logits (k >= K hold -inf), y, w, e
// out: g_z1, g_z2, gl[k] = d loss / d logits[k]; lw[3] = weighted per-sample losses (inv, env-aware, nll)
// DIST (fused user pass: the 16 lanes of a group hold identical inputs): the K softmax exponentials, the single most
// expensive piece (round-1 SASS: 5.6 % of the kernel's instructions), are evaluated ONCE per group -- lane l takes
// k = l mod 2/4/8 -- and shared with shuffles: same function, same argument, same bits as every lane computing all
// K.  The logs that only feed the REPORTED loss values (never a gradient) use the hardware lg2 (abs. error 2^-21.4:
// <= 5e-7 of the O(1) per-sample losses they enter, against the 1e-5 loss tolerance), and the softmax normaliser
// the hardware reciprocal (1 ulp).
template <int KT, bool DIST = false>
__device__ __forceinline__ void loss_grads(const LossCfg& c, float z1, float z2, const float (&lg)[KT], float y,
                                           float w, int e, float& g_z1, float& g_z2, float (&gl)[KT],
                                           float (&lw)[3], int lane = 0, unsigned gmask = 0u) {
    // softmax over K (models.py:208): ex[k] = exp(l_k - max), soft = ex / sum, lse = max + log(sum)
    float mx = lg[0];
#pragma unroll
    for (int k = 1; k < KT; ++k) mx = fmaxf(mx, lg[k]);
    float ex[KT];
    float se = 0.f;
    if (DIST) {
        constexpr int KP = (KT <= 2) ? 2 : ((KT <= 4) ? 4 : 8);
        const int kk = lane & (KP - 1);
        float mine = lg[0];
#pragma unroll
        for (int k = 1; k < KT; ++k) mine = (kk == k) ? lg[k] : mine;
        const float em = (kk < c.K && kk < KT) ? expf(mine - mx) : 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) { ex[k] = __shfl_sync(gmask, em, k, GROUP); se += ex[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < KT; ++k) { ex[k] = (k < c.K) ? expf(lg[k] - mx) : 0.f; se += ex[k]; }
    }
    const float inv_se = DIST ? rcp_approx(se) : 1.f / se;
    const float wr = c.use_rec_rw ? w : 1.f;      // train.py:817-819
    const float wc = c.use_class_rw ? w : 1.f;    // train.py:814-815
    float l_inv, l_ea;
    if (c.implicit) {
        const float s_inv = sigmoidf_(z1), s2 = sigmoidf_(z2);
        const float s_env = s_inv * s2;
        // nn.BCELoss: log clamped at -100; its backward clamps x(1-x) at 1e-12
        if (DIST) {
            l_inv = -(y * fmaxf(__logf(s_inv), -100.f) + (1.f - y) * fmaxf(__logf(1.f - s_inv), -100.f));
            l_ea = -(y * fmaxf(__logf(s_env), -100.f) + (1.f - y) * fmaxf(__logf(1.f - s_env), -100.f));
        } else {
            l_inv = -(y * fmaxf(logf(s_inv), -100.f) + (1.f - y) * fmaxf(logf(1.f - s_inv), -100.f));
            l_ea = -(y * fmaxf(logf(s_env), -100.f) + (1.f - y) * fmaxf(logf(1.f - s_env), -100.f));
        }
        const float r1 = (s_inv - y) / fmaxf(s_inv * (1.f - s_inv), 1e-12f);
        const float r2 = (s_env - y) / fmaxf(s_env * (1.f - s_env), 1e-12f);
        const float g_s1 = wr * c.invB * (c.c_inv * r1 + c.c_ea * r2 * s2);
        const float g_s2 = wr * c.invB * c.c_ea * r2 * s_inv;
        g_z1 = g_s1 * s_inv * (1.f - s_inv);
        g_z2 = g_s2 * s2 * (1.f - s2);
    } else {
        const float d1 = z1 - y, d2 = (z1 + z2) - y;
        l_inv = d1 * d1;
        l_ea = d2 * d2;
        g_z1 = wr * c.invB * 2.f * (c.c_inv * d1 + c.c_ea * d2);
        g_z2 = wr * c.invB * 2.f * c.c_ea * d2;
    }
    const float coef = c.c_env * wc * c.invB;
    float le = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        gl[k] = coef * (ex[k] * inv_se - ((k == e) ? 1.f : 0.f));
        if (k == e) le = lg[k];
    }
    lw[0] = l_inv * wr;
    lw[1] = l_ea * wr;
    lw[2] = ((mx + (DIST ? __logf(se) : logf(se))) - le) * wc;          // -log_softmax[e]
}

}  // namespace invpref
