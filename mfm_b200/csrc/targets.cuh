// Target log-densities on device.  Mirrors include/mfm_b200.h::mfm_target_t.
//
// Small targets (Gaussian mixture d=2, phi-four, independent Gaussian) are evaluated by ONE WARP
// PER CHAIN from a shared-memory copy of the position.  The pines target (dense 1600x1600 prior)
// is evaluated as a GEMM against K^-1 (see pines.cu).
// Reference: distributions.py:58-67 (GMM), :131-160 (PhiFour), :89-90 (IndepGaussian).
#pragma once
#include "common.cuh"
#include "../../include/mfm_b200.h"

namespace mfm {

// value + gradient of (loglik, logprior) at xs[0..d) (shared memory, warp-visible).
// Writes the *untempered* loglik gradient into gs[0..d) (prior gradient is 0 for these targets)
// and returns loglik in all lanes.  Caller must __syncwarp() before reading gs.
__device__ __forceinline__ float small_target_loglik_grad(const mfm_target_t& T, const float* xs, float* gs, int lane) {
    const int d = T.dim;
    if (T.kind == MFM_TARGET_GMM) {
        // log sum_k w_k prod_j pdf(x_j; m_kj, s_kj), probability domain as coded (distributions.py:58-61)
        float S = 0.0f, gx[2] = {0.0f, 0.0f};
        const float x0 = xs[0], x1 = xs[1];
        for (int k = lane; k < T.n_modes; k += 32) {
            const float m0 = T.modes[2 * k], m1 = T.modes[2 * k + 1];
            const float s0 = T.stds[2 * k], s1 = T.stds[2 * k + 1];
            const float v0 = s0 * s0, v1 = s1 * s1;
            const float d0 = x0 - m0, d1 = x1 - m1;
            const float lp0 = -(logf(6.28318530717958647692f * v0) + d0 * d0 / v0) / 2.0f;
            const float lp1 = -(logf(6.28318530717958647692f * v1) + d1 * d1 / v1) / 2.0f;
            const float pk = T.weights[k] * (expf(lp0) * expf(lp1));
            S += pk;
            gx[0] += pk * (-d0 / v0);
            gx[1] += pk * (-d1 / v1);
        }
        S = warp_sum(S); gx[0] = warp_sum(gx[0]); gx[1] = warp_sum(gx[1]);
        if (lane == 0) { gs[0] = gx[0] / S; gs[1] = gx[1] / S; }
        return logf(S);
    } else if (T.kind == MFM_TARGET_PHI4) {
        const float coef = T.phi_a * (float)d;
        float U = 0.0f, V = 0.0f;
        for (int i = lane; i < d; i += 32) {
            const float xi = xs[i];
            const float xl = i > 0 ? xs[i - 1] : 0.0f;
            const float xr = i + 1 < d ? xs[i + 1] : 0.0f;
            const float df = xi - xl;
            U += df * df;
            if (i == d - 1) U += xi * xi;            // last padded difference (0 - x_{d-1})^2
            const float q = 1.0f - xi * xi;
            V += q * q;
            gs[i] = -T.phi_beta * (coef * (2.0f * xi - xl - xr) - xi * q / coef);
        }
        U = warp_sum(U); V = warp_sum(V);
        return -T.phi_beta * (U / 2.0f * coef + V / 4.0f / coef);
    } else {  // MFM_TARGET_GAUSS
        const float var = T.gauss_std * T.gauss_std;
        float s = 0.0f;
        for (int i = lane; i < d; i += 32) {
            const float df = xs[i] - T.gauss_mean;
            s += -(logf(6.28318530717958647692f * var) + df * df / var) / 2.0f;
            gs[i] = -df / var;
        }
        return warp_sum(s);
    }
}

// Hessian-vector product (hv != null) and/or Hessian diagonal (hd != null) of the UNTEMPERED
// logprob at xs, for the flow's d/dx[nn_t * grad logprob] term (exe_flow_matching.py:88-90,213,216).
__device__ __forceinline__ void small_target_hess(const mfm_target_t& T, const float* xs, const float* zs,
                                                  float* hv, float* hd, int lane) {
    const int d = T.dim;
    if (T.kind == MFM_TARGET_GMM) {
        float S = 0.0f, g0 = 0.0f, g1 = 0.0f, h00 = 0.0f, h01 = 0.0f, h11 = 0.0f;
        const float x0 = xs[0], x1 = xs[1];
        for (int k = lane; k < T.n_modes; k += 32) {
            const float s0 = T.stds[2 * k], s1 = T.stds[2 * k + 1];
            const float v0 = s0 * s0, v1 = s1 * s1;
            const float d0 = x0 - T.modes[2 * k], d1 = x1 - T.modes[2 * k + 1];
            const float lp0 = -(logf(6.28318530717958647692f * v0) + d0 * d0 / v0) / 2.0f;
            const float lp1 = -(logf(6.28318530717958647692f * v1) + d1 * d1 / v1) / 2.0f;
            const float pk = T.weights[k] * (expf(lp0) * expf(lp1));
            const float a0 = -d0 / v0, a1 = -d1 / v1;
            S += pk; g0 += pk * a0; g1 += pk * a1;
            h00 += pk * (a0 * a0 - 1.0f / v0); h01 += pk * a0 * a1; h11 += pk * (a1 * a1 - 1.0f / v1);
        }
        S = warp_sum(S); g0 = warp_sum(g0) / S; g1 = warp_sum(g1) / S;
        h00 = warp_sum(h00) / S - g0 * g0; h01 = warp_sum(h01) / S - g0 * g1; h11 = warp_sum(h11) / S - g1 * g1;
        if (lane == 0) {
            if (hv) { hv[0] = h00 * zs[0] + h01 * zs[1]; hv[1] = h01 * zs[0] + h11 * zs[1]; }
            if (hd) { hd[0] = h00; hd[1] = h11; }
        }
    } else if (T.kind == MFM_TARGET_PHI4) {
        const float coef = T.phi_a * (float)d;
        for (int i = lane; i < d; i += 32) {
            const float xi = xs[i];
            const float diag = -T.phi_beta * (2.0f * coef - (1.0f - 3.0f * xi * xi) / coef);
            if (hd) hd[i] = diag;
            if (hv) {
                const float zl = i > 0 ? zs[i - 1] : 0.0f, zr = i + 1 < d ? zs[i + 1] : 0.0f;
                hv[i] = diag * zs[i] + T.phi_beta * coef * (zl + zr);
            }
        }
    } else {
        const float var = T.gauss_std * T.gauss_std;
        for (int i = lane; i < d; i += 32) {
            if (hd) hd[i] = -1.0f / var;
            if (hv) hv[i] = -zs[i] / var;
        }
    }
}

}  // namespace mfm
