// ogb_guess.cuh -- initial guesses, multi-start jitter and trajectory unpacking on the device
// (include/ogb200.h: ogb_guess_fill, ogb_jitter, ogb_trajectories).
//
// Reference being replaced (file:line = /root/reference/OpenGoddard/optimize.py): Guess.zeros / constant /
// linear / cubic :883-956, the setters that store a guess :377-440, time_update :518-531,
// states_all_section / controls_all_section :286-331, to_csv :844-863.  All three kernels are plain
// element-wise HBM-bound maps (a few bytes read per 8 bytes written); they exist so that a batch of
// thousands of starts never has to be built on the host and copied over PCIe.
#pragma once
#include "ogb_core.h"

// ------------------------------------------------------------------ Guess.*
// linear: scipy.interpolate.interp1d([t0, tf], [y0, yf]) evaluated with numpy.interp's formula
//         (interp1d._call_linear_np -> compiled_base.c arr_interp): the end points exactly, in between
//         slope * (t - t0) + y0.
// cubic : the reference solves the 4 x 4 Hermite system with numpy.linalg.inv and evaluates
//         C0 + C1 t + C2 t^2 + C3 t^3; here the same polynomial in the Hermite basis on s = (t - t0) / (tf - t0)
//         (better conditioned than the monomial basis; agrees with the reference to ~cond(A) * eps).
OGB_HD double ogb_guess_value(int kind, double t, double t0, double tf, const double* q) {
    switch (kind) {
        case OGB_GUESS_CONSTANT: return 1.0 * q[0];
        case OGB_GUESS_LINEAR: {
            if (t >= tf) return q[1];
            if (t == t0) return q[0];
            const double slope = (q[1] - q[0]) / (tf - t0);
            return slope * (t - t0) + q[0];
        }
        case OGB_GUESS_CUBIC: {
            const double h = tf - t0, s = (t - t0) / h, s2 = s * s, s3 = s2 * s;
            const double h00 = 2.0 * s3 - 3.0 * s2 + 1.0, h10 = s3 - 2.0 * s2 + s;
            const double h01 = -2.0 * s3 + 3.0 * s2, h11 = s3 - s2;
            return h00 * q[0] + h10 * h * q[1] + h01 * q[2] + h11 * h * q[3];
        }
        default: return 0.0;
    }
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256)
ogb_guess_kernel(OgbProb P, const ogb_guess_spec* __restrict__ specs, int nspec, const double* __restrict__ params,
                 const double* __restrict__ time, const double* __restrict__ tfinal, int B, double* __restrict__ out) {
    // one thread per (instance, spec, global node); the final times ride along as "spec nspec"
    const long per = (long)(nspec + 1) * P.gtot;
    const long total = (long)B * per;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / per;
        const int rem = (int)(idx - b * per);
        const int sp = rem / P.gtot, g = rem - sp * P.gtot;
        if (sp == nspec) {
            if (tfinal != nullptr && g < P.nsec)
                out[b * P.n + P.n - P.nsec + g] = tfinal[b * P.nsec + g] / P.unit_time;      // set_time_final
            continue;
        }
        const ogb_guess_spec S = specs[sp];
        const int s = ogb_sec_of_node(P, g);
        if (S.sec >= 0 && S.sec != s) continue;
        const OgbSec& sec = P.sec[s];
        if (S.blk >= sec.nb) continue;
        // the time axis the guess was requested on: one phase, or all phases concatenated
        const int glo = S.sec >= 0 ? sec.g0 : 0, ghi = S.sec >= 0 ? sec.g0 + sec.N : P.gtot;
        const double v = ogb_guess_value(S.kind, time[g], time[glo], time[ghi - 1],
                                         params + ((size_t)b * nspec + sp) * 4);
        double unit = 1.0;
        if (S.blk < sec.ns) unit = P.ustate[sec.us_off + S.blk];
        else {
            int uc = 0;
            for (int q = 0; q < s; ++q) uc += P.sec[q].nc;
            unit = P.ucontrol[uc + S.blk - sec.ns];
        }
        out[b * P.n + sec.off + S.blk * sec.N + (g - sec.g0)] = v / unit;                     // set_states / set_controls
    }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------ multi-start jitter
// Philox4x32-10 (Salmon et al., SC'11): 10 rounds of two 32x32 -> 64 multiplies; counter-based, so every
// (instance, variable) draws its own 128 random bits from (seed, instance, variable) alone.
OGB_HD void ogb_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t* o) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
// two 32-bit words -> a double in [0, 1) with 53 random bits
OGB_HD double ogb_u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256)
ogb_jitter_kernel(OgbProb P, double* __restrict__ X, int B, unsigned long long seed, long long first, double rel_x,
                  double rel_t, const double* __restrict__ lb, const double* __restrict__ ub) {
    const long total = (long)B * P.n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / P.n;
        const int j = (int)(idx - b * P.n);
        const unsigned long long inst = (unsigned long long)(first + b);
        uint32_t r[4];
        ogb_philox4x32((uint32_t)j, 0u, (uint32_t)inst, (uint32_t)(inst >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u1 = ogb_u53(r[0], r[1]), u2 = ogb_u53(r[2], r[3]);
        double x = X[idx];
        if (j < P.n - P.nsec) {
            const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586477 * u2);     // Box-Muller
            x = x * (1.0 + rel_x * z);
        } else {
            x = x * (1.0 + rel_t * (2.0 * u1 - 1.0));
        }
        if (lb != nullptr && x < lb[j]) x = lb[j];
        if (ub != nullptr && x > ub[j]) x = ub[j];
        X[idx] = x;
    }
}

// ------------------------------------------------------------------ trajectories
__global__ void __launch_bounds__(256)
ogb_traj_kernel(OgbProb P, const double* __restrict__ X, int B, double* __restrict__ out) {
    const int ns0 = P.sec[0].ns, nc0 = P.sec[0].nc, W = 1 + ns0 + nc0;
    const long total = (long)B * P.gtot * W;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long bg = idx / W;
        const int col = (int)(idx - bg * W);
        const long b = bg / P.gtot;
        const int g = (int)(bg - b * P.gtot);
        const int s = ogb_sec_of_node(P, g);
        const OgbSec& sec = P.sec[s];
        const double* x = X + b * P.n;
        double v = 0.0;
        if (col == 0) {                      // time_update (:518-531): t = [0] + final times
            const double tb = x[sec.tf_idx] * P.unit_time;
            const double ta = s == 0 ? 0.0 : x[sec.t0_idx] * P.unit_time;
            v = (tb - ta) / 2.0 * P.tau[g] + (tb + ta) / 2.0;
        } else {
            const int blk = col - 1 < ns0 ? col - 1 : sec.ns + (col - 1 - ns0);      // state a / control c of this phase
            const bool is_state = col - 1 < ns0;
            if ((is_state && blk < sec.ns) || (!is_state && blk < sec.nb)) {
                double unit;
                if (is_state) unit = P.ustate[sec.us_off + blk];
                else {
                    int uc = 0;
                    for (int q = 0; q < s; ++q) uc += P.sec[q].nc;
                    unit = P.ucontrol[uc + blk - sec.ns];
                }
                v = x[sec.off + blk * sec.N + (g - sec.g0)] * unit;                   // states() / controls(), :284 / :315
            }
        }
        out[idx] = v;
    }
}
#endif  // __CUDACC__
