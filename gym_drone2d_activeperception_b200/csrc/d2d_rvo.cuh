// RVO motion profile (params.motion_profile == "RVO"): RVO.RVO_update / intersect / in_between, utils.py:299-460, called
// from Drone2DEnv2.step (drone_v2.py:169-175) before Agent.step.  One block per env, one warp per agent at a time:
// the lanes build the agent's velocity-obstacle cones in shared memory, then evaluate the ~32 x 5 (+1) candidate
// velocities in parallel; the arg-min keeps Python's first-minimum rule.  All agents read the positions / velocities from
// BEFORE the update (the reference takes its lists once and rebinds agent.velocity), so the new velocities go to
// `avel_next`; the step kernels consume them in Agent.step and publish them as the current velocity.
// Trigonometry: cos / sin of the 32 constant candidate angles come from a host (glibc) table; atan2 / asin / sin / cos of
// run-time values are CUDA's double-precision functions (<= 2 ulp, not glibc's bits): continuous agent state is held
// to 1e-9, not bit-exact, under this profile.  The N * 160 * (N - 1) cone tests of a step compare angles through cross
// products (d2d_rvo_inside_fast) and evaluate atan2 only where that is not certain.
#pragma once
#include "d2d_state.cuh"
#include "d2d_math.cuh"
#include "d2d_rvo_math.cuh"

#define D2D_RVO_MAX_WARPS 8
// warps per block (= agents in flight per env): the count in [4, 8] that leaves the fewest idle warp slots in the last round
// (10 agents: 5 warps x 2 rounds instead of 4 x 3)
__host__ __device__ inline int d2d_rvo_warps(int N) {
    int best = 4, waste = (N + 3) / 4 * 4 - N;
    for (int w = 5; w <= D2D_RVO_MAX_WARPS; w++) {
        const int x = (N + w - 1) / w * w - N;
        if (x < waste) { best = w; waste = x; }
    }
    return best;
}
#define D2D_RVO_MAX_CAND (D2D_RVO_THETAS * 8 + 1)

__host__ __device__ inline size_t d2d_rvo_smem_bytes(int N) {
    const size_t cones = (size_t)(N + D2D_RVO_MAX_OBS) * sizeof(RvoCone);
    const size_t cand = (size_t)D2D_RVO_MAX_CAND * 17;
    return (size_t)N * 32 + d2d_rvo_warps(N) * ((cones + cand + 15) / 16 * 16) + 64;
}

__global__ void __launch_bounds__(D2D_RVO_MAX_WARPS * 32) d2d_rvo_kernel(const DevP P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = P.N, NP = P.NP;
    if (e >= P.B || N <= 0) return;                 // drone_v2.py:170: only with at least one agent
    double2 *spos = (double2 *)smem, *svel = spos + N;
    const size_t per_warp = ((size_t)(N + D2D_RVO_MAX_OBS) * sizeof(RvoCone) + (size_t)D2D_RVO_MAX_CAND * 17 + 15) / 16 * 16;
    unsigned char *wbase = smem + (size_t)N * 32 + (size_t)wid * per_warp;
    RvoCone *cones = (RvoCone *)wbase;
    double *candx = (double *)(cones + N + D2D_RVO_MAX_OBS), *candy = candx + D2D_RVO_MAX_CAND;
    unsigned char *bad = (unsigned char *)(candy + D2D_RVO_MAX_CAND);
    // an env that is about to be re-initialised by this step (lazy reset) plans from its snapshot
    const bool rs = P.rec[e].pending_reset != 0 || (P.auto_reset && P.done[e] != 0);
    const double2 *gpos = (rs ? P.apos0 : P.apos) + (size_t)e * NP, *gvel = (rs ? P.avel0 : P.avel) + (size_t)e * NP;
    const double2 *gpref = (rs ? P.apref0 : P.apref) + (size_t)e * NP;
    for (int k = tid; k < N; k += blockDim.x) { spos[k] = gpos[k]; svel[k] = gvel[k]; }
    __syncthreads();
    const double ROB_RAD = P.arad[(size_t)e * NP] + 0.01;       // agents[0].radius + 0.01, utils.py:309
    const int nobs = P.rvo_nobs[e];
    const double *obs = P.rvo_obs + (size_t)e * D2D_RVO_MAX_OBS * 3;
    const int n_warps = blockDim.x >> 5;
    for (int i = wid; i < N; i += n_warps) {
        const double pAx = spos[i].x, pAy = spos[i].y, vAx = svel[i].x, vAy = svel[i].y;
        const int nc = N - 1 + nobs;
        // ---- cones (utils.py:314-352); their order does not matter: every use is an `any` or a `min`
        for (int j = lane; j < nc; j += 32) {
            if (j < N - 1) {
                const int b = j < i ? j : j + 1;
                d2d_rvo_make_cone(cones[j], pAx, pAy, pAx + 0.5 * (svel[b].x + vAx), pAy + 0.5 * (svel[b].y + vAy), spos[b].x,
                                  spos[b].y, 2 * ROB_RAD);
            } else {
                const double *hh = obs + 3 * (j - (N - 1));
                d2d_rvo_make_cone(cones[j], pAx, pAy, pAx + 0, pAy + 0, hh[0], hh[1], hh[2] * 1.5 + ROB_RAD);
            }
        }
        __syncwarp();
        // ---- candidates (utils.py:362-389): np.arange(0, 2*3.14, 0.2) x np.arange(0.02, norm_v + 0.02, norm_v / 5), then vA
        const double2 w = gpref[i];
        const double norm_v = d2d_norm2(w.x, w.y);
        const double r_start = 0.02, r_step = norm_v / 5.0;
        int n_rad = (int)ceil(((norm_v + 0.02) - r_start) / r_step);      // arange length
        n_rad = n_rad < 0 ? 0 : (n_rad > 8 ? 8 : n_rad);
        const double r_next = r_start + r_step, r_delta = r_next - r_start;   // NumPy fill: start + i * (a[1] - a[0])
        const int ncand = D2D_RVO_THETAS * n_rad + 1;
        bool suit_any = false;
        for (int q = lane; q < ncand - 1; q += 32) {
            const int it = q / n_rad, ir = q - it * n_rad;
            const double rad = ir == 0 ? r_start : (ir == 1 ? r_next : r_start + (double)ir * r_delta);
            const double cx = rad * P.tab->rvo_cos[it], cy = rad * P.tab->rvo_sin[it];
            bool b = false;
            for (int k = 0; k < nc && !b; k++) {
                const double dy = cy + pAy - cones[k].ty, dx = cx + pAx - cones[k].tx;
                const int f = d2d_rvo_inside_fast(cones[k], dx, dy);
                if (f >= 0) b = f != 0;
                else b = d2d_rvo_in_between(cones[k].th_right, atan2(dy, dx), cones[k].th_left);
            }
            candx[q] = cx; candy[q] = cy; bad[q] = b ? 1 : 0;
            suit_any |= !b;
        }
        {   // the last candidate is the preferred velocity itself (utils.py:377-389): its cones go one per lane instead of
            // costing a sixth round of the loop above for a single lane (160 = 5 x 32 candidates come before it)
            bool b = false;
            for (int k = lane; k < nc; k += 32) {
                const double dy = w.y + pAy - cones[k].ty, dx = w.x + pAx - cones[k].tx;
                const int f = d2d_rvo_inside_fast(cones[k], dx, dy);
                if (f >= 0) b |= f != 0;
                else b |= d2d_rvo_in_between(cones[k].th_right, atan2(dy, dx), cones[k].th_left);
            }
            b = __any_sync(0xffffffffu, b);
            if (lane == 0) { candx[ncand - 1] = w.x; candy[ncand - 1] = w.y; bad[ncand - 1] = b ? 1 : 0; }
            suit_any |= !b;
        }
        suit_any = __any_sync(0xffffffffu, suit_any);
        __syncwarp();
        // ---- selection (utils.py:391-431): Python min() keeps the FIRST minimum
        double best_key = 0.0;
        int best = 0x7fffffff;
        for (int q = lane; q < ncand; q += 32) {
            const double ux = candx[q], uy = candy[q];
            double key;
            if (suit_any) {
                if (bad[q]) continue;
                key = d2d_norm2(ux - w.x, uy - w.y);
            } else {
                double tc_min = 0.0;
                bool have = false;
                for (int k = 0; k < nc; k++) {
                    const double dx = ux + pAx - cones[k].tx, dy = uy + pAy - cones[k].ty;
                    const int f = d2d_rvo_inside_fast(cones[k], dx, dy);
                    if (f == 0) continue;                       // certainly outside this cone: no angle needed
                    const double theta_dif = atan2(dy, dx);
                    if (f < 0 && !d2d_rvo_in_between(cones[k].th_right, theta_dif, cones[k].th_left)) continue;
                    const double small_theta = fabs(theta_dif - 0.5 * (cones[k].th_left + cones[k].th_right));
                    double rad = cones[k].rad;
                    const double a = fabs(cones[k].dist * sin(small_theta));
                    if (a >= rad) rad = a;
                    const double big_theta = asin(a / rad);
                    double dist_tg = fabs(cones[k].dist * cos(small_theta)) - fabs(rad * cos(big_theta));
                    if (dist_tg < 0) dist_tg = 0;
                    const double tc_v = dist_tg / d2d_norm2(dx, dy);
                    if (!have || tc_v < tc_min) { tc_min = tc_v; have = true; }
                }
                key = (0.2 / (tc_min + 0.001)) + d2d_norm2(ux - w.x, uy - w.y);
            }
            if (best == 0x7fffffff || key < best_key) { best = q; best_key = key; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double ok = __shfl_down_sync(0xffffffffu, best_key, off);
            const int oq = __shfl_down_sync(0xffffffffu, best, off);
            if (oq != 0x7fffffff && (best == 0x7fffffff || ok < best_key || (ok == best_key && oq < best))) { best = oq; best_key = ok; }
        }
        if (lane == 0) {
            double2 nv;
            nv.x = candx[best]; nv.y = candy[best];
            P.avel_next[(size_t)e * NP + i] = nv;
        }
        __syncwarp();
    }
}
