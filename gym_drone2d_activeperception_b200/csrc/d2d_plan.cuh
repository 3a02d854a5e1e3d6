// d2d_plan.cuh -- kernels for the Primitive planner path and the Oxford gaze policy (sm_100a).
//
// With planner == Primitive one env step is several launches (the planner verdict sits in the middle of
// Drone2DEnv2.step, drone_v2.py:194-204, and an A* search is ~1000x longer than the rest of the step, so it gets
// its own kernel over a compacted list of the envs that actually need a plan).  Default (one warp per env):
//
//   d2d_step_prim_warp_kernel   agents, rays, trackers + Primitive.replan_check (traj_planner.py:220-233); envs whose
//                               trajectory stays valid -- and envs whose search is decided by its start position alone --
//                               finish their step here; the others are appended to the planning list
//   d2d_plan_small_kernel       Primitive.plan (traj_planner.py:125-218): A* over motion primitives with the reference's
//                               insertion-ordered-dict tie breaking, the whole search in shared memory (4 warps, 5 per SM)
//   d2d_plan_kernel (mode 1)    the searches the small kernel abandoned for lack of node slots (usually none); mode 0: every
//                               search, where the small kernel's packed keys / primitive count do not apply
//   d2d_step_post_list_kernel   brake / step_pos / step_yaw / is_collide / flags / done / observation of the planning envs
//
// Legacy block-per-E-envs path (envs_per_block > 0): d2d_step_pre_kernel -> d2d_plan_kernel -> d2d_step_post_kernel.
//
// d2d_oxford_kernel is Oxford.plan (yaw_planner.py:81-127), one block per env; d2d_oxford_list_kernel the same for the envs
// of the planning list (d2d_step_plan_oxford runs the searches beside the scoring of the other envs).
#pragma once
#include "d2d_state.cuh"
#include "d2d_math.cuh"
#include "d2d_step.cuh"
#include "d2d_plan_math.cuh"

#define D2D_PLAN_SLOTS 296          // concurrent A* workspaces (2 per SM)
#define D2D_PLAN_THREADS 128
#define D2D_HASH_EMPTY 0xFFFFFFFFFFFFFFFFull

// ------------------------------------------------------------------------------------------ shared helpers
// waypoint `a` (absolute index) of the stored trajectory: segment a / n_way, sample time t_way[n_way-1 - a % n_way]
// (the reference builds each segment with t = arange(2, 0, -dt) and reverses the whole list, traj_planner.py:212-217)
// q / d for 0 <= q < 2^15, 1 <= d <= 64 without the integer-division sequence: (q + 0.5) / d is at least 0.5 / 64 away
// from every integer, three orders of magnitude more than the fp32 error of the product at these sizes
__device__ __forceinline__ int d2d_div_small(int q, int d) {
    return __float2int_rz(((float)q + 0.5f) * __frcp_rn((float)d));
}

__device__ __forceinline__ void d2d_waypoint_pos(const DevP &P, int e, int a, double &x, double &y) {
    const int seg = d2d_div_small(a, P.n_way), ws = a - seg * P.n_way, ti = P.n_way - 1 - ws;
    const double *cf = P.traj_coeff + ((size_t)e * D2D_MAX_SEGMENTS + seg) * 6;
    const double t = P.tab->t_way[ti], t2 = P.tab->t_way2[ti];
    x = rint(D2D_FMA(t2, cf[2], cf[0] + t * cf[1]));
    y = rint(D2D_FMA(t2, cf[5], cf[3] + t * cf[4]));
}

// OccupancyGridMap.get_grid on a belief grid in shared memory (utils.py:545-548)
__device__ __forceinline__ int d2d_belief_probe(const DevP &P, const uint8_t *bel, double x, double y) {
    if (x >= P.map_w || x < 0 || y >= P.map_h || y < 0) return 1;
    return bel[d2d_cell(x, P.scale, P.inv_scale) * D2D_GRID + d2d_cell(y, P.scale, P.inv_scale)];
}

// Planner.is_free (traj_planner.py:28-59); trk = active trackers as [mu0, mu1, mu2, mu3, radius]
__device__ __forceinline__ bool d2d_is_free(const DevP &P, const uint8_t *bel, double px, double py, double t,
                                            const double *trk, int nact) {
    if (px != px || py != py) return false;
    const double sd = P.drone_r + 10.0;
    if (d2d_belief_probe(P, bel, px - sd, py) == 1) return false;
    if (d2d_belief_probe(P, bel, px, py) == 1) return false;
    if (d2d_belief_probe(P, bel, px + sd, py) == 1) return false;
    if (d2d_belief_probe(P, bel, px, py - sd) == 1) return false;
    if (d2d_belief_probe(P, bel, px, py + sd) == 1) return false;
    for (int k = 0; k < nact; k++) {
        const double *m = trk + 5 * k;
        const double ex = m[0] + t * m[2], ey = m[1] + t * m[3];   // estimate_pos utils.py:220-223
        if (d2d_norm2(px - ex, py - ey) <= P.drone_r + m[4] + 5.0 + P.var_cam) return false;
    }
    return true;
}

// gathers the env's active trackers into shared memory; returns via *nact (block must sync afterwards)
__device__ __forceinline__ void d2d_gather_trackers(const DevP &P, int e, double *trk, int *nact, int tid, int T) {
#pragma unroll 1
    for (int k = tid; k < P.N; k += T) {
        const size_t g = (size_t)e * P.NP + k;
        if (P.trk_active[g]) {
            const int slot = atomicAdd(nact, 1);
            const double *mu = P.trk_mu + g * 4;
            double *d = trk + 5 * slot;
            d[0] = mu[0]; d[1] = mu[1]; d[2] = mu[2]; d[3] = mu[3]; d[4] = P.trk_radius[g];
        }
    }
}

// ------------------------------------------------------------------------------------------ pre kernel
__host__ __device__ inline size_t d2d_pre_smem_bytes(int E, int NP, int HW) {
    return d2d_step_smem_bytes(E, NP, HW) + (size_t)E * NP * 5 * 8 + (size_t)E * 16;
}

template <int E>
__global__ void __launch_bounds__((E * 50 + 31) / 32 * 32, (E == 4 ? 4 : (E == 8 ? 2 : 1)))
d2d_step_pre_kernel(const DevP P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const BlockCtx c = d2d_carve(smem, E, P.NP, P.HW);
    double *trk = (double *)(smem + d2d_step_smem_bytes(E, P.NP, P.HW));   // [E][NP][5]
    int *nact = (int *)(trk + (size_t)E * P.NP * 5);                        // [E]
    int *rep = nact + E;                                                    // [E] replan flags
    const int tid = threadIdx.x, T = blockDim.x;
    const int env0 = blockIdx.x * E;

    if (tid == 0) { d2d_mbar_init(c.mbar, 1); c.misc[0] = 0; c.misc[1] = 0; }
    if (blockIdx.x == 0 && tid == 0) P.plan_list[P.B + 5] = 0;              // ticket counter of d2d_plan_kernel
    for (int w = tid; w < E * P.HW; w += T) c.hitw[w] = 0u;
    if (tid < E) { nact[tid] = 0; rep[tid] = 0; }
    __syncthreads();
    if (tid < E) {
        d2d_load_env_scalars(P, c.S[tid], env0 + tid);
        if (c.S[tid].reset) c.misc[1] = 1;
    }
    __syncthreads();
    if (tid == 0) d2d_issue_bulk(P, c, env0, E, true);
    d2d_reset_arrays(P, c, env0, E, tid, T);
    d2d_phase_agents<false>(P, c, env0, E, tid, T);
    if (tid < E && c.S[tid].valid) d2d_leader_begin(P, c.S[tid]);
    __syncthreads();
    d2d_mbar_wait(c.mbar, 0);
    d2d_phase_rays(P, c, env0, E, tid, T);
    __syncthreads();
    d2d_phase_trackers(P, c, env0, E, tid, T);
    __syncthreads();
    // ---- Primitive.replan_check (traj_planner.py:220-233)
    for (int i = 0; i < E; i++)
        if (c.S[i].valid) d2d_gather_trackers(P, env0 + i, trk + (size_t)i * P.NP * 5, &nact[i], tid, T);
    __syncthreads();
    for (int i = 0; i < E; i++) {
        const EnvS &s = c.S[i];
        if (!s.valid) continue;
        const int len = s.nseg * P.n_way - s.cursor;
        const uint8_t *bel = c.belief + (size_t)i * D2D_BELIEF_STRIDE;
        const double *tk = trk + (size_t)i * P.NP * 5;
        const int na = nact[i];
        bool hit = false;
        for (int w = tid; w < len && !hit; w += T) {
            double x, y;
            d2d_waypoint_pos(P, env0 + i, s.cursor + w, x, y);
            const double ti = (double)w * P.dt;
            // static overlap: swep_map is uint8 (zeros_like of the belief grid), so i*dt is truncated (:222-224,230)
            const int ci = d2d_cell(x, P.scale, P.inv_scale), cj = d2d_cell(y, P.scale, P.inv_scale);
            if ((unsigned)ci < (unsigned)D2D_GRID && (unsigned)cj < (unsigned)D2D_GRID) {
                if (bel[ci * D2D_GRID + cj] == 1 && (uint8_t)ti > 0) hit = true;
            }
            for (int k = 0; k < na && !hit; k++) {
                const double *m = tk + 5 * k;
                const double ex = m[0] + ti * m[2], ey = m[1] + ti * m[3];
                if (d2d_norm2(ex - x, ey - y) <= P.drone_r + m[4]) hit = true;   // :225-229
            }
        }
        if (hit) rep[i] = 1;
    }
    __syncthreads();
    if (tid < E && c.S[tid].valid) {
        EnvS &s = c.S[tid];
        const int e = env0 + tid;
        if (rep[tid]) { s.nseg = 0; s.cursor = 0; atomicAdd(&P.stats[D2D_STAT_REPLANS], 1ull); }   // trajectory.clear()
        const int need = (s.nseg * P.n_way - s.cursor) == 0;
        P.replan[e] = (uint8_t)rep[tid];
        P.need_plan[e] = (uint8_t)need;
        P.plan_ok[e] = 1;
        if (need) {
            const int slot = atomicAdd(&P.plan_list[P.B], 1);
            P.plan_list[slot] = e;
        }
        // tracker bookkeeping of this step; the still-active totals travel to the post kernel
        s.bufc += s.arch_cnt; s.bufts += s.arch_ts; s.tracked += s.newly;
        P.tmp_act_cnt[e] = s.act_cnt; P.tmp_act_ts[e] = s.act_ts;
        d2d_store_env_scalars(P, s, e);
        P.done[e] = 0;   // consumed by the lazy reset above; the post kernel writes this step's verdict
    }
}

// ------------------------------------------------------------------------------------------ warp-per-env Primitive step
// One warp per env.  Everything up to the planner verdict runs as in the fused kernel; an env whose trajectory is
// still valid (the vast majority) then finishes its step right here -- step_pos / step_yaw / is_collide / flags /
// observation -- with the belief grid still in shared memory.  Only envs that need a plan are deferred: they are
// appended to the compacted list for d2d_plan_kernel and completed by d2d_step_post_list_kernel.
// The cells changed by this step's rays are recorded, so when the drone's cell (the window origin) is unchanged the
// observation tensor is patched instead of rewritten.
// (list of the env's active trackers [NP][5] +) changed-cell list + two counters.  d2d_step_prim_warp_kernel keeps the tracker list
// in the agent arrays sx / sy / sr2 / mx / my (5 contiguous [NP] doubles: exactly its size), which nothing reads once the
// tracker phase is over -- the drone-vs-agent test of the finish reads the new positions back from P.apos instead.  At 142
// agents that takes a warp's slice from 15.5 KB to 9.8 KB: 5 blocks (20 warps) per SM instead of 3 (12).
__host__ __device__ inline size_t d2d_trk_warp_extra(int NP) { return (size_t)NP * 5 * 8 + D2D_CHG_CAP * 4 + 32; }
__host__ __device__ inline size_t d2d_prim_warp_extra(int NP) { (void)NP; return (size_t)D2D_CHG_CAP * 4 + 32; }

__device__ D2D_COLD void d2d_finish_env_warp(const DevP &P, const BlockCtx &c, EnvS &s, int e, int lane,
                                                    double action, bool success, bool agents_in_smem, int nchg,
                                                    const uint32_t *chg, bool allow_patch) {
    if (lane == 0) d2d_leader_finish(P, s, c.gt, e, action, success);
    __syncwarp();
    // Drone2D.is_collide (utils.py:764-778) against the drone's NEW position
    bool hit = false;
#pragma unroll 1
    for (int k = lane; k < P.N; k += 32) {
        const size_t g = (size_t)e * P.NP + k;
        double ax, ay;
        if (agents_in_smem) { ax = c.sx[k]; ay = c.sy[k]; }
        else { const double2 q = P.apos[g]; ax = q.x; ay = q.y; }
        if (d2d_norm2(ax - s.px, ay - s.py) < P.arad[g] + P.drone_r) hit = true;
    }
    const bool any_hit = __any_sync(0xffffffffu, hit);
    const int shit = __any_sync(0xffffffffu, lane < 5 ? d2d_static_probe(P, c.gt, s.px, s.py, lane) : 0);
    if (lane == 0) {
        s.coll_agent = any_hit ? 1 : 0;
        d2d_leader_flags(P, s, c.gt, e, shit);
    }
    __syncwarp();
    // the observation tensor holds the window of cell (obs_ix, obs_iy): patch it if that is still the drone's cell
    const bool do_patch = allow_patch && !s.reset && s.obs_ix == s.ix && s.obs_iy == s.iy && nchg <= D2D_CHG_CAP;
    __syncwarp();
    if (!do_patch && lane == 0) { s.obs_ix = s.ix; s.obs_iy = s.iy; }
    __syncwarp();
    d2d_store_env_warp(P, s, e, lane);
    if (do_patch) {
        uint8_t *out = P.local_map + (size_t)e * D2D_LOCAL_CELLS;
        uint8_t *out_m = P.lm_mirror ? P.lm_mirror + (size_t)e * D2D_LOCAL_CELLS : nullptr;
#pragma unroll 1
        for (int q = lane; q < nchg; q += 32) {
            const int cell = (int)(chg[q] & 0xFFFFu), v = (int)(chg[q] >> 16);
            const int u = cell / D2D_GRID - (s.ix - 16), w = cell % D2D_GRID - (s.iy - 16);
            if ((unsigned)u < (unsigned)D2D_LOCAL && (unsigned)w < (unsigned)D2D_LOCAL) {
                out[u * D2D_LOCAL + w] = (uint8_t)v;
                if (out_m) out_m[u * D2D_LOCAL + w] = (uint8_t)v;
            }
        }
        if (out_m && lane == 0 && nchg) atomicAdd(&P.stats[D2D_STAT_MIRROR_BYTES], (unsigned long long)nchg);
    } else {
        d2d_obs_env_warp(P, c.belief, s.ix, s.iy, e, lane);
    }
    if (s.done_now) d2d_count_explored_warp(P, c.belief, lane);
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 28 / WPB) d2d_step_prim_warp_kernel(const DevP P,
                                                                               const double *__restrict__ actions) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int e = blockIdx.x * WPB + wid;
    if (e >= P.B) return;
    unsigned char *slice = smem + (size_t)wid * d2d_warp_slice_bytes(P.NP, P.HW, d2d_prim_warp_extra(P.NP));
    const BlockCtx c = d2d_carve(slice, 1, P.NP, P.HW);
    double *trk = c.sx;                                                    // [NP][5] over sx / sy / sr2 / mx / my, from the gather on
    uint32_t *chg = (uint32_t *)(slice + d2d_step_smem_bytes(1, P.NP, P.HW));   // [D2D_CHG_CAP]
    int *cnt = (int *)(chg + D2D_CHG_CAP);                                 // [0] nact, [1] nchg
    EnvS &s = c.S[0];
    // plan-list counters are double buffered by a step parity that lives in DEVICE memory (plan_list[B+3], advanced by
    // the post kernel), so a captured CUDA graph replays correctly; [B+4] publishes this step's parity to the later kernels
    const int par = P.plan_list[P.B + 3] & 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.plan_list[P.B + 4] = par; P.plan_list[P.B + 1 + (par ^ 1)] = 0;
        P.plan_list[P.B + 5] = 0;                                            // ticket counter of the A* kernels
        P.plan_over[P.B] = 0; P.plan_over[P.B + 1] = 0;                      // overflow list of d2d_plan_small_kernel: count, ticket
        atomicAdd(&P.stats[D2D_STAT_ENV_STEPS], (unsigned long long)P.B);    // every env steps once per d2d_step
    }

    // all HBM requests up front (see d2d_step_fused_warp_kernel): agents speculatively from the live arrays, bulk copies,
    // then the scalars
    double2 pf_pos = double2{0.0, 0.0}, pf_pref = double2{0.0, 0.0};
    double pf_r = 0.0;
    uint8_t pf_act = 0;
    if (lane < P.N) {
        const size_t g = (size_t)e * P.NP + lane;
        pf_pos = P.apos[g]; pf_pref = P.apref[g]; pf_r = P.arad[g];
        if (P.trackers) pf_act = P.trk_active[g];
    }
    const double action = actions[e];
    if (lane == 0) {
        d2d_mbar_init(c.mbar, 1);
        d2d_mbar_expect_tx(c.mbar, D2D_GT_ROW_BYTES + D2D_BELIEF_STRIDE);
        d2d_bulk_g2s(c.gt, P.gt_rows + (size_t)e * D2D_GRID, D2D_GT_ROW_BYTES, c.mbar);
        d2d_bulk_g2s(c.belief, P.belief + (size_t)e * D2D_BELIEF_STRIDE, D2D_BELIEF_STRIDE, c.mbar);
        c.misc[0] = 0;
        cnt[0] = 0; cnt[1] = 0;
    }
    d2d_load_env_warp(P, s, e, lane);
    if (lane == 0) c.misc[1] = s.reset;
#pragma unroll 1
    for (int w = lane; w < P.HW; w += 32) c.hitw[w] = 0u;
    __syncwarp();
    d2d_reset_prefetch(P, s, e, lane, pf_pos, pf_pref);
    d2d_reset_arrays(P, c, e, 1, lane, 32, c.mbar);
    d2d_phase_agents<false, true>(P, c, e, 1, lane, 32, pf_pos, pf_pref, pf_r);
    if (pf_act && !s.reset) d2d_prefetch_tracker(P, (size_t)e * P.NP + lane);
    if (lane == 0) d2d_leader_begin(P, s);
    __syncwarp();
    RayOut ro;
    ro.bel_s = c.belief; ro.e = e; ro.patch = 0; ro.wi = 0; ro.wj = 0; ro.chg = chg; ro.nchg = &cnt[1]; ro.defer_mirror = 0;
    d2d_mbar_wait(c.mbar, 0);
    ro.border_ok = d2d_border_intact(c.gt, lane);
    d2d_phase_rays_warp<false>(P, c, ro, lane);
    __syncwarp();
    if (P.var_cam != 0.0) {
        if (lane == 0) d2d_measure_env(P, c, e);
        __syncwarp();
    }
    d2d_phase_trackers<true>(P, c, e, 1, lane, 32, pf_act);
    __syncwarp();
    // ---- Primitive.replan_check (traj_planner.py:220-233)
    d2d_gather_trackers(P, e, trk, &cnt[0], lane, 32);
    __syncwarp();
    const int na = cnt[0];
    const int len = s.nseg * P.n_way - s.cursor;
    bool hit = false;
#pragma unroll 1
    for (int w = lane; w < len && !hit; w += 32) {
        double x, y;
        d2d_waypoint_pos(P, e, s.cursor + w, x, y);
        const double ti = (double)w * P.dt;
        const int ci = d2d_cell(x, P.scale, P.inv_scale), cj = d2d_cell(y, P.scale, P.inv_scale);
        if ((unsigned)ci < (unsigned)D2D_GRID && (unsigned)cj < (unsigned)D2D_GRID)
            if (c.belief[ci * D2D_GRID + cj] == 1 && (uint8_t)ti > 0) hit = true;        // uint8 swep_map (:222-224,230)
#pragma unroll 1
        for (int k = 0; k < na && !hit; k++) {
            const double *m = trk + 5 * k;
            const double ex = m[0] + ti * m[2], ey = m[1] + ti * m[3];
            if (d2d_norm2_le(ex - x, ey - y, P.drone_r + m[4])) hit = true;              // :225-229
        }
    }
    const bool rep = __any_sync(0xffffffffu, hit);
    __syncwarp();
    if (lane == 0) {
        if (rep) { s.nseg = 0; s.cursor = 0; atomicAdd(&P.stats[D2D_STAT_REPLANS], 1ull); }   // trajectory.clear()
        s.bufc += s.arch_cnt; s.bufts += s.arch_ts; s.tracked += s.newly;
        s.arch_cnt = 0; s.arch_ts = 0; s.newly = 0;
        P.replan[e] = rep ? 1 : 0;
        P.plan_ok[e] = 1;
    }
    __syncwarp();
    const bool need = (s.nseg * P.n_way - s.cursor) == 0;
    // A search whose start position is itself not free ends after one expansion: sample 0 of EVERY primitive is the start
    // position at global time 0 (t = arange(0, 2, ...)[0], traj_planner.py:179-183), so no successor is added, the open set
    // is empty and plan() returns False (:149-153) -- unless the start already lies within the goal threshold (:160).  74 %
    // of the searches of BASELINE config 3 and 44 % of config 4 are of this kind; they get their verdict right here, with
    // the belief grid and the trackers in shared memory, instead of a slot of d2d_plan_kernel + d2d_step_post_list_kernel.
    bool blocked = false;
    if (need && !d2d_norm2_le(s.px - s.tgx, s.py - s.tgy, 10.0)) {
        const double qx = rint(s.px), qy = rint(s.py);
        bool occ = false;
        if (qx != qx || qy != qy || !(fabs(qx) < 1e6 && fabs(qy) < 1e6)) occ = true;
        else {
            if (lane < 5) {       // Planner.is_free's five probes (traj_planner.py:35-47), one per lane
                const int sd = (int)(P.drone_r + 10.0), x = (int)qx + (lane == 0 ? -sd : (lane == 2 ? sd : 0)),
                          y = (int)qy + (lane == 3 ? -sd : (lane == 4 ? sd : 0));
                occ = d2d_belief_probe_int(c.belief, x, y, (int)P.map_w, (int)P.map_h) == 1;
            }
#pragma unroll 1
            for (int k = lane; k < na; k += 32) {
                const double *m = trk + 5 * k;
                const double ex = m[0] + 0.0 * m[2], ey = m[1] + 0.0 * m[3];            // estimate_pos(0) utils.py:220-223
                if (d2d_norm2_le(qx - ex, qy - ey, P.drone_r + m[4] + 5.0 + P.var_cam)) occ = true;
            }
        }
        blocked = __any_sync(0xffffffffu, occ);
    }
    __syncwarp();                 // every lane has read the trajectory length before lane 0 pops a waypoint (step_pos)
    if (!need || blocked) {
        if (lane == 0) {
            P.need_plan[e] = blocked ? 2 : 0;     // 0: trajectory still valid, 1: waits for an A* search, 2: search decided here
            if (blocked) {
                s.nseg = 0; s.cursor = 0; P.plan_ok[e] = 0;
                atomicAdd(&P.stats[D2D_STAT_PLANS], 1ull);
                atomicAdd(&P.stats[D2D_STAT_PLAN_FAILURES], 1ull);
            }
        }
        __syncwarp();
        d2d_finish_env_warp(P, c, s, e, lane, action, !blocked, false, cnt[1], chg, true);    // agents from P.apos: sx / sy hold the tracker list
    } else {
        if (lane == 0) {
            P.need_plan[e] = 1;
            const int slot = atomicAdd(&P.plan_list[P.B + 1 + par], 1);
            P.plan_list[slot] = e;
            P.tmp_act_cnt[e] = s.act_cnt; P.tmp_act_ts[e] = s.act_ts;
            P.done[e] = 0;
        }
        d2d_store_env_warp(P, s, e, lane);
    }
}


// ------------------------------------------------------------------------------------------ Jerk_Primitive  traj_planner.py:403-516
// Every step: rank the 72 headings by squared angular distance to the goal bearing (:471-478), take the first heading whose
// minimum-jerk primitive (Mueller's closed form, :413-460) is collision-free at every sample (:482-491) and append its FIRST
// sample as the waypoint step_pos consumes in the same step (:496-499).  One warp per env, whole step in one launch.
//   * heading order: unique (ascending cost) unless the bearing is an exact multiple of 2.5 degrees, where pairs tie and the
//     order is the host numpy's recorded argsort (d2d_jerk_tables.tie_order);
//   * per-heading constants (end-point offset, T and its powers, sample times and their powers) come from the host tables;
//   * evaluation: lanes = (heading in rank order, sample); PACK headings are tested per pass, the lowest-ranked free one wins.
// replan_check (:503-516) always sees an empty trajectory (plan appends one waypoint, step_pos pops it) and returns False.
__device__ __forceinline__ bool d2d_jerk_plan_warp(const DevP &P, EnvS &s, const uint8_t *bel, const double *trk, int nact,
                                                   int e, int lane, double *cost_s, uint8_t *order_s) {
    const d2d_jerk_tables &J = *P.jerk;
    const double RAD2DEG = 180.0 / D2D_PI;
    const double p0x = s.px, p0y = s.py, v0x = s.vx, v0y = s.vy;
    const double2 a0 = s.reset ? double2{0.0, 0.0} : P.drone_acc[e];
    const double phi_h = d2d_atan2_cr(s.tgy - p0y, s.tgx - p0x) * RAD2DEG;
    const double pm = d2d_pymod(phi_h, 360.0);
    const double q = pm / 2.5;
    if (q == floor(q) && q >= 0.0 && q < 144.0) {
        const uint8_t *row = J.tie_order[(int)q];
        for (int i = lane; i < D2D_JERK_H; i += 32) order_s[i] = row[i];
    } else {
        for (int i = lane; i < D2D_JERK_H; i += 32) {
            const double d = fabs(5.0 * (double)i - pm);
            const double c = d <= 180.0 ? d : 360.0 - d;
            cost_s[i] = c * c;
        }
        __syncwarp();
        for (int i = lane; i < D2D_JERK_H; i += 32) {          // rank by counting (stable; the costs are distinct here)
            const double ci = cost_s[i];
            int r = 0;
#pragma unroll 1
            for (int j = 0; j < D2D_JERK_H; j++) { const double cj = cost_s[j]; r += (cj < ci || (cj == ci && j < i)) ? 1 : 0; }
            order_s[r] = (uint8_t)i;
        }
    }
    __syncwarp();
    int tmax = 1;
    for (int i = lane; i < D2D_JERK_H; i += 32) tmax = max(tmax, J.times[i]);
    tmax = __reduce_max_sync(0xffffffffu, tmax);
    const int pack = tmax <= 32 ? 32 / tmax : 1;               // headings tested per pass
    const double v_max = P.max_speed;
    bool found = false;
#pragma unroll 1
    for (int r0 = 0; r0 < D2D_JERK_H && !found; r0 += pack) {
        // samples beyond 32 (drone_max_speed < ~12): the lane walks them with stride 32 (pack == 1 then)
        const int sub = pack > 1 ? lane / tmax : 0, jj0 = pack > 1 ? lane - sub * tmax : lane;
        const int r = r0 + sub;
        const bool live = sub < pack && r < D2D_JERK_H;
        const int h = live ? order_s[r] : 0;
        bool coll = false;
        double fpx = 0, fpy = 0, fvx = 0, fvy = 0, fax = 0, fay = 0;
        if (live) {
            const double T = J.T[h];
            const double *Tp = J.Tp[h];
            const double pfx = p0x + J.dx[h], pfy = p0y + J.dy[h];
            const double lx = s.tgx - pfx, ly = s.tgy - pfy;
            const double sc = 0.5 * v_max / d2d_norm2(lx, ly);
            const double vfx = sc * lx, vfy = sc * ly;
            double al[2], be[2], ga[2];
            const double a0v[2] = {a0.x, a0.y}, v0v[2] = {v0x, v0y}, p0v[2] = {p0x, p0y}, pfv[2] = {pfx, pfy}, vfv[2] = {vfx, vfy};
#pragma unroll
            for (int ii = 0; ii < 2; ii++) {
                const double delt_a = 0.0 - a0v[ii];
                const double delt_v = vfv[ii] - v0v[ii] - a0v[ii] * T;
                const double delt_p = pfv[ii] - p0v[ii] - v0v[ii] * T - 0.5 * a0v[ii] * Tp[0];
                al[ii] = delt_a * 60.0 / Tp[1] - delt_v * 360.0 / Tp[2] + delt_p * 720.0 / Tp[3];
                be[ii] = -delt_a * 24.0 / Tp[0] + delt_v * 168.0 / Tp[1] - delt_p * 360.0 / Tp[2];
                ga[ii] = delt_a * 3.0 / T - delt_v * 24.0 / Tp[0] + delt_p * 60.0 / Tp[1];
            }
            const int times = J.times[h];
#pragma unroll 1
            for (int jj = jj0; jj < times; jj += (pack > 1 ? D2D_JERK_MAXT : 32)) {
                const double tt = J.tt[h][jj];
                const double *tp = J.ttp[h][jj];
                double pos[2], vel[2], acc[2];
#pragma unroll
                for (int ii = 0; ii < 2; ii++) {
                    pos[ii] = al[ii] / 120.0 * tp[3] + be[ii] / 24.0 * tp[2] + ga[ii] / 6.0 * tp[1] + a0v[ii] / 2.0 * tp[0] + v0v[ii] * tt + p0v[ii];
                    vel[ii] = al[ii] / 24.0 * tp[2] + be[ii] / 6.0 * tp[1] + ga[ii] / 2.0 * tp[0] + a0v[ii] * tt + v0v[ii];
                    acc[ii] = al[ii] / 6.0 * tp[1] + be[ii] / 2.0 * tp[0] + ga[ii] * tt + a0v[ii];
                }
                if (jj == 0) { fpx = pos[0]; fpy = pos[1]; fvx = vel[0]; fvy = vel[1]; fax = acc[0]; fay = acc[1]; }
                if (!d2d_is_free(P, bel, pos[0], pos[1], tt, trk, nact)) coll = true;
            }
        }
        const unsigned cm = __ballot_sync(0xffffffffu, coll);
        int win = -1;
        for (int k = 0; k < pack && win < 0; k++) {
            if (r0 + k >= D2D_JERK_H) break;
            const unsigned m = pack > 1 ? (((tmax >= 32 ? 0xffffffffu : ((1u << tmax) - 1u))) << (k * tmax)) : 0xffffffffu;
            if ((cm & m) == 0u) win = k;
        }
        if (win >= 0) {
            found = true;
            const int src = pack > 1 ? win * tmax : 0;             // the lane that evaluated sample 0 of the winning heading
            fpx = __shfl_sync(0xffffffffu, fpx, src); fpy = __shfl_sync(0xffffffffu, fpy, src);
            fvx = __shfl_sync(0xffffffffu, fvx, src); fvy = __shfl_sync(0xffffffffu, fvy, src);
            fax = __shfl_sync(0xffffffffu, fax, src); fay = __shfl_sync(0xffffffffu, fay, src);
            if (lane == 0) {
                s.jerk_has = 1; s.jerk_px = fpx; s.jerk_py = fpy; s.jerk_vx = fvx; s.jerk_vy = fvy; s.jerk_ax = fax; s.jerk_ay = fay;
            }
        }
    }
    __syncwarp();
    return found;
}

__host__ __device__ inline size_t d2d_jerk_warp_extra(int NP) { return d2d_trk_warp_extra(NP) + D2D_JERK_H * 8 + 96; }

template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 28 / WPB) d2d_step_jerk_warp_kernel(const DevP P,
                                                                               const double *__restrict__ actions) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int e = blockIdx.x * WPB + wid;
    if (e >= P.B) return;
    unsigned char *slice = smem + (size_t)wid * d2d_warp_slice_bytes(P.NP, P.HW, d2d_jerk_warp_extra(P.NP));
    const BlockCtx c = d2d_carve(slice, 1, P.NP, P.HW);
    double *trk = (double *)(slice + d2d_step_smem_bytes(1, P.NP, P.HW));   // [NP][5]
    uint32_t *chg = (uint32_t *)(trk + (size_t)P.NP * 5);                  // [D2D_CHG_CAP]
    int *cnt = (int *)(chg + D2D_CHG_CAP);                                 // [0] nact, [1] nchg
    double *cost_s = (double *)(cnt + 8);                                  // [72]
    uint8_t *order_s = (uint8_t *)(cost_s + D2D_JERK_H);                   // [72]
    EnvS &s = c.S[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[D2D_STAT_ENV_STEPS], (unsigned long long)P.B);
    double2 pf_pos = double2{0.0, 0.0}, pf_pref = double2{0.0, 0.0};
    double pf_r = 0.0;
    uint8_t pf_act = 0;
    if (lane < P.N) {
        const size_t g = (size_t)e * P.NP + lane;
        pf_pos = P.apos[g]; pf_pref = P.apref[g]; pf_r = P.arad[g];
        if (P.trackers) pf_act = P.trk_active[g];
    }
    const double action = actions[e];
    if (lane == 0) {
        d2d_mbar_init(c.mbar, 1);
        d2d_mbar_expect_tx(c.mbar, D2D_GT_ROW_BYTES + D2D_BELIEF_STRIDE);
        d2d_bulk_g2s(c.gt, P.gt_rows + (size_t)e * D2D_GRID, D2D_GT_ROW_BYTES, c.mbar);
        d2d_bulk_g2s(c.belief, P.belief + (size_t)e * D2D_BELIEF_STRIDE, D2D_BELIEF_STRIDE, c.mbar);
        c.misc[0] = 0;
        cnt[0] = 0; cnt[1] = 0;
    }
    d2d_load_env_warp(P, s, e, lane);
    if (lane == 0) c.misc[1] = s.reset;
#pragma unroll 1
    for (int w = lane; w < P.HW; w += 32) c.hitw[w] = 0u;
    __syncwarp();
    d2d_reset_prefetch(P, s, e, lane, pf_pos, pf_pref);
    d2d_reset_arrays(P, c, e, 1, lane, 32, c.mbar);
    d2d_phase_agents<false, true>(P, c, e, 1, lane, 32, pf_pos, pf_pref, pf_r);
    if (pf_act && !s.reset) d2d_prefetch_tracker(P, (size_t)e * P.NP + lane);
    if (lane == 0) d2d_leader_begin(P, s);
    __syncwarp();
    RayOut ro;
    ro.bel_s = c.belief; ro.e = e; ro.patch = 0; ro.wi = 0; ro.wj = 0; ro.chg = chg; ro.nchg = &cnt[1]; ro.defer_mirror = 0;
    d2d_mbar_wait(c.mbar, 0);
    ro.border_ok = d2d_border_intact(c.gt, lane);
    d2d_phase_rays_warp<false>(P, c, ro, lane);
    __syncwarp();
    if (P.var_cam != 0.0) {
        if (lane == 0) d2d_measure_env(P, c, e);
        __syncwarp();
    }
    d2d_phase_trackers<true>(P, c, e, 1, lane, 32, pf_act);
    __syncwarp();
    d2d_gather_trackers(P, e, trk, &cnt[0], lane, 32);
    __syncwarp();
    if (lane == 0) {
        s.bufc += s.arch_cnt; s.bufts += s.arch_ts; s.tracked += s.newly;
        s.arch_cnt = 0; s.arch_ts = 0; s.newly = 0;
    }
    __syncwarp();
    const bool ok = d2d_jerk_plan_warp(P, s, c.belief, trk, cnt[0], e, lane, cost_s, order_s);
    if (lane == 0) {
        P.replan[e] = 0; P.plan_ok[e] = ok ? 1 : 0; P.need_plan[e] = 0;
        if (s.reset) P.drone_acc[e] = double2{0.0, 0.0};
        atomicAdd(&P.stats[D2D_STAT_PLANS], 1ull);
        if (!ok) atomicAdd(&P.stats[D2D_STAT_PLAN_FAILURES], 1ull);
    }
    __syncwarp();
    d2d_finish_env_warp(P, c, s, e, lane, action, ok, true, cnt[1], chg, true);
}

// completes the step of the envs that went through d2d_plan_kernel (one warp per list entry, grid-stride)
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) d2d_step_post_list_kernel(const DevP P, const double *__restrict__ actions) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int par = P.plan_list[P.B + 4];
    const int count = min(P.plan_list[P.B + 1 + par], P.B);
    if (blockIdx.x == 0 && threadIdx.x == 0) P.plan_list[P.B + 3] += 1;     // next step uses the other counter
    const BlockCtx c = d2d_carve(smem + (size_t)wid * d2d_warp_slice_bytes(1, 1, 0), 1, 1, 1);
    EnvS &s = c.S[0];
    uint32_t phase = 0;
    if (lane == 0) d2d_mbar_init(c.mbar, 1);
    __syncwarp();
    for (int li = blockIdx.x * WPB + wid; li < count; li += gridDim.x * WPB) {
        const int e = P.plan_list[li];
        if (lane == 0) {
            d2d_load_env_scalars(P, s, e);       // done == 0 and pending_reset == 0 here: plain reload
            s.act_cnt = P.tmp_act_cnt[e]; s.act_ts = P.tmp_act_ts[e];
        }
        __syncwarp();
        if (lane == 0) d2d_issue_bulk(P, c, e, 1, true);
        d2d_mbar_wait(c.mbar, phase);
        phase ^= 1u;
        d2d_finish_env_warp(P, c, s, e, lane, actions[e], P.plan_ok[e] != 0, false, 0, nullptr, false);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ A* workspace
struct PlanWs {
    double *px, *py, *vx, *vy, *cost, *total, *open_total;
    int *parent, *itr, *act;
    unsigned long long *hkeys;
    int *hvals;
    int cap, hcap;
};

__host__ __device__ inline int d2d_plan_cap(int n_u) { return 99 * n_u * n_u + 8; }
__host__ __device__ inline int d2d_plan_hcap(int n_u) {
    int need = 2 * d2d_plan_cap(n_u), h = 1024;
    while (h < need) h <<= 1;
    return h;
}
__host__ __device__ inline size_t d2d_plan_workspace_bytes(int n_u) {
    const size_t cap = (size_t)d2d_plan_cap(n_u), hcap = (size_t)d2d_plan_hcap(n_u);
    size_t b = cap * 8 * 7 + cap * 4 * 3 + hcap * 8 + hcap * 4;
    return (b + 255) / 256 * 256;
}
__device__ __forceinline__ PlanWs d2d_plan_carve(unsigned char *base, int n_u) {
    PlanWs w;
    w.cap = d2d_plan_cap(n_u); w.hcap = d2d_plan_hcap(n_u);
    double *d = (double *)base;
    w.px = d; w.py = d + w.cap; w.vx = d + 2 * (size_t)w.cap; w.vy = d + 3 * (size_t)w.cap;
    w.cost = d + 4 * (size_t)w.cap; w.total = d + 5 * (size_t)w.cap; w.open_total = d + 6 * (size_t)w.cap;
    w.hkeys = (unsigned long long *)(d + 7 * (size_t)w.cap);
    int *ip = (int *)(w.hkeys + w.hcap);
    w.hvals = ip; ip += w.hcap;
    w.parent = ip; w.itr = ip + w.cap; w.act = ip + 2 * (size_t)w.cap;
    return w;
}

__device__ __forceinline__ unsigned long long d2d_node_key(double px, double py, double vx, double vy) {
    const long long a = d2d_floordiv10((long long)rint(px)), b = d2d_floordiv10((long long)rint(py));
    const long long c = (long long)rint(vx), d = (long long)rint(vy);
    return ((unsigned long long)(unsigned short)(short)a << 48) | ((unsigned long long)(unsigned short)(short)b << 32) |
           ((unsigned long long)(unsigned short)(short)c << 16) | (unsigned long long)(unsigned short)(short)d;
}
__device__ __forceinline__ int d2d_hash_slot0(unsigned long long k, int hcap) {
    return (int)(((k * 0x9E3779B97F4A7C15ull) >> 40) & (unsigned long long)(hcap - 1));
}
// returns slot; *existed tells whether the key was already present
__device__ __forceinline__ int d2d_hash_find_or_insert(const PlanWs &w, unsigned long long key, bool *existed) {
    int s = d2d_hash_slot0(key, w.hcap);
    for (;;) {
        const unsigned long long cur = w.hkeys[s];
        if (cur == key) { *existed = true; return s; }
        if (cur == D2D_HASH_EMPTY) {
            const unsigned long long old = atomicCAS(&w.hkeys[s], D2D_HASH_EMPTY, key);
            if (old == D2D_HASH_EMPTY) { *existed = false; return s; }
            if (old == key) { *existed = true; return s; }
        }
        s = (s + 1) & (w.hcap - 1);
    }
}

// total_cost of a node (traj_planner.py:88)
__device__ __forceinline__ double d2d_node_total(double cost, double px, double py, double vx, double vy, double tx,
                                                 double ty) {
    return cost + 0.5 * d2d_norm2(px - tx, py - ty) + 0.1 * d2d_norm2(vx, vy);
}

// ------------------------------------------------------------------------------------------ A* fast collision test
// largest double T with sqrt_rn(T) <= R: then `np.linalg.norm(d) <= R`  <=>  fma(dy,dy,dx*dx) <= T  exactly
// (sqrt is correctly rounded and monotone), which removes the square root from the per-sample tracker test.
__device__ __forceinline__ double d2d_sq_threshold(double R) {
    if (!(R >= 0.0)) return -1.0;
    double c = R * R;
    for (int it = 0; it < 8 && d2d_sqrt(c) > R; it++) c = __longlong_as_double(__double_as_longlong(c) - 1);
    for (int it = 0; it < 8; it++) {
        const double up = __longlong_as_double(__double_as_longlong(c) + 1);
        if (d2d_sqrt(up) <= R) c = up; else break;
    }
    return c;
}

// Planner.is_free (traj_planner.py:28-59) for the A* samples, whose coordinates are integer-valued doubles
// (np.around, traj_planner.py:181): cells by integer arithmetic.  trk = [mu0, mu1, mu2, mu3, radius, T] per tracker.
__device__ __forceinline__ bool d2d_is_free_int(const DevP &P, const uint8_t *bel, double px, double py, double t,
                                                const double *trk6, int nact) {
    if (px != px || py != py) return false;
    if (!(fabs(px) < 1e6 && fabs(py) < 1e6)) return false;       // far outside the map: every probe returns OCCUPIED
    const int x = (int)px, y = (int)py, sd = (int)(P.drone_r + 10.0), w = (int)P.map_w, h = (int)P.map_h;
    if (d2d_belief_probe_int(bel, x - sd, y, w, h) == 1) return false;
    if (d2d_belief_probe_int(bel, x, y, w, h) == 1) return false;
    if (d2d_belief_probe_int(bel, x + sd, y, w, h) == 1) return false;
    if (d2d_belief_probe_int(bel, x, y - sd, w, h) == 1) return false;
    if (d2d_belief_probe_int(bel, x, y + sd, w, h) == 1) return false;
#pragma unroll 1
    for (int k = 0; k < nact; k++) {
        const double *m = trk6 + 6 * k;
        const double ex = m[0] + t * m[2], ey = m[1] + t * m[3];   // estimate_pos utils.py:220-223
        const double ddx = px - ex, ddy = py - ey;
        if (D2D_FMA(ddy, ddy, ddx * ddx) <= m[5]) return false;    // norm(...) <= drone_r + radius + 5 + var_cam
    }
    return true;
}

// ------------------------------------------------------------------------------------------ A* kernel
// One block per planning env (persistent over the compacted list).  Per-search hot state lives in shared memory when
// it fits (8x8 primitives: open-set totals 51 KB + 32-bit-key hash 48 KB, two blocks per SM): the argmin over the open set, the dict
// lookups and the replace-if-cheaper test never leave the SM.  Cold node fields (position, velocity, parent, action)
// are SoA in an HBM workspace that stays L2-resident.  Collision checks are spread over (primitive, sample) pairs.
#define D2D_PLAN_SMEM_NODES 6400       // >= 99 * 64 + 8
#define D2D_PLAN_SMEM_HASH 8192
#define D2D_HASH32_EMPTY 0xFFFFFFFFu

struct PlanHot {                        // pointers into shared memory (fast path) or the HBM workspace (fallback)
    double *open_total;                 // +inf once the node is closed
    double *cost;
    uint32_t *hkeys32; uint16_t *hvals16;          // fast path
    unsigned long long *hkeys64; int *hvals32;     // fallback
    int hcap;
    bool fast;
};

__host__ __device__ inline size_t d2d_plan_smem_bytes(int NP, int n_u, int n_samp) {
    size_t b = D2D_BELIEF_STRIDE + (size_t)NP * 6 * 8 + 256 + (size_t)n_u * n_u;   // belief, trackers, scratch, prim_ok
    if (d2d_plan_cap(n_u) <= D2D_PLAN_SMEM_NODES)
        b += (size_t)D2D_PLAN_SMEM_NODES * 8 + (size_t)D2D_PLAN_SMEM_HASH * 6;
    return (b + 15) / 16 * 16;
}

// dict lookup / insert; returns the node index stored for the key (>= 0) or -1 after inserting a fresh slot (*slot)
__device__ __forceinline__ int d2d_plan_find_or_insert(const PlanHot &h, const PlanWs &w, double px, double py, double vx,
                                                       double vy, int *slot) {
    if (h.fast) {
        const uint32_t key = d2d_node_key32(px, py, vx, vy);
        int s = (int)((key * 2654435761u) >> 19) & (h.hcap - 1);
        for (;;) {
            const uint32_t cur = h.hkeys32[s];
            if (cur == key) { *slot = s; return h.hvals16[s]; }
            if (cur == D2D_HASH32_EMPTY) {
                const uint32_t old = atomicCAS(&h.hkeys32[s], D2D_HASH32_EMPTY, key);
                if (old == D2D_HASH32_EMPTY) { *slot = s; return -1; }
                if (old == key) { *slot = s; return h.hvals16[s]; }
            }
            s = (s + 1) & (h.hcap - 1);
        }
    } else {
        bool existed;
        const int s = d2d_hash_find_or_insert(w, d2d_node_key(px, py, vx, vy), &existed);
        *slot = s;
        return existed ? w.hvals[s] : -1;
    }
}

#define D2D_PLAN_THREADS2 256
// mode 0: the step's planning list (plan_list); mode 1: the searches d2d_plan_small_kernel abandoned (plan_over)
__global__ void __launch_bounds__(D2D_PLAN_THREADS2) d2d_plan_kernel(const DevP P, const int mode) {
    extern __shared__ __align__(16) unsigned char psm[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = T >> 5;
    const int nu = P.n_u, nprim = nu * nu, nsamp = P.n_samp;
    uint8_t *bel = psm;                                         // [2560]
    double *trk = (double *)(psm + D2D_BELIEF_STRIDE);          // [NP][6]
    double *red_v = trk + (size_t)P.NP * 6;                     // [8] warp partials
    int *red_i = (int *)(red_v + 8);                            // [8]
    int *sh = red_i + 8;                                        // [8] cur, n_nodes, n_open, -, nact, -, -, ticket
    int *wsum = sh + 8;                                         // [8] warp prefix
    uint8_t *prim_ok = (uint8_t *)(wsum + 8);                   // [nprim]
    const PlanWs w = d2d_plan_carve(P.plan_ws + (size_t)blockIdx.x * d2d_plan_workspace_bytes(P.n_u), P.n_u);
    PlanHot h;
    h.fast = d2d_plan_cap(nu) <= D2D_PLAN_SMEM_NODES && P.max_speed < 60.0;
    if (h.fast) {
        unsigned char *q = psm + ((size_t)D2D_BELIEF_STRIDE + (size_t)P.NP * 48 + 256 + nprim + 15) / 16 * 16;
        h.open_total = (double *)q; q += (size_t)D2D_PLAN_SMEM_NODES * 8;
        h.cost = w.cost;                                        // only read on dict hits: stays in the L2-resident workspace
        h.hkeys32 = (uint32_t *)q; q += (size_t)D2D_PLAN_SMEM_HASH * 4;
        h.hvals16 = (uint16_t *)q;
        h.hkeys64 = nullptr; h.hvals32 = nullptr; h.hcap = D2D_PLAN_SMEM_HASH;
    } else {
        h.open_total = w.open_total; h.cost = w.cost; h.hkeys32 = nullptr; h.hvals16 = nullptr;
        h.hkeys64 = w.hkeys; h.hvals32 = w.hvals; h.hcap = w.hcap;
    }
    const int *list = mode ? P.plan_over : P.plan_list;
    int *ticket = mode ? &P.plan_over[P.B + 1] : &P.plan_list[P.B + 5];
    const int count = mode ? min(P.plan_over[P.B], P.B)
                           : min(P.plan_list[P.B + (P.use_parity ? 1 + P.plan_list[P.B + 4] : 0)], P.B);

    // Searches differ widely in length (1 .. 99 expansions), so the list entries are handed out dynamically: a block takes
    // the next one when it is done (ticket counter plan_list[B+5], zeroed by the step kernel that fills the list).  With
    // the static grid-stride assignment the SMs were busy 48 % of the kernel.  Which block runs a search does not matter:
    // every search is self-contained in the block's own workspace.
    for (;;) {
        __syncthreads();
        if (tid == 0) sh[7] = atomicAdd(ticket, 1);                     // sh[7]: this block's ticket
        __syncthreads();
        const int li = sh[7];
        if (li >= count) break;
        const int e = list[li];
        for (int o = tid; o < D2D_BELIEF_STRIDE / 4; o += T)
            ((uint32_t *)bel)[o] = ((const uint32_t *)(P.belief + (size_t)e * D2D_BELIEF_STRIDE))[o];
        if (h.fast) {
            for (int o = tid; o < h.hcap; o += T) h.hkeys32[o] = D2D_HASH32_EMPTY;
        } else {
            for (int o = tid; o < w.hcap; o += T) { w.hkeys[o] = D2D_HASH_EMPTY; w.hvals[o] = -1; }
        }
        if (tid == 0) sh[4] = 0;
        __syncthreads();
#pragma unroll 1
        for (int k = tid; k < P.N; k += T) {        // active trackers + the exact squared clearance threshold
            const size_t g = (size_t)e * P.NP + k;
            if (P.trk_active[g]) {
                const int slot = atomicAdd(&sh[4], 1);
                const double *mu = P.trk_mu + g * 4;
                double *d = trk + 6 * slot;
                const double rad = P.trk_radius[g];
                d[0] = mu[0]; d[1] = mu[1]; d[2] = mu[2]; d[3] = mu[3]; d[4] = rad;
                d[5] = d2d_sq_threshold(P.drone_r + rad + 5.0 + P.var_cam);
            }
        }
        const double tx = P.rec[e].tgx, ty = P.rec[e].tgy;
        if (tid == 0) {   // start node (traj_planner.py:136-146)
            const double x = P.rec[e].px, y = P.rec[e].py, vx = P.rec[e].vx, vy = P.rec[e].vy;
            w.px[0] = x; w.py[0] = y; w.vx[0] = vx; w.vy[0] = vy; h.cost[0] = 0.0;
            h.open_total[0] = d2d_node_total(0.0, x, y, vx, vy, tx, ty);
            w.parent[0] = -1; w.itr[0] = 0; w.act[0] = 0;
            int slot;
            d2d_plan_find_or_insert(h, w, x, y, vx, vy, &slot);
            if (h.fast) h.hvals16[slot] = 0; else w.hvals[slot] = 0;
            sh[1] = 1; sh[2] = 1;
        }
        __syncthreads();
        const int nact = sh[4];
        int goal = -1;
        bool success = false;
        // loop-invariant per-thread work items: up to two (primitive, sample) pairs and one successor primitive
        const int nitems = nprim * nsamp;
        const bool two_items = nitems <= 2 * T;
        int it_p[2] = {0, 0};
        double it_xh[2] = {0, 0}, it_yh[2] = {0, 0}, it_t[2] = {0, 0}, it_t2[2] = {0, 0};
        bool it_ok[2] = {false, false};
        if (two_items) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int q = tid + j * T;
                if (q < nitems) {
                    const int pidx = q / nsamp, sI = q - pidx * nsamp, ia = pidx / nu, ib = pidx - ia * nu;
                    it_ok[j] = true; it_p[j] = pidx;
                    it_xh[j] = P.tab->u_space[ia] / 2.0; it_yh[j] = P.tab->u_space[ib] / 2.0;
                    it_t[j] = P.tab->t_samp[sI]; it_t2[j] = P.tab->t_samp2[sI];
                }
            }
        }
        // norm([nvx, nvy]) < max_speed  <=>  fma(nvy, nvy, nvx*nvx) <= speed_thr   (exact, no sqrt per item)
        const double speed_thr = d2d_sq_threshold(__longlong_as_double(__double_as_longlong(P.max_speed) - 1));
        int n_nodes = 1, n_open = 1;                     // replicated in every thread (all take the same decisions)
        for (int itr = 1;; itr++) {
            if (n_open == 0 || itr >= 100) break;                // traj_planner.py:149
            // ---- first minimal total_cost in insertion order (min() over a dict, :155-158)
            double bv = INFINITY;
            int bi = 0x7fffffff;
            for (int i = tid; i < n_nodes; i += T) {
                const double v = h.open_total[i];
                if (v < bv) { bv = v; bi = i; }
            }
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, bv, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; }
            for (int q = tid; q < nprim; q += T) prim_ok[q] = 1;
            __syncthreads();
            bv = red_v[0]; bi = red_i[0];
            for (int q = 1; q < NW; q++)
                if (red_v[q] < bv || (red_v[q] == bv && red_i[q] < bi)) { bv = red_v[q]; bi = red_i[q]; }
            const int cur = bi;
            if (cur == 0x7fffffff) break;                        // only non-finite costs left (cannot happen)
            const double cpx = w.px[cur], cpy = w.py[cur], cvx = w.vx[cur], cvy = w.vy[cur], ccost = h.cost[cur];
            const int citr = w.itr[cur];
            if (d2d_norm2(cpx - tx, cpy - ty) <= 10.0) {         // :160
                goal = cur; success = true; break;
            }
            if (tid == 0) h.open_total[cur] = INFINITY;          // open -> closed (:167-170); visible after the next sync
            n_open -= 1;
            // ---- collision checks of all (primitive, sample) pairs (:174-185)
            const double gt0 = (double)(citr * 2);
            if (two_items) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (!it_ok[j]) continue;
                    const double nvx = 1.0 * cvx + 4.0 * it_xh[j], nvy = 1.0 * cvy + 4.0 * it_yh[j];
                    bool ok = D2D_FMA(nvy, nvy, nvx * nvx) <= speed_thr;     // :176
                    if (ok && prim_ok[it_p[j]]) {
                        const double qx = rint(D2D_FMA(it_t2[j], it_xh[j], 1.0 * cpx + it_t[j] * cvx));
                        const double qy = rint(D2D_FMA(it_t2[j], it_yh[j], 1.0 * cpy + it_t[j] * cvy));
                        ok = d2d_is_free_int(P, bel, qx, qy, it_t[j] + gt0, trk, nact);
                    }
                    if (!ok) prim_ok[it_p[j]] = 0;
                }
            } else {
                for (int q = tid; q < nitems; q += T) {
                    const int pidx = q / nsamp, sI = q - pidx * nsamp;
                    const int ia = pidx / nu, ib = pidx - ia * nu;
                    const double xh = P.tab->u_space[ia] / 2.0, yh = P.tab->u_space[ib] / 2.0;
                    const double nvx = 1.0 * cvx + 4.0 * xh, nvy = 1.0 * cvy + 4.0 * yh;
                    bool ok = D2D_FMA(nvy, nvy, nvx * nvx) <= speed_thr;
                    if (ok && prim_ok[pidx]) {
                        const double t = P.tab->t_samp[sI], t2 = P.tab->t_samp2[sI];
                        const double qx = rint(D2D_FMA(t2, xh, 1.0 * cpx + t * cvx));
                        const double qy = rint(D2D_FMA(t2, yh, 1.0 * cpy + t * cvy));
                        ok = d2d_is_free_int(P, bel, qx, qy, t + gt0, trk, nact);
                    }
                    if (!ok) prim_ok[pidx] = 0;
                }
            }
            __syncthreads();
            // ---- successors in (x_acc, y_acc) loop order, T primitives at a time (:187-206)
            for (int base = 0; base < nprim; base += T) {
                const int pidx = base + tid;
                const bool ok = pidx < nprim && prim_ok[pidx];
                double nvx = 0, nvy = 0, spx = 0, spy = 0, scost = 0;
                int slot = -1, exist_idx = -1;
                bool is_new = false;
                if (ok) {
                    const double xa = P.tab->u_space[pidx / nu], ya = P.tab->u_space[pidx - (pidx / nu) * nu];
                    nvx = 1.0 * cvx + 4.0 * (xa / 2.0);
                    nvy = 1.0 * cvy + 4.0 * (ya / 2.0);
                    spx = rint((1.0 * cpx + 2.0 * cvx) + 4.0 * (xa / 2.0));   // :188
                    spy = rint((1.0 * cpy + 2.0 * cvy) + 4.0 * (ya / 2.0));
                    scost = ccost + (xa * xa + ya * ya) / 100.0 + 10.0;      // :190
                    exist_idx = d2d_plan_find_or_insert(h, w, spx, spy, nvx, nvy, &slot);
                    is_new = exist_idx < 0;
                }
                // ordered slot assignment for the new nodes of this chunk (insertion order == primitive order)
                const unsigned bal = __ballot_sync(0xffffffffu, is_new);
                const int wrank = __popc(bal & ((1u << lane) - 1u));
                if (lane == 0) wsum[wid] = __popc(bal);
                __syncthreads();
                int before = 0, tot_new = 0;
                for (int q = 0; q < NW; q++) {
                    if (q < wid) before += wsum[q];
                    tot_new += wsum[q];
                }
                int idx = -1;
                if (is_new) {
                    idx = n_nodes + before + wrank;
                    if (h.fast) h.hvals16[slot] = (uint16_t)idx; else w.hvals[slot] = idx;
                } else if (ok) {
                    // in closed_set -> skip; in open_set -> replace if cheaper, keeping the dict slot (:197-206)
                    if (h.open_total[exist_idx] != INFINITY && h.cost[exist_idx] > scost) idx = exist_idx;
                }
                if (idx >= 0 && idx < w.cap) {
                    w.px[idx] = spx; w.py[idx] = spy; w.vx[idx] = nvx; w.vy[idx] = nvy; h.cost[idx] = scost;
                    h.open_total[idx] = d2d_node_total(scost, spx, spy, nvx, nvy, tx, ty);
                    w.parent[idx] = cur; w.itr[idx] = citr + 1; w.act[idx] = pidx;
                }
                n_nodes += tot_new; n_open += tot_new;
                __syncthreads();      // new nodes / wsum reuse visible before the next chunk or the next argmin
            }
        }
        __syncthreads();
        // ---- result: segments from the start outwards (the reference walks parents and reverses, :208-217)
        if (tid == 0) {
            atomicAdd(&P.stats[D2D_STAT_PLANS], 1ull);
            if (success) {
                int depth = 0;
                for (int c2 = goal; c2 != 0; c2 = w.parent[c2]) depth++;
                int seg = depth;
                for (int c2 = goal; c2 != 0; c2 = w.parent[c2]) {
                    seg--;
                    const int par = w.parent[c2], pidx = w.act[c2];
                    double *cf = P.traj_coeff + ((size_t)e * D2D_MAX_SEGMENTS + seg) * 6;
                    cf[0] = w.px[par]; cf[1] = w.vx[par]; cf[2] = P.tab->u_space[pidx / nu] / 2.0;
                    cf[3] = w.py[par]; cf[4] = w.vy[par]; cf[5] = P.tab->u_space[pidx - (pidx / nu) * nu] / 2.0;
                }
                P.rec[e].nseg = depth; P.rec[e].cursor = 0; P.plan_ok[e] = 1;
            } else {
                P.rec[e].nseg = 0; P.rec[e].cursor = 0; P.plan_ok[e] = 0;
                atomicAdd(&P.stats[D2D_STAT_PLAN_FAILURES], 1ull);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ A* kernel, small footprint
// The same search with EVERYTHING in shared memory, sized for what the searches actually need: on BASELINE configs 3 and 4 no
// search of 5.6e4 sampled ones (instrumented oracle run) holds more than 384 nodes, while d2d_plan_kernel reserves room for the
// worst case (99 expansions x 64 primitives = 6336 nodes, 104 KB of shared memory: two searches per SM, 16 warps, schedulers
// issuing 30 % of the time, every expansion paying two L2 round trips for the cold node fields).  Here a search holds
// D2D_PS_NODES nodes (all fields) and a D2D_PS_HASH-slot dict in ~37 KB, runs on 4 warps, and six of them share an SM.  A
// search that would need a 513th node is abandoned and appended to the overflow list, which d2d_plan_kernel (mode 1) works
// off right afterwards -- every search is self-contained, so who runs it does not matter.
//   * an expansion is a latency chain and the failed 99-expansion searches set the length of the kernel, so the chain is short:
//     two block barriers per expansion; every warp first lists the primitives that pass the speed test (typically 10-15 of the
//     64 -- the others used to occupy lanes that did nothing), then one (feasible primitive, sample) pair per thread with the five
//     belief probes in flight together; warp 0 alone inserts the successors and picks the next node;
//   * trackers that cannot come near ANY sample of this expansion are dropped from the test per warp: a primitive that passes
//     the speed test (:176) has |v(t)| <= max(|v(0)|, v_max) (|v| is convex along it), so its samples stay within
//     2 max(|v0|, v_max) + 0.71 (rounding) of the node, a tracker estimate moves 2 |v_trk| in the same 2 s, and the test is
//     only skipped when the clearance at the node's time exceeds all that by a further whole unit.
#define D2D_PS_THREADS 128
#ifndef D2D_PS_NODES             // -DD2D_PS_NODES=40 -DD2D_PS_HASH=128: test build in which most searches overflow
#define D2D_PS_NODES 512
#define D2D_PS_HASH 1024
#endif
#define D2D_PS_MINB 5

__host__ __device__ inline bool d2d_plan_small_ok(int n_u, double max_speed) {
    return n_u * n_u <= 64 && n_u <= 8 && max_speed < 60.0;      // key32 needs |v| < 64
}
__host__ __device__ inline size_t d2d_plan_small_smem_bytes(int NP) {
    const size_t NPe = (size_t)(NP + 7) / 8 * 8;
    size_t b = D2D_BELIEF_STRIDE + NPe * 6 * 8 + (size_t)D2D_PS_NODES * 6 * 8 + (8 + 2 * D2D_MAX_SAMP + 3 * 64 + 2 + 6 * 64 + 32) * 8   // doubles (+ 64 keys)
             + (size_t)D2D_PS_HASH * 4 + 8 * 4                                                                    // words
             + (size_t)D2D_PS_HASH * 2 + (size_t)D2D_PS_NODES * 2 + (D2D_PS_THREADS / 32) * NPe * 2               // halves
             + (size_t)D2D_PS_NODES * 2 + 64 + (D2D_PS_THREADS / 32) * 64;                                        // bytes
    return (b + 15) / 16 * 16;
}

// first minimum over a warp of (v, i) pairs with v >= +0.0 (or +inf): the bit patterns of non-negative doubles order like
// unsigned integers, so three redux.sync steps (high word, low word, index) replace five shuffle rounds on the latency chain.
// Every lane returns the winner.
__device__ __forceinline__ void d2d_warp_first_min(double &v, int &i) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const bool mine = hi == mh && lo == ml;
    i = (int)__reduce_min_sync(0xffffffffu, mine ? (unsigned)i : 0x7fffffffu);
    v = __hiloint2double((int)mh, (int)ml);
}

__global__ void __launch_bounds__(D2D_PS_THREADS, D2D_PS_MINB) d2d_plan_small_kernel(const DevP P) {
    extern __shared__ __align__(16) unsigned char psm[];
    constexpr int T = D2D_PS_THREADS, NW = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nu = P.n_u, nprim = nu * nu, nsamp = P.n_samp;
    const int NPe = (P.NP + 7) / 8 * 8;
    unsigned char *q8 = psm;
    uint8_t *bel = q8; q8 += D2D_BELIEF_STRIDE;                          // [2560]
    double *trk = (double *)q8; q8 += (size_t)NPe * 48;                  // [NP][6]
    double *n_px = (double *)q8; q8 += (size_t)D2D_PS_NODES * 48;        // node fields, SoA
    double *n_py = n_px + D2D_PS_NODES, *n_vx = n_py + D2D_PS_NODES, *n_vy = n_vx + D2D_PS_NODES;
    double *n_cost = n_vy + D2D_PS_NODES, *n_open_total = n_cost + D2D_PS_NODES;   // open_total: +inf once closed
    double *uh = (double *)q8; q8 += 8 * 8;                              // u_space / 2
    double *ts = (double *)q8; q8 += (size_t)D2D_MAX_SAMP * 8;
    double *ts2 = (double *)q8; q8 += (size_t)D2D_MAX_SAMP * 8;
    double *hx = (double *)q8; q8 += 64 * 8;                             // per primitive: x_acc / 2, y_acc / 2,
    double *hy = (double *)q8; q8 += 64 * 8;
    double *pc = (double *)q8; q8 += 64 * 8;                             // (x_acc^2 + y_acc^2) / 100  (:190)
    double *cand_v = (double *)q8; q8 += 2 * 8;                          // next-node candidates of warp 0 / warp 1
    double *su = (double *)q8; q8 += 6 * 64 * 8;                         // successor per feasible primitive (by rank): px py vx vy cost total
    uint32_t *su_key = (uint32_t *)q8; q8 += 64 * 4;                     // ... and its dict key
    uint32_t *hkeys = (uint32_t *)q8; q8 += (size_t)D2D_PS_HASH * 4;
    int *sh = (int *)q8; q8 += 8 * 4;                                    // [1] nodes, [2] open, [3] overflow, [4] nact, [5], [6] candidate nodes, [7] ticket
    uint16_t *hvals = (uint16_t *)q8; q8 += (size_t)D2D_PS_HASH * 2;
    uint16_t *n_parent = (uint16_t *)q8; q8 += (size_t)D2D_PS_NODES * 2;
    uint16_t *live = (uint16_t *)q8 + (size_t)wid * NPe; q8 += (size_t)NW * NPe * 2;   // this warp's tracker list
    uint8_t *n_itr = q8; q8 += D2D_PS_NODES;
    uint8_t *n_act = q8; q8 += D2D_PS_NODES;
    uint8_t *vok = q8; q8 += 64;                                         // [64] verdict per speed-feasible primitive (by rank)
    uint8_t *vlist = q8 + (size_t)wid * 64;                              // [NW][64] this warp's list of speed-feasible primitives

    for (int i = tid; i < nu; i += T) uh[i] = P.tab->u_space[i] / 2.0;
    for (int i = tid; i < nsamp; i += T) { ts[i] = P.tab->t_samp[i]; ts2[i] = P.tab->t_samp2[i]; }
    for (int pp = tid; pp < nprim; pp += T) {
        const int ia = d2d_div_small(pp, nu), ib = pp - ia * nu;
        const double xa = P.tab->u_space[ia], ya = P.tab->u_space[ib];
        hx[pp] = xa / 2.0; hy[pp] = ya / 2.0; pc[pp] = (xa * xa + ya * ya) / 100.0;
    }
    const int count = min(P.plan_list[P.B + (P.use_parity ? 1 + P.plan_list[P.B + 4] : 0)], P.B);
    // norm([nvx, nvy]) < max_speed  <=>  fma(nvy, nvy, nvx*nvx) <= speed_thr   (exact, no sqrt per item)
    const double speed_thr = d2d_sq_threshold(__longlong_as_double(__double_as_longlong(P.max_speed) - 1));
    const int p_sd = (int)(P.drone_r + 10.0), p_w = (int)P.map_w, p_h = (int)P.map_h;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh[7] = atomicAdd(&P.plan_list[P.B + 5], 1);      // next list entry (see d2d_plan_kernel)
        __syncthreads();
        const int li = sh[7];
        if (li >= count) break;
        const int e = P.plan_list[li];
#ifdef D2D_PLAN_PROF
        unsigned long long prof_t0 = 0;
        if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_t0));
#endif
        for (int o = tid; o < D2D_BELIEF_STRIDE / 4; o += T)
            ((uint32_t *)bel)[o] = ((const uint32_t *)(P.belief + (size_t)e * D2D_BELIEF_STRIDE))[o];
        for (int o = tid; o < D2D_PS_HASH; o += T) hkeys[o] = D2D_HASH32_EMPTY;
        if (tid == 0) sh[4] = 0;
        __syncthreads();
#pragma unroll 1
        for (int k = tid; k < P.N; k += T) {        // active trackers + the exact squared clearance threshold
            const size_t g = (size_t)e * P.NP + k;
            if (P.trk_active[g]) {
                const int slot = atomicAdd(&sh[4], 1);
                const double *mu = P.trk_mu + g * 4;
                double *d = trk + 6 * slot;
                const double rad = P.trk_radius[g];
                const double clr = P.drone_r + rad + 5.0 + P.var_cam;
                d[0] = mu[0]; d[1] = mu[1]; d[2] = mu[2]; d[3] = mu[3];
                d[4] = clr + 2.0 * d2d_norm2(mu[2], mu[3]) + 1.0;       // + how far the estimate moves in 2 s + a whole unit of margin
                d[5] = d2d_sq_threshold(clr);
            }
        }
        const double tx = P.rec[e].tgx, ty = P.rec[e].tgy;
        if (tid == 0) {   // start node (traj_planner.py:136-146)
            const double x = P.rec[e].px, y = P.rec[e].py, vx = P.rec[e].vx, vy = P.rec[e].vy;
            n_px[0] = x; n_py[0] = y; n_vx[0] = vx; n_vy[0] = vy; n_cost[0] = 0.0;
            n_open_total[0] = d2d_node_total(0.0, x, y, vx, vy, tx, ty);
            n_parent[0] = 0; n_itr[0] = 0; n_act[0] = 0;
            const uint32_t key = d2d_node_key32i(x, y, vx, vy);
            const int s0 = (int)((key * 2654435761u) >> 19) & (D2D_PS_HASH - 1);
            hkeys[s0] = key; hvals[s0] = 0;
            sh[1] = 1; sh[2] = 1; sh[3] = 0;                     // nodes, open nodes, overflow
            cand_v[0] = n_open_total[0]; sh[5] = 0; cand_v[1] = INFINITY; sh[6] = 0x7fffffff;
        }
        if (tid < 64) vok[tid] = 1;
        __syncthreads();
        const int nact = sh[4];
        int goal = -1;
        bool success = false, overflow = false;
        int n_nodes = 1, n_open = 1;                     // kept by warp 0, published through sh[1], sh[2] at the end of an expansion
        int itr = 1;
        int pend_i0 = -1, pend_i1 = -1;                  // warp 0: open nodes this lane has made cheaper, totals not yet stored
        double pend_t0 = 0.0, pend_t1 = 0.0;
        // An expansion is a latency chain, and the failed 99-expansion searches set the length of the kernel, so the chain is
        // kept short: two block barriers per expansion.  Between (B) and (A) every warp tests samples; between (A) and (B) warp 0
        // closes the node and inserts the successors while warp 1 scans the older nodes for the next one to expand.
        for (;; itr++) {
            if (n_open == 0 || itr >= 100) break;                // traj_planner.py:149
            // first minimal total_cost in insertion order (min() over a dict, :155-158): the better of the two candidates
            const int cur = (cand_v[1] < cand_v[0] || (cand_v[1] == cand_v[0] && sh[6] < sh[5])) ? sh[6] : sh[5];
            if (cur == 0x7fffffff) break;                        // only non-finite costs left (cannot happen)
            const double cpx = n_px[cur], cpy = n_py[cur], cvx = n_vx[cur], cvy = n_vy[cur];
            const int citr = n_itr[cur];
            if (d2d_norm2_le(cpx - tx, cpy - ty, 10.0)) {        // :160
                goal = cur; success = true; break;
            }
            const double gt0 = (double)(citr * 2);
            // ---- per warp (no barrier): the primitives that pass the speed test (:176), in primitive order.  Typically 10-15
            //      of the 64 do; only their samples are tested, one (primitive, sample) pair per thread.
            int nv = 0;
#pragma unroll 2
            for (int base = 0; base < nprim; base += 32) {
                const int pp = base + lane;
                bool pass = false;
                if (pp < nprim) {
                    const double nvx = 1.0 * cvx + 4.0 * hx[pp], nvy = 1.0 * cvy + 4.0 * hy[pp];
                    pass = D2D_FMA(nvy, nvy, nvx * nvx) <= speed_thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) vlist[nv + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)pp;
                nv += __popc(bal);
            }
            // ---- per warp: the trackers that can come near a sample of this expansion (see the kernel's header)
            int nlive = 0;
            {
                const double reach = 2.0 * fmax(P.max_speed, fabs(cvx) + fabs(cvy)) + 0.71;     // |v0| <= |v0x| + |v0y|: no square root
#pragma unroll 1
                for (int base = 0; base < nact; base += 32) {
                    const int k = base + lane;
                    bool keep = false;
                    if (k < nact) {
                        const double *m = trk + 6 * k;
                        const double ddx = cpx - (m[0] + gt0 * m[2]), ddy = cpy - (m[1] + gt0 * m[3]);
                        const double lim = reach + m[4];             // m[4] = clearance + 2 |v_trk| + 1
                        keep = !(D2D_FMA(ddy, ddy, ddx * ddx) > lim * lim);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) live[nlive + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)k;
                    nlive += __popc(bal);
                }
            }
            __syncwarp();
            // ---- collision checks of the (primitive, sample) pairs (:174-185); vok[r]: primitive vlist[r] is still free
            const int nit = nv * nsamp;
#pragma unroll 1
            for (int q = tid; q < nit; q += T) {
                const int sI = d2d_div_small(q, nv), r = q - sI * nv;      // (the samples of a primitive run side by side: nothing to skip)
                const int pp = vlist[r];
                const double xh = hx[pp], yh = hy[pp];
                const double t = ts[sI], t2 = ts2[sI];
                const double qx = rint(D2D_FMA(t2, xh, 1.0 * cpx + t * cvx));
                const double qy = rint(D2D_FMA(t2, yh, 1.0 * cpy + t * cvy));
                const double tg = t + gt0;
                bool occ = !(fabs(qx) < 1e6 && fabs(qy) < 1e6);  // NaN or far outside the map: every probe reads OCCUPIED
                if (!occ) occ = d2d_probe5_occ(bel, (int)qx, (int)qy, p_sd, p_w, p_h) != 0;
#pragma unroll 1
                for (int jj = 0; jj < nlive && !occ; jj++) {
                    const double *m = trk + 6 * (int)live[jj];
                    const double ex = m[0] + tg * m[2], ey = m[1] + tg * m[3];   // estimate_pos utils.py:220-223
                    const double ddx = qx - ex, ddy = qy - ey;
                    occ = D2D_FMA(ddy, ddy, ddx * ddx) <= m[5];                  // norm(...) <= drone_r + radius + 5 + var_cam
                }
                if (occ) vok[r] = 0;
            }
            // ---- the successor of every feasible primitive (:187-195), computed by the last warp -- the one with the fewest samples
            //      to test -- whatever the verdicts will be, so that warp 0 only has the dict and the node arrays left to do
            if (wid == NW - 1) {
                const double ccost = n_cost[cur];
#pragma unroll 1
                for (int r = lane; r < nv; r += 32) {
                    const int pp = vlist[r];
                    const double xh = hx[pp], yh = hy[pp];                       // x_acc / 2, y_acc / 2
                    const double nvx = 1.0 * cvx + 4.0 * xh, nvy = 1.0 * cvy + 4.0 * yh;
                    const double spx = rint((1.0 * cpx + 2.0 * cvx) + 4.0 * xh); // :188
                    const double spy = rint((1.0 * cpy + 2.0 * cvy) + 4.0 * yh);
                    const double scost = ccost + pc[pp] + 10.0;                  // :190
                    su[r] = spx; su[64 + r] = spy; su[128 + r] = nvx; su[192 + r] = nvy; su[256 + r] = scost;
                    su[320 + r] = d2d_node_total(scost, spx, spy, nvx, nvy, tx, ty);
                    su_key[r] = d2d_node_key32i(spx, spy, nvx, nvy);
                }
            }
            __syncthreads();                                     // (A) verdicts of all samples and the successors in place
            if (wid == 0) {
                if (lane == 0) n_open_total[cur] = INFINITY;     // open -> closed (:167-170)
                n_open -= 1;
                double wbv = INFINITY;
                int wbi = 0x7fffffff;
                __syncwarp();
                // ---- successors in (x_acc, y_acc) loop order (:187-206): vlist is in that order, 32 per round
#pragma unroll 1
                for (int base = 0; base < nv; base += 32) {
                    const int r = base + lane;
                    const bool ok = r < nv && vok[r];
                    double scost = 0;
                    int slot = -1, exist_idx = -1;
                    bool is_new = false;
                    if (ok) {
                        scost = su[256 + r];
                        const uint32_t key = su_key[r];
                        int s = (int)((key * 2654435761u) >> 19) & (D2D_PS_HASH - 1);
                        for (;;) {
                            const uint32_t curk = hkeys[s];
                            if (curk == key) { exist_idx = hvals[s]; break; }
                            if (curk == D2D_HASH32_EMPTY) {
                                const uint32_t old = atomicCAS(&hkeys[s], D2D_HASH32_EMPTY, key);
                                if (old == D2D_HASH32_EMPTY) { is_new = true; break; }
                                if (old == key) { exist_idx = hvals[s]; break; }
                            }
                            s = (s + 1) & (D2D_PS_HASH - 1);
                        }
                        slot = s;
                    }
                    // ordered slot assignment for the new nodes (insertion order == primitive order)
                    const unsigned bal = __ballot_sync(0xffffffffu, is_new);
                    const int tot_new = __popc(bal);
                    if (n_nodes + tot_new > D2D_PS_NODES) { overflow = true; break; }    // same verdict in every lane
                    int idx = -1;
                    if (is_new) {
                        idx = n_nodes + __popc(bal & ((1u << lane) - 1u));
                        hvals[slot] = (uint16_t)idx;
                    } else if (ok) {
                        // in closed_set -> skip; in open_set -> replace if cheaper, keeping the dict slot (:197-206)
                        if (n_open_total[exist_idx] != INFINITY && n_cost[exist_idx] > scost) idx = exist_idx;
                    }
                    if (idx >= 0) {
                        const double tot = su[320 + r];
                        n_px[idx] = su[r]; n_py[idx] = su[64 + r]; n_vx[idx] = su[128 + r]; n_vy[idx] = su[192 + r]; n_cost[idx] = scost;
                        // a new node's total is stored now; the total of an older node made cheaper only after barrier (B), so that
                        // warp 1's scan of the older nodes reads values nobody is writing (warp 0's candidate carries the new one)
                        if (is_new) n_open_total[idx] = tot;
                        else if (base == 0) { pend_i0 = idx; pend_t0 = tot; }
                        else { pend_i1 = idx; pend_t1 = tot; }
                        if (tot < wbv || (tot == wbv && idx < wbi)) { wbv = tot; wbi = idx; }
                        n_parent[idx] = (uint16_t)cur; n_itr[idx] = (uint8_t)(citr + 1); n_act[idx] = vlist[r];
                    }
                    n_nodes += tot_new; n_open += tot_new;
                }
                vok[lane] = 1; vok[lane + 32] = 1;               // every primitive starts the next expansion as free
                // warp 0's candidate for the next node: the best of the nodes it has just written (new or made cheaper)
                d2d_warp_first_min(wbv, wbi);
                if (lane == 0) { cand_v[0] = wbv; sh[5] = wbi; sh[1] = n_nodes; sh[2] = n_open; sh[3] = overflow ? 1 : 0; }
            } else if (wid == 1) {
                // ---- meanwhile warp 1 scans the nodes that existed before this expansion (without the one being closed).  A node
                //      warp 0 is making cheaper right now still shows its old, higher total: its new total is in warp 0's
                //      candidate, so the better of the two candidates is the first minimum over the final values.
                double bv = INFINITY;
                int bi = 0x7fffffff;
#pragma unroll 2
                for (int i = lane; i < n_nodes; i += 32) {
                    if (i == cur) continue;                      // being closed by warp 0 right now
                    const double v = n_open_total[i];
                    if (v < bv) { bv = v; bi = i; }
                }
                d2d_warp_first_min(bv, bi);
                if (lane == 0) { cand_v[1] = bv; sh[6] = bi; }
            }
            __syncthreads();                                     // (B) next node, counts and the overflow flag published
            n_nodes = sh[1]; n_open = sh[2];
            if (sh[3]) { overflow = true; break; }
            if (wid == 0) {                                      // the deferred totals, well before the next barrier (A)
                if (pend_i0 >= 0) { n_open_total[pend_i0] = pend_t0; pend_i0 = -1; }
                if (pend_i1 >= 0) { n_open_total[pend_i1] = pend_t1; pend_i1 = -1; }
            }
        }
        __syncthreads();
        if (tid == 0) {
#ifdef D2D_PLAN_PROF
            // tools/plan_prof.py: per search start / end (ns), nodes, outcome, SM, ticket
            unsigned long long t1; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long *pr = P.prof + (size_t)e * 12;
            pr[0] = prof_t0; pr[1] = t1; pr[2] = (unsigned long long)n_nodes; pr[3] = success ? 1 : (overflow ? 2 : 0);
            pr[4] = smid; pr[5] = (unsigned long long)li; pr[6] = (unsigned long long)n_open; pr[7] = (unsigned long long)nact; pr[8] = (unsigned long long)itr;
#endif
            if (overflow) {
                P.plan_over[atomicAdd(&P.plan_over[P.B], 1)] = e;        // d2d_plan_kernel (mode 1) redoes this search
                atomicAdd(&P.stats[D2D_STAT_PLAN_OVERFLOWS], 1ull);
            } else {
                atomicAdd(&P.stats[D2D_STAT_PLANS], 1ull);
                if (success) {
                    // segments from the start outwards (the reference walks parents and reverses, :208-217)
                    int depth = 0;
                    for (int c2 = goal; c2 != 0; c2 = n_parent[c2]) depth++;
                    int seg = depth;
                    for (int c2 = goal; c2 != 0; c2 = n_parent[c2]) {
                        seg--;
                        const int par = n_parent[c2], pidx = n_act[c2];
                        double *cf = P.traj_coeff + ((size_t)e * D2D_MAX_SEGMENTS + seg) * 6;
                        cf[0] = n_px[par]; cf[1] = n_vx[par]; cf[2] = hx[pidx];
                        cf[3] = n_py[par]; cf[4] = n_vy[par]; cf[5] = hy[pidx];
                    }
                    P.rec[e].nseg = depth; P.rec[e].cursor = 0; P.plan_ok[e] = 1;
                } else {
                    P.rec[e].nseg = 0; P.rec[e].cursor = 0; P.plan_ok[e] = 0;
                    atomicAdd(&P.stats[D2D_STAT_PLAN_FAILURES], 1ull);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ post kernel
template <int E>
__global__ void __launch_bounds__(256) d2d_step_post_kernel(const DevP P, const double *__restrict__ actions) {
    extern __shared__ __align__(128) unsigned char smem[];
    const BlockCtx c = d2d_carve(smem, E, 1, 1);
    const int tid = threadIdx.x, T = blockDim.x;
    const int env0 = blockIdx.x * E;
    if (tid == 0) { d2d_mbar_init(c.mbar, 1); c.misc[0] = 0; c.misc[1] = 0; }
    __syncthreads();
    if (tid < E) {
        EnvS &s = c.S[tid];
        const int e = env0 + tid;
        s.valid = e < P.B;
        s.reset = 0; s.coll_agent = 0; s.done_now = 0;
        if (s.valid) {
            const uint4 *src = (const uint4 *)(P.rec + e);   // the whole record: it is stored back verbatim below
#pragma unroll
            for (int i = 0; i < 8; i++) ((uint4 *)&s)[i] = src[i];
            s.ix = 0; s.iy = 0; s.ncull = 0; s.ox_was_fresh = 0;
            s.arch_cnt = 0; s.arch_ts = 0; s.newly = 0;      // already applied by the pre kernel
            s.act_cnt = P.tmp_act_cnt[e]; s.act_ts = P.tmp_act_ts[e];
        }
    }
    __syncthreads();
    if (tid == 0) d2d_issue_bulk(P, c, env0, E, true);
    if (tid < E && c.S[tid].valid) {
        const int e = env0 + tid;
        const bool success = P.need_plan[e] ? (P.plan_ok[e] != 0) : true;
        d2d_leader_finish(P, c.S[tid], nullptr, e, actions[e], success);
    }
    __syncthreads();
    // Drone2D.is_collide, dynamic part (utils.py:773-776) against the drone's NEW position
    for (int w = tid; w < E * P.N; w += T) {
        const int i = w / P.N, k = w - i * P.N;
        EnvS &s = c.S[i];
        if (!s.valid) continue;
        const size_t g = (size_t)(env0 + i) * P.NP + k;
        const double2 pos = P.apos[g];
        if (d2d_norm2(pos.x - s.px, pos.y - s.py) < P.arad[g] + P.drone_r) atomicOr(&s.coll_agent, 1);
    }
    d2d_mbar_wait(c.mbar, 0);
    __syncthreads();
    if (tid < E && c.S[tid].valid) {
        EnvS &s = c.S[tid];
        const int e = env0 + tid;
        d2d_leader_flags(P, s, c.gt + (size_t)tid * D2D_GRID, e);
        d2d_store_env_scalars(P, s, e);
        if (s.done_now) c.misc[0] = 1;
    }
    if (tid == 0) {
        int nv = 0;
        for (int i = 0; i < E; i++) nv += c.S[i].valid;
        atomicAdd(&P.stats[D2D_STAT_ENV_STEPS], (unsigned long long)nv);
        if (blockIdx.x == 0) P.plan_list[P.B] = 0;   // reset the compaction counter for the next step
    }
    __syncthreads();
    d2d_phase_obs(P, c, env0, E, tid, T);
    if (c.misc[0]) d2d_phase_done_stats(P, c, E, tid, T);
}

// ------------------------------------------------------------------------------------------ Oxford gaze policy
// Oxford.plan (yaw_planner.py:81-127), one block per env.  np.sum over the 50x50 product array is reproduced in
// NumPy's pairwise order: the host supplies the leaf blocks (offset, length <= 128) of the recursion for n = 2500 and
// the post-order combine program; each leaf is summed by one thread with NumPy's 8-accumulator loop.
#define D2D_OX_MAX_LEAVES 64
struct OxProgram {
    int n_leaves, n_ops;
    int leaf_off[D2D_OX_MAX_LEAVES], leaf_len[D2D_OX_MAX_LEAVES];
    int ops[2 * D2D_OX_MAX_LEAVES];     // post-order: >= 0 push leaf, -1 add top two
    // the same program as a tree: node ids 0 .. n_leaves-1 are the leaves, internal nodes follow in post-order; internal
    // nodes sorted by level (1 = both children are leaves ...), level L occupies [lev_off[L-1], lev_off[L]) of node_*
    int n_levels, root;
    int lev_off[16];
    int node_id[D2D_OX_MAX_LEAVES], node_l[D2D_OX_MAX_LEAVES], node_r[D2D_OX_MAX_LEAVES];
};

__device__ __forceinline__ double d2d_np_leaf_sum(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; i++) res += a[i];
    return res;
}

// exact decision of `np.arccos(cc) <= view_angle` with cc = num / den (yaw_planner.py:78): cc >= c* and cc <= 1
// (arccos of cc > 1 is NaN -> False).  The IEEE division is only performed when num is within a few ulps of the two
// decision boundaries c* * den and den; everywhere else the comparison of the products decides (margins 8e-16 relative
// cover the rounding of the division and of the products).
__device__ __forceinline__ bool d2d_ox_wedge(double num, double den, double cstar) {
    const double t = cstar * den;
    const double slack = 8e-16;
    if (num > t + fabs(t) * slack && num < den - den * slack) return true;
    if (num < t - fabs(t) * slack) return false;
    const double cc = num / den;
    return cc >= cstar && cc <= 1.0;
}

__device__ __forceinline__ bool d2d_ox_visible(const DevP &P, int cell, double dx, double dy, double cs, double msn) {
    const int i = cell / D2D_GRID, j = cell - i * D2D_GRID;
    const double x = (double)i * P.scale, y = (double)j * P.scale;
    const double ex = dx - x, ey = dy - y;
    const double d2 = ex * ex + ey * ey;
    if (d2 <= 0.0) return true;
    if (!(d2 <= P.depth2)) return false;
    return d2d_ox_wedge((x - dx) * cs + (y - dy) * msn, D2D_SQRT(d2), P.ox_cos_thresh);
}

// Only cells whose corner lies within the view depth of the pose can be visible: rows/cols [lo, hi] of the grid.
__device__ __forceinline__ void d2d_ox_window(const DevP &P, double dx, int &lo, int &hi) {
    const double depth = D2D_SQRT(P.depth2);
    lo = (int)floor((dx - depth) * P.inv_scale) - 1;
    hi = (int)floor((dx + depth) * P.inv_scale) + 1;
    if (lo < 0) lo = 0;
    if (hi > D2D_GRID - 1) hi = D2D_GRID - 1;
}

#define D2D_OX_ROWS 21          // window rows (2 * depth / scale + slack) handled per pose
#define D2D_OX_SPAN (D2D_OX_ROWS * D2D_GRID + 2 * 128)   // materialised flattened range, extended to whole leaves
#define D2D_OX_WORDS ((D2D_OX_SPAN + 31) / 32)
#define D2D_OX_THREADS 128
// candidate yaw rates the kernel holds per env (Oxford.v_yaw_space has 6, yaw_planner.py:65); sized so that the kernel's static
// shared memory (21.5 KB) and registers (48) let 10 blocks = 40 warps share an SM instead of 8 = 32
#define D2D_OX_MAX_YAW 8
#ifndef D2D_OX_MINB
#define D2D_OX_MINB 10
#endif

// last_time_observed of a cell after policy call n (yaw_planner.py:95-97): 0 if visible at call n, else the += dt
// sequence continued from 0.0 (last seen at call s) or from the initial 5.0 (never seen), read from the exact table
__device__ __forceinline__ double d2d_ox_last_value(const DevP &P, int s, int n) {
    if (s == n && n > 0) return 0.0;
    int k = (s > 0) ? n - s : n;
    const double *tab = P.ox_tab + (s > 0 ? 0 : D2D_OX_TAB);
    if (k < D2D_OX_TAB) return tab[k];
    double v = tab[D2D_OX_TAB - 1];
    for (int q = D2D_OX_TAB - 1; q < k; q++) v = v + 1.0 * P.dt;       // beyond the table: continue the exact sequence
    return v;
}

__global__ void d2d_oxford_export_kernel(const DevP P, double *__restrict__ out) {
    const int e = blockIdx.x;
    if (e >= P.B) return;
    const int n = P.ox_calls[e];
    const uint16_t *seen = P.ox_seen + (size_t)e * D2D_OX_SEEN_STRIDE;
    for (int c = threadIdx.x; c < D2D_CELLS; c += blockDim.x) out[(size_t)e * D2D_CELLS + c] = d2d_ox_last_value(P, seen[c], n);
}

// Oxford.plan (yaw_planner.py:81-127), one block (4 warps) per env.
//  A. warp 0: sin/cos of the 6 candidate yaws + current yaw, first waypoint, leaf range, clears.  last_time_observed is
//     kept compact (call index of the last sighting per cell), so no per-call ageing pass over the 2500 cells exists.
//  B. swep_map from the remaining waypoints (largest waypoint index per cell: later assignments win, :87-89); cells of
//     the pose's depth window that are visible are reset to 0 (only those can be visible).
//  C. candidate scores np.sum(view_k * reward) (:120-125): view_k vanishes outside the depth window of the first
//     waypoint, so reward (:108-110) is materialised only on the flattened range of rows that window touches (extended
//     to whole leaves of NumPy's pairwise recursion) and each candidate gets a visibility BITMASK over that range.
//  D. leaf sums in NumPy's order -- 8 strided accumulators (one lane each; a lane only visits the set bits of its
//     residue class), combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail; leaves outside the range are exactly
//     +0.0 (x + 0.0 == x).
//  E. the recursion's combine as a level-parallel tree (same operand pairs), then the strict-< argmax.
// One env's Oxford.plan on the calling block (every thread of the block enters and leaves together).
__device__ __forceinline__ void d2d_oxford_env(const DevP &P, const OxProgram *__restrict__ prog, double *__restrict__ actions_out,
                                               const int e) {
    __shared__ double reward[D2D_OX_SPAN];
    __shared__ uint32_t vmask[D2D_OX_MAX_YAW][D2D_OX_WORDS];
    __shared__ int swep_w[D2D_OX_SPAN];                 // largest waypoint index per cell of the range (-1: none)
    __shared__ double leaf[D2D_OX_MAX_YAW][D2D_OX_MAX_LEAVES];
    __shared__ double cs_s[D2D_OX_MAX_YAW + 1], sn_s[D2D_OX_MAX_YAW + 1];
    __shared__ double wpt[2];
    __shared__ int geo[12];                             // r0 r1 q0 q1 lo hi l0 l1 i0 i1 j0 j1
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, wid = tid >> 5;
    // an env that reported done and will be re-initialised by its next step is seen by the policy as freshly reset
    // (the reference builds a new env + policy per episode, experiment.py:27-34)
    const bool fresh = P.rec[e].pending_reset || (P.auto_reset && P.done[e]);
    const double dx = fresh ? P.rec[e].p0x : P.rec[e].px, dy = fresh ? P.rec[e].p0y : P.rec[e].py;
    const double yaw = fresh ? P.rec[e].p0yaw : P.rec[e].yaw;
    const int len = fresh ? 0 : P.rec[e].nseg * P.n_way - P.rec[e].cursor, cursor = fresh ? 0 : P.rec[e].cursor;
    const int ny = P.n_yaw;
    uint16_t *seen = P.ox_seen + (size_t)e * D2D_OX_SEEN_STRIDE;
    const int ncall = (fresh ? 0 : P.ox_calls[e]) + 1;          // index of this policy call within the episode

    // ---- A: three independent pieces on three warps (warp 0 alone used to run them back to back -- ~2000 cycles of
    //      double-double sin/cos, then the DRAM round trip of the first waypoint -- with the other warps at the barrier)
    if (wid == 0) {
        if (lane <= ny) {
            // candidate yaws: Drone2D(..., yaw_i) stores yaw_i % 360 (utils.py:718); entry n_yaw is the current pose
            const double y = (lane < ny) ? d2d_pymod(yaw + P.tab->v_yaw_space[lane] * P.dt, 360.0) : yaw;
            double sn, cs;
            d2d_sincos(y * D2D_DEG2RAD, &sn, &cs);
            cs_s[lane] = cs; sn_s[lane] = -sn;    // vec_yaw = [cos, -sin] (yaw_planner.py:72)
        }
    } else if (wid == 1) {
        int r0 = 0, r1 = -1, q0 = 0, q1 = -1, i0, i1, j0, j1;
        d2d_ox_window(P, dx, i0, i1);
        d2d_ox_window(P, dy, j0, j1);
        double wx = 0, wy = 0;
        if (len > 0) {
            d2d_waypoint_pos(P, e, cursor, wx, wy);
            d2d_ox_window(P, wx, r0, r1);
            d2d_ox_window(P, wy, q0, q1);
            if (r1 - r0 + 1 > D2D_OX_ROWS) r1 = r0 + D2D_OX_ROWS - 1;       // cannot happen for depth <= 9 cells
#pragma unroll 1
            for (int l = lane; l < prog->n_leaves; l += 32) {
                const int b0 = r0 * D2D_GRID, b1 = (r1 + 1) * D2D_GRID - 1;
                const int off = prog->leaf_off[l], n = prog->leaf_len[l];
                if (b0 >= off && b0 < off + n) { geo[6] = l; geo[4] = off; }
                if (b1 >= off && b1 < off + n) { geo[7] = l; geo[5] = off + n; }
            }
        }
        if (lane == 0) {
            geo[0] = r0; geo[1] = r1; geo[2] = q0; geo[3] = q1; geo[8] = i0; geo[9] = i1; geo[10] = j0; geo[11] = j1;
            wpt[0] = wx; wpt[1] = wy;
        }
    } else {
#pragma unroll 1
        for (int o = tid - 64; o < D2D_OX_MAX_YAW * D2D_OX_WORDS; o += D2D_OX_THREADS - 64) (&vmask[0][0])[o] = 0u;
        if (fresh) {
            // first call of a new episode: nothing has been seen yet
#pragma unroll 1
            for (int o = tid - 64; o < D2D_OX_SEEN_STRIDE / 2; o += D2D_OX_THREADS - 64) ((uint32_t *)seen)[o] = 0u;
        }
    }
    __syncthreads();
    const int r0 = geo[0], r1 = geo[1], q0 = geo[2], q1 = geo[3], lo = geo[4], hi = geo[5], l0 = geo[6], l1 = geo[7];
    const int i0 = geo[8], i1 = geo[9], j0 = geo[10], j1 = geo[11];
    const double wx = wpt[0], wy = wpt[1];
    const int span = (len > 0) ? hi - lo : 0;
    // ---- B
#pragma unroll 1
    for (int o = tid; o < span; o += T) swep_w[o] = -1;
    {
        const int wi = i1 - i0 + 1, wj = j1 - j0 + 1;
#pragma unroll 1
        for (int q = tid; q < wi * wj; q += T) {
            const int qi = d2d_div_small(q, wj);
            const int i = i0 + qi, j = j0 + (q - qi * wj);
            const int c = i * D2D_GRID + j;
            if (d2d_ox_visible(P, c, dx, dy, cs_s[ny], sn_s[ny])) seen[c] = (uint16_t)ncall;
        }
    }
    if (len == 0) {                                                  // :117-118
        if (tid == 0) { actions_out[e] = 0.0; P.ox_calls[e] = ncall; if (fresh) P.rec[e].ox_fresh = 1; }
        return;
    }
    __syncthreads();
#pragma unroll 1
    for (int w = tid; w < len; w += T) {
        double x, y;
        d2d_waypoint_pos(P, e, cursor + w, x, y);
        const int ci = d2d_cell(x, P.scale, P.inv_scale), cj = d2d_cell(y, P.scale, P.inv_scale);
        if ((unsigned)ci < (unsigned)D2D_GRID && (unsigned)cj < (unsigned)D2D_GRID) {
            const int c = ci * D2D_GRID + cj;
            if (c >= lo && c < hi) atomicMax(&swep_w[c - lo], w);
        }
    }
    __syncthreads();
    // ---- C: window cells of the first waypoint only
    {
        const int wr = r1 - r0 + 1, wq = q1 - q0 + 1;
#pragma unroll 1
        for (int q = tid; q < wr * wq; q += T) {
            const int qi = d2d_div_small(q, wq);
            const int i = r0 + qi, j = q0 + (q - qi * wq);
            const int c = i * D2D_GRID + j, o = c - lo;
            const double x = (double)i * P.scale, y = (double)j * P.scale;
            const double ex = wx - x, ey = wy - y;
            const double d2 = ex * ex + ey * ey;
            const bool zero_d = d2 <= 0.0;
            if (!(zero_d || d2 <= P.depth2)) continue;
            const double lt = d2d_ox_last_value(P, seen[c], ncall);
            const double sw = swep_w[o] >= 0 ? (double)swep_w[o] * P.dt : 0.0;
            double r;
            if (sw > 0.0 && sw <= 3.0 && lt >= 0.5) r = 1000000.0;           // :108-110
            else if (sw > 3.0 && lt >= 0.5) r = 1000.0;
            else r = (lt > 1.0) ? 1.0 : lt;
            reward[o] = r;
            const double num_x = x - wx, num_y = y - wy, den = D2D_SQRT(d2);
#pragma unroll 1
            for (int k = 0; k < ny; k++) {
                const bool vis = zero_d || d2d_ox_wedge(num_x * cs_s[k] + num_y * sn_s[k], den, P.ox_cos_thresh);
                if (vis) atomicOr(&vmask[k][o >> 5], 1u << (o & 31));
            }
        }
    }
    __syncthreads();
    // ---- D: leaf sums: task = (candidate, leaf, accumulator lane u in 0..7)
    const int nleaf = l1 - l0 + 1;
#pragma unroll 1
    for (int qb = 0; qb < ny * nleaf * 8; qb += T) {
        const int q = qb + tid;
        const bool act = q < ny * nleaf * 8;
        const int u = q & 7, kl = q >> 3;
        const int k = act ? d2d_div_small(kl, nleaf) : 0, l = l0 + (act ? kl - k * nleaf : 0);
        const int off = prog->leaf_off[l], n = prog->leaf_len[l];
        const int nb = n - (n % 8);                     // elements covered by the 8 accumulators
        double r = 0.0;
        if (act && n >= 8) {
            // accumulator u sums elements off+u, off+8+u, ... in ascending order; zeros are skipped (x + 0.0 == x):
            // walk the words of the leaf and visit only set bits whose leaf-relative index is = u (mod 8)
            const int o0 = off - lo, o1 = o0 + nb;
            const uint32_t cls = 0x01010101u << ((u + o0) & 7);   // bit positions b with (32w + b - o0) = u (mod 8)
            bool first = true;
#pragma unroll 1
            for (int w = o0 >> 5; w <= (o1 - 1) >> 5; w++) {
                uint32_t m = vmask[k][w] & cls;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int o = (w << 5) + b;
                    if (o < o0 || o >= o1) continue;
                    if (first) { r = reward[o]; first = false; }     // r[j] = a[j] then += : same value as 0.0 + a
                    else r += reward[o];
                }
            }
        }
        // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) across the 8 lanes of the group
        double t = r + __shfl_down_sync(0xffffffffu, r, 1);      // lanes 0,2,4,6 hold pair sums
        double t2 = t + __shfl_down_sync(0xffffffffu, t, 2);     // lanes 0,4 hold quad sums
        double res = t2 + __shfl_down_sync(0xffffffffu, t2, 4);  // lane 0 holds the block sum
        if (act && u == 0) {
            if (n < 8) res = 0.0;
            for (int i = (n < 8 ? 0 : nb); i < n; i++) {          // tail (or the whole leaf when n < 8), sequential
                const int o = off + i - lo;
                if ((vmask[k][o >> 5] >> (o & 31)) & 1u) res += reward[o];
            }
            leaf[k][l] = res;
        }
    }
    __syncthreads();
    // ---- E: NumPy's pairwise combine, evaluated as a tree: every internal node adds its two children -- the operand pairs
    //      of the recursion's post-order program, so the same roundings -- and the nodes of one level are independent
    //      (the serial stack program on one thread per candidate kept the other 120 threads at the barrier for ~2500 cycles)
    {
        const int nl = prog->n_leaves;
#pragma unroll 1
        for (int q = tid; q < ny * nl; q += T) {                  // leaves outside the materialised range: exactly +0.0
            const int k = d2d_div_small(q, nl), l = q - k * nl;
            if (l < l0 || l > l1) leaf[k][l] = 0.0;
        }
        __syncthreads();
        const int nlev = prog->n_levels;
#pragma unroll 1
        for (int lev = 0; lev < nlev; lev++) {
            const int a = prog->lev_off[lev], cnt = prog->lev_off[lev + 1] - a;
#pragma unroll 1
            for (int q = tid; q < ny * cnt; q += T) {
                const int k = d2d_div_small(q, cnt), j = a + (q - k * cnt);
                leaf[k][prog->node_id[j]] = leaf[k][prog->node_l[j]] + leaf[k][prog->node_r[j]];
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        double max_reward = 0.0;
        int best = 0;
        const int root = prog->root;
        for (int k = 0; k < ny; k++) {
            const double sc = 0.0 + leaf[k][root];
            if (max_reward < sc) { best = k; max_reward = sc; }               // strict <, first maximum (:123-125)
        }
        actions_out[e] = P.tab->v_yaw_space[best] / P.max_yaw_speed;
        P.ox_calls[e] = ncall;
    }
}

// mode 0: every env (d2d_plan_oxford).  mode 1: every env whose step is complete when d2d_step_prim_warp_kernel has run,
// i.e. all but the ones waiting for an A* search (need_plan == 1) -- d2d_step_plan_oxford runs this beside the searches.
__global__ void __launch_bounds__(D2D_OX_THREADS, D2D_OX_MINB) d2d_oxford_kernel(const DevP P, const OxProgram *__restrict__ prog,
                                                                    double *__restrict__ actions_out, const int mode) {
    const int e = blockIdx.x;
    if (e >= P.B || (mode == 1 && P.need_plan[e] == 1)) return;
    d2d_oxford_env(P, prog, actions_out, e);
}
// ... and the envs of the step's planning list once their searches and d2d_step_post_list_kernel are done
__global__ void __launch_bounds__(D2D_OX_THREADS, D2D_OX_MINB) d2d_oxford_list_kernel(const DevP P, const OxProgram *__restrict__ prog,
                                                                         double *__restrict__ actions_out) {
    const int count = min(P.plan_list[P.B + 1 + P.plan_list[P.B + 4]], P.B);
#pragma unroll 1
    for (int li = blockIdx.x; li < count; li += gridDim.x) {
        d2d_oxford_env(P, prog, actions_out, P.plan_list[li]);
        __syncthreads();          // the block's shared arrays are reused by its next env
    }
}


// ------------------------------------------------------------------------------------------ scalar gaze policies
// NoControl / Rotating / LookAhead / LookGoal (yaw_planner.py:10-39, 136-142, 225-255), one warp per env.
// atan2 is d2d_atan2_cr (correctly rounded; glibc's is < 1 ulp and differs from it on ~0.09 % of inputs by one ulp): actions
// are held to 1e-12, and are bit-exact wherever glibc rounds correctly or the action saturates at +-1 (the common case).
__device__ __forceinline__ double d2d_turn_towards(const DevP &P, double target_yaw, double yaw) {
    const double m = P.max_yaw_speed;
    double v = (target_yaw - yaw) / P.dt;
    v = v < m ? v : m;
    v = v > -m ? v : -m;
    if (!(fabs(target_yaw - yaw) < 180.0)) v = -v;
    return v / m;
}

__global__ void __launch_bounds__(128) d2d_gaze_kernel(const DevP P, int policy, double *__restrict__ actions_out) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= P.B) return;
    const double RAD2DEG = 180.0 / D2D_PI;      // CPython math.degrees
    // an env that reported done and will be re-initialised by its next step is seen as freshly reset
    const bool fresh = P.rec[e].pending_reset || (P.auto_reset && P.done[e]);
    double a = 0.0;
    if (policy == D2D_GAZE_ROTATING) a = 1.0;
    else if (policy == D2D_GAZE_LOOKAHEAD) {
        const double vx = fresh ? 0.0 : P.rec[e].vx, vy = fresh ? 0.0 : P.rec[e].vy;
        if (!(vy == 0.0 && vx == 0.0)) {
            const double ty = d2d_pymod(d2d_atan2_cr(-vy, vx) * RAD2DEG, 360.0);
            a = d2d_turn_towards(P, ty, P.rec[e].yaw);
        }
    } else if (policy == D2D_GAZE_LOOKGOAL) {
        const int cursor = P.rec[e].cursor;
        const int len = fresh ? 0 : P.rec[e].nseg * P.n_way - cursor;
        if (len > 0) {
            const uint8_t *bel = P.belief + (size_t)e * D2D_BELIEF_STRIDE;
            int first = len;                       // first waypoint lying in an UNEXPLORED belief cell
#pragma unroll 1
            for (int w = lane; w < len && w < first; w += 32) {
                double x, y;
                d2d_waypoint_pos(P, e, cursor + w, x, y);
                const bool inb = !(x >= P.map_w || x < 0 || y >= P.map_h || y < 0);
                if (inb && bel[d2d_cell(x, P.scale, P.inv_scale) * D2D_GRID + d2d_cell(y, P.scale, P.inv_scale)] == 0)
                    first = w;
            }
            for (int off = 16; off > 0; off >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
            double xl, yl;
            d2d_waypoint_pos(P, e, cursor + (first < len ? first : len - 1), xl, yl);
            const double ty = d2d_pymod(d2d_atan2_cr(-(yl - P.rec[e].py), xl - P.rec[e].px) * RAD2DEG, 360.0);
            a = d2d_turn_towards(P, ty, P.rec[e].yaw);
        }
    }
    if (lane == 0) actions_out[e] = a;
}


// ------------------------------------------------------------------------------------------ Owl  yaw_planner.py:144-222
// Owl.plan for every env, called the way experiment.py:33-34 calls it (the class object is the instance).  One warp per env:
//   queue    a chosen action is repeated owl_repeat (= int(0.8 // dt) - 1) more times before the policy plans again (:193-196)
//   update_U the 36 direction-uncertainty bins move by -dp.d_i / depth + (l_hit inside the view wedge | l_miss outside),
//            clamped to [0, 1] (:178-185); lanes own bins
//   bearings d_g (goal), d_v (velocity; NaN for a drone at rest, which makes every cost NaN and argmin return 0),
//            d_o (k-th ACTIVE tracker), all through d2d_atan2_cr
//   costs    lane i owns candidate yaw rate u_space[i]: f = [G*(1-U), |v/10|^2*G*(1-U), sum_k ratio_k*G, U(yaw), |rad|],
//            cost = f . lamb as a sequential FMA chain (ndarray.dot on the reference image), first-minimum / first-NaN argmin.
//            zip(d_o, trackers) pairs the k-th active tracker's bearing with trackers[k] (:212-213): reproduced.
// Discrete output (one of 20 actions): bit-exact vs the reference wherever glibc's atan2 is correctly rounded; pow(x, 2) is x*x.
#define D2D_OWL_WARPS 4
#define D2D_OWL_MAX_TRK 480

__device__ __forceinline__ double d2d_owl_between(double a1m, double a2m) {   // angle_between on `% 360`-reduced operands
    const double diff = fabs(a1m - a2m), o = 360.0 - diff;
    if (diff != diff) return diff;                                           // np.minimum propagates NaN
    return diff <= o ? diff : o;
}
__device__ __forceinline__ double d2d_owl_G(double theta, double theta_h, double hp, double hm) {   // :172-177
    const double t = d2d_pymod(theta, 360.0);
    if (d2d_owl_between(t, 0.0) <= theta_h * 0.5) return 0.0;
    return (d2d_owl_between(t, hp) * D2D_DEG2RAD) * (d2d_owl_between(t, hm) * D2D_DEG2RAD);
}
__device__ __forceinline__ double d2d_owl_U(const double *U, double theta) {  // :186-189
    const double t = d2d_pymod(theta, 360.0);
    double bv = d2d_owl_between(0.0, t);
    if (bv != bv) return U[0];
    int best = 0;
#pragma unroll 1
    for (int i = 1; i < D2D_OWL_BINS; i++) {
        const double v = d2d_owl_between(10.0 * (double)i, t);
        if (v < bv) { bv = v; best = i; }
    }
    return U[best];
}

__global__ void __launch_bounds__(D2D_OWL_WARPS * 32) d2d_owl_kernel(const DevP P, double *__restrict__ actions_out) {
    __shared__ double U_s[D2D_OWL_WARPS][D2D_OWL_BINS];
    __shared__ double do_s[D2D_OWL_WARPS][D2D_OWL_MAX_TRK];      // bearing of the k-th active tracker
    __shared__ double ratio_s[D2D_OWL_WARPS][D2D_OWL_MAX_TRK];   // |mu_vel| / |mu_pos - p| of tracker k
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int e = blockIdx.x * D2D_OWL_WARPS + wid;
    if (e >= P.B) return;
    const EnvRec &r = P.rec[e];
    const bool fresh = r.pending_reset || (P.auto_reset && P.done[e]);
    const double maxyaw = P.max_yaw_speed;
    const int q = fresh ? 0 : P.owl_q[e];
    if (q > 0) {                                                 // :193-196
        if (lane == 0) { P.owl_q[e] = q - 1; actions_out[e] = P.owl_u[e] / maxyaw; }
        return;
    }
    const double x = fresh ? r.p0x : r.px, y = fresh ? r.p0y : r.py, yaw = fresh ? r.p0yaw : r.yaw;
    const double vx = fresh ? 0.0 : r.vx, vy = fresh ? 0.0 : r.vy;
    const double tgx = fresh ? r.p0x : r.tgx, tgy = fresh ? r.p0y : r.tgy;   // Planner.__init__: target = drone position
    const double dt = 0.8, theta_h = P.view_range_deg;
    const double RAD2DEG = 180.0 / D2D_PI;
    double *U = U_s[wid];
    // ---- update_U
    {
        const double dp0 = vx * dt, dp1 = vy * dt;
        const double nyaw = d2d_pymod(-yaw, 360.0);
        const double depth = P.view_depth;
#pragma unroll 1
        for (int i = lane; i < D2D_OWL_BINS; i += 32) {
            const double u0 = fresh ? 0.0 : P.owl_U[(size_t)e * D2D_OWL_BINS + i];
            double L = -D2D_FMA(dp1, P.tab->owl_sin[i], dp0 * P.tab->owl_cos[i]) / depth;
            L += (d2d_owl_between(10.0 * (double)i, nyaw) < theta_h * 0.5) ? 0.4 : -0.05;
            double u = u0 + L;
            u = (1.0 < u) ? 1.0 : u;
            u = (0.0 > u) ? 0.0 : u;
            P.owl_U[(size_t)e * D2D_OWL_BINS + i] = u;
            U[i] = u;
        }
    }
    // ---- bearings
    const double d_g = d2d_atan2_cr(tgy - y, tgx - x) * RAD2DEG;
    const double nv = d2d_norm2(vx, vy);
    const double d_v = d2d_atan2_cr(vy / nv, vx / nv) * RAD2DEG;
    int n_o = 0;
    if (!fresh && P.trackers) {
#pragma unroll 1
        for (int base = 0; base < P.N; base += 32) {
            const int k = base + lane;
            const size_t g = (size_t)e * P.NP + k;
            const bool act = k < P.N && P.trk_active[g] != 0;
            const unsigned b = __ballot_sync(0xffffffffu, act);
            const int pos = n_o + __popc(b & ((1u << lane) - 1u));
            if (act && pos < D2D_OWL_MAX_TRK) do_s[wid][pos] = d2d_atan2_cr(P.trk_mu[g * 4 + 1] - y, P.trk_mu[g * 4] - x) * RAD2DEG;
            n_o += __popc(b);
        }
        if (n_o > D2D_OWL_MAX_TRK) n_o = D2D_OWL_MAX_TRK;
#pragma unroll 1
        for (int k = lane; k < n_o; k += 32) {
            const double *mu = P.trk_mu + ((size_t)e * P.NP + k) * 4;
            ratio_s[wid][k] = d2d_norm2(mu[2], mu[3]) / d2d_norm2(mu[0] - x, mu[1] - y);
        }
    }
    __syncwarp();
    // ---- candidate costs
    const double hp = d2d_pymod(theta_h * 0.5, 360.0), hm = d2d_pymod(-(theta_h * 0.5), 360.0);
    const double nv10 = d2d_norm2(vx / 10.0, vy / 10.0);
    const double f1a = nv10 * nv10;
    const double ug = 1.0 - d2d_owl_U(U, d_g), uv = 1.0 - d2d_owl_U(U, d_v);
    double cost = 0.0;
    const bool cand = lane < P.n_owl_u;
    if (cand) {
        const double us = P.tab->owl_u_space[lane];
        const double cy = -(yaw + us * dt);
        const double f0 = d2d_owl_G(cy - d_g, theta_h, hp, hm) * ug;
        const double f1 = f1a * d2d_owl_G(cy - d_v, theta_h, hp, hm) * uv;
        double f2 = 0.0;
#pragma unroll 1
        for (int k = 0; k < n_o; k++) f2 += ratio_s[wid][k] * d2d_owl_G(cy - do_s[wid][k], theta_h, hp, hm);
        const double f3 = d2d_owl_U(U, cy);
        const double f4 = fabs((us * dt) * D2D_DEG2RAD);
        cost = D2D_FMA(f0, 0.2, 0.0);
        cost = D2D_FMA(f1, 0.9, cost);
        cost = D2D_FMA(f2, 1.0, cost);
        cost = D2D_FMA(f3, 0.1, cost);
        cost = D2D_FMA(f4, 0.0, cost);
    }
    // ---- np.argmin: the first NaN if any, else the first minimum
    const unsigned nanm = __ballot_sync(0xffffffffu, cand && cost != cost);
    int best;
    if (nanm) best = __ffs(nanm) - 1;
    else {
        double bc = cand ? cost : 1.0e308 * 10.0;     // +inf for idle lanes
        int bi = lane;
        for (int off = 16; off > 0; off >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, bc, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        best = bi;
    }
    if (lane == 0) {
        const double u = P.tab->owl_u_space[best];
        P.owl_q[e] = P.owl_repeat;
        P.owl_u[e] = u;
        actions_out[e] = u / maxyaw;
        if (fresh) P.rec[e].owl_fresh = 1;
    }
}
