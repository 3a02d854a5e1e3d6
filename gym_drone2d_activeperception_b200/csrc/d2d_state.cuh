// d2d_state.cuh -- device-side view of one handle's arena (passed by value to every kernel).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/drone2d.h"

// state_machine values (utils.py:24-29)
#define SM_WAIT_FOR_GOAL 0
#define SM_GOAL_REACHED 1
#define SM_PLANNING 2
#define SM_EXECUTING 3

#define D2D_GT_ROW_BYTES 400   // 50 rows x uint64 per env (bit j of row i == ground truth cell (i,j) == OCCUPIED)
#define D2D_GRID 50            // cells per side (map_size // map_scale); the reference's maps are 50x50
#define D2D_CELLS 2500
#define D2D_LOCAL 33
#define D2D_LOCAL_CELLS 1089
#define D2D_OX_TAB 4096        // table length of the last_time_observed accumulation sequences
#define D2D_OX_SEEN_STRIDE 2560
#define D2D_RVO_THETAS 32       // len(np.arange(0, 2*3.14, 0.2))
#define D2D_RVO_MAX_OBS 16      // circular obstacles (pillars) per env under the RVO motion profile
#define D2D_OWL_BINS 36         // len(np.arange(0, 360, 10)): Owl.U_list (yaw_planner.py:171)
#define D2D_MIRCNT_OFF 2552    // int counter in the zero padding behind the 2500 belief bytes of a shared-memory belief copy

struct DevTables {   // lookup tables in global memory (read through the read-only path)
    double u_space[D2D_MAX_U];
    double t_samp[D2D_MAX_SAMP], t_samp2[D2D_MAX_SAMP];
    double t_way[D2D_MAX_WAY], t_way2[D2D_MAX_WAY], t_way_x2[D2D_MAX_WAY];
    double v_yaw_space[D2D_MAX_YAW];
    double rvo_cos[D2D_RVO_THETAS], rvo_sin[D2D_RVO_THETAS];   // glibc cos / sin of np.arange(0, 2*3.14, 0.2) (utils.py:365)
    double owl_u_space[D2D_MAX_OWL_U];                         // np.arange(-max_yaw_speed, max_yaw_speed, max_yaw_speed/10)
    double owl_cos[D2D_OWL_BINS], owl_sin[D2D_OWL_BINS];       // glibc cos / sin of radians(np.arange(0, 360, 10)) (yaw_planner.py:181-182)
};

// Everything persistent and scalar about ONE env, as one 128-byte line: a warp loads / stores it with a single coalesced
// request (16 lanes x 8 B) instead of ~24 scattered sectors.  d2d_get_buffer exposes the fields as strided [B] views
// ("drone_x", "steps", ...).  The step kernels copy it verbatim into the head of their shared-memory EnvS.
struct __align__(16) EnvRec {
    double px, py, yaw, vx, vy;      // drone pose / velocity (utils.py:714-731)
    double tgx, tgy;                 // planner target (traj_planner.py:22)
    double p0x, p0y, p0yaw;          // reset pose (drone_v2.py:88-117)
    int steps, sm, fail, tcur;       // step counter, state machine, fail_count, next target index (drone_v2.py:153-163)
    int bufc, bufts, tracked;        // tracker_buffer count / summed ts, newly tracked agents (drone_v2.py:187, 232-235)
    int nseg, cursor;                // trajectory: remaining waypoints = nseg*n_way - cursor
    int obs_ix, obs_iy;              // drone cell for which the local_map tensor content is currently valid
    uint8_t pending_reset;           // reset requested by the host (d2d_request_reset), applied at the start of the next step
    uint8_t ox_fresh;                // Oxford state already re-initialised for a pending reset
    uint8_t owl_fresh;               // same for the Owl state
    uint8_t pad_[1];
};
static_assert(sizeof(EnvRec) == 128, "EnvRec must be exactly one 128-byte line");

struct DevP {
    int B, N, NP, HW;            // envs, agents, padded agents, hit words per env
    int n_rays, planner, trackers, auto_reset, n_targets;
    int n_u, n_samp, n_way, n_yaw;
    int step_parity, use_parity; // use_parity: warp-per-env Primitive path (parity itself lives in plan_list[B+3])
    int m_far;                   // first ray sample index at which the view-depth test can fire
    double dt, scale, inv_scale, map_w, map_h, agent_radius, max_acc, drone_r, max_yaw_speed;
    double ray_a0, ray_da;       // -FOV/2 and FOV/n_rays (utils.py:594)
    double depth2, fov, max_steps, var_cam, max_speed, cull_reach, ox_cos_thresh;
    double view_depth, view_range_deg;   // params.drone_view_depth, params.drone_view_range as given (Owl: yaw_planner.py:168, 183)
    double targets[D2D_MAX_TARGETS][2];
    // agents [B][NP] (env-major)
    double2 *apos, *apref, *apos0, *apref0;
    // RVO motion profile (motion_rvo != 0): agent.velocity is its own array (current / reset snapshot / output of d2d_rvo_kernel)
    int motion_rvo;
    double2 *avel, *avel0, *avel_next;
    double *rvo_obs;             // [B][D2D_RVO_MAX_OBS][3] x, y, rad
    int *rvo_nobs;               // [B]
    double *arad, *trk_radius0;
    uint64_t *gt_rows;           // [B][50]
    uint8_t *belief;             // [B][D2D_BELIEF_STRIDE]
    // drone + env bookkeeping: one 128-B record per env, plus the per-step verdict arrays [B]
    EnvRec *rec;
    uint8_t *collision, *dead_lock, *freezing, *done;
    // observation
    uint8_t *local_map;          // [B][1][33][33]
    float *yaw_obs;              // [B][1]
    // optional zero-copy HOST mirror of the observation (d2d_bind_host_mirror): device-visible addresses of pinned host
    // buffers; every store to local_map / yaw_obs / done is repeated there, so d2d_step_host has nothing to copy back
    uint8_t *lm_mirror;          // [B][1][33][33] or null
    float *yaw_mirror;           // [B] or null
    uint8_t *done_mirror;        // [B] or null
    // pipelined host path (d2d_step_pipelined): the kernel of a step is launched BEFORE the caller has chosen that step's
    // actions and runs everything that does not depend on them; just before the yaw update every warp waits until its
    // action has replaced the sentinel in the device staging buffer (delivered by the host's copy engine)
    unsigned long long *gate;            // the device staging buffer viewed as 64-bit slots (null: actions are valid at launch)
    unsigned int *gate_fault;            // pinned host word, set when a warp gave up waiting (the actions never came)
    // resident form (d2d_rollout_warp_kernel<GATED>): one kernel serves a whole run of host-driven steps
    unsigned long long *gate_count;      // device: (env, step) pairs completed since the kernel started
    unsigned int *gate_done;             // pinned host word: sequence number of the last step whose stores are all visible
    // courier block of the resident kernel (when an SM is free for it): polls the stamp word in pinned HOST memory and pulls
    // the step's actions out of the caller's pinned buffer itself -- no copy-engine transfer, no driver call per step
    const unsigned long long *gate_src;          // device-visible address of the caller's pinned action buffer (null: copy engine)
    const unsigned long long *gate_stamp_host;   // device-visible address of the pinned stamp word the host stores per step
#if defined(D2D_WARP_PROF) || defined(D2D_PLAN_PROF)
    unsigned long long *prof;    // [B][12]: 10 globaltimer stamps, smid, warpid of the fused warp kernel (tools/warp_prof.py builds with -DD2D_WARP_PROF)
#endif
    float *reward;               // [B] zeros (drone_v2.py:257)
    int8_t *hit;                 // [B][NP]
    // trackers [B][NP]
    uint8_t *trk_active;
    double *trk_mu, *trk_sigma, *trk_radius;   // [B][NP][4], [B][NP][16], [B][NP]
    int *trk_ts;
    // trajectory as A* segments
    double *traj_coeff;          // [B][D2D_MAX_SEGMENTS][6]
    uint8_t *need_plan, *plan_ok, *replan;
    int *tmp_act_cnt, *tmp_act_ts;   // still-active tracker totals, pre kernel -> post kernel
    // legacy np.random stream per env (MT19937 + cached gaussian), consumed by noisy measurements (var_cam != 0)
    uint32_t *rng_key, *rng_key0;   // [B][624] current / reset snapshot
    int *rng_pos, *rng_pos0, *rng_has, *rng_has0;
    double *rng_gauss, *rng_gauss0;
    // Oxford
    // Oxford policy state (yaw_planner.py:49): last_time_observed kept COMPACT -- per cell the index of the policy call at
    // which it was last visible (0 = never); the exact accumulated double is ox_tab[base][calls since] (the += dt sequence)
    uint16_t *ox_seen;           // [B][2560] (2500 used)
    int *ox_calls;               // [B] policy calls since reset
    const double *ox_tab;        // [2][D2D_OX_TAB] : from 0.0 (seen) and from 5.0 (never seen)
    double *ox_last;             // [B][2500] materialised on demand by d2d_oxford_export_kernel (may be null)
    // Owl policy state (yaw_planner.py:160-172), null unless cfg.oxford & D2D_POLICY_OWL
    double *owl_U;               // [B][36] direction-uncertainty bins U_list
    int *owl_q;                  // [B] entries left in the repeated-action queue `self.u`
    double *owl_u;               // [B] the queued action (deg/s)
    int n_owl_u, owl_repeat;
    // Jerk_Primitive planner (traj_planner.py:403-516)
    const d2d_jerk_tables *jerk; // per-heading tables in device memory (null unless planner == D2D_PLANNER_JERK)
    double2 *drone_acc;          // [B] Drone2D.acceleration (utils.py:724, 735): only this planner reads it
    unsigned long long *stats;   // [D2D_NUM_STATS]
    unsigned char *plan_ws;      // A* workspaces (Primitive planner)
    int *plan_list;              // [B+8]: compacted list of envs that need a plan; [B] count (block path); [B+1], [B+2]
                                 // double-buffered counts, [B+3] step counter, [B+4] this step's parity (warp path)
    int *plan_over;              // [B+8]: searches abandoned by d2d_plan_small_kernel (node capacity); [B] count, [B+1] ticket
    const DevTables *tab;
};
