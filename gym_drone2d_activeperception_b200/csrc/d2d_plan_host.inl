// d2d_plan_host.inl -- host-side launch code of the Primitive planner path and the Oxford policy (included by drone2d.cu)

template <int E>
static int launch_pre(d2d_handle *h, cudaStream_t st) {
    const int rc = ensure_smem_attr(h, (const void *)d2d_step_pre_kernel<E>, "pre");
    if (rc != D2D_OK) return rc;
    d2d_step_pre_kernel<E><<<(h->B + E - 1) / E, h->T, h->smem_pre, st>>>(h->P);
    h->launches++;
    return D2D_OK;
}

// ox_next == nullptr: the step (d2d_step).  ox_next != nullptr (d2d_step_plan_oxford): the step, then Oxford.plan of every env into
// ox_next -- the searches of the planning envs and their completion run on the handle's side stream while the main stream
// already scores the envs whose step d2d_step_prim_warp_kernel has completed (95 % on BASELINE config 4); the A* kernel is
// a latency chain that leaves most of every SM idle, which is where the Oxford blocks run.
template <int WPB>
static int launch_prim_warp(d2d_handle *h, const double *actions, cudaStream_t st, double *ox_next = nullptr) {
    const size_t smem = (size_t)WPB * d2d_warp_slice_bytes(h->NP, h->HW, d2d_prim_warp_extra(h->NP));
    if (smem > 227 * 1024) { h->err = "warp-per-env Primitive kernel: shared memory per block exceeds 227 KB"; return D2D_ERR_INVALID; }
    int rca = ensure_smem_attr(h, (const void *)d2d_step_prim_warp_kernel<WPB>, "prim warp");
    if (rca == D2D_OK) rca = ensure_smem_attr(h, (const void *)d2d_plan_kernel, "plan");
    if (rca != D2D_OK) return rca;
    h->P.use_parity = 1;
    if (h->plan_small < 0) {
        // the small-footprint A* kernel runs first wherever its packed keys are valid (D2D_PLAN_SMALL=0: A/B switch)
        const char *ev = getenv("D2D_PLAN_SMALL");
        h->plan_small = 0;
        if (cudaDeviceGetAttribute(&h->plan_sms, cudaDevAttrMultiProcessorCount, h->cfg.device) != cudaSuccess) h->plan_sms = 148;
        if (d2d_plan_small_ok(h->cfg.n_u, h->cfg.drone_max_speed) && !(ev && ev[0] == '0')) {
            const size_t ssm = d2d_plan_small_smem_bytes(h->NP);
            int per_sm = 0;
            if (ssm <= 227 * 1024 && ensure_smem_attr(h, (const void *)d2d_plan_small_kernel, "plan small") == D2D_OK &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, d2d_plan_small_kernel, D2D_PS_THREADS, ssm) == cudaSuccess &&
                per_sm > 0)
                h->plan_small = per_sm * h->plan_sms;
        }
    }
    cudaStream_t ps = st;
    int small_grid = h->plan_small, large_grid = D2D_PLAN_SLOTS;
    if (ox_next) {
        if (!h->side_stream) {
            CUDA_TRY(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; i++) CUDA_TRY(h, cudaEventCreateWithFlags(&h->side_ev[i], cudaEventDisableTiming));
        }
        ps = h->side_stream;
        // D2D_PLAN_OVERLAP_SLOTS caps the searches per SM beside the Oxford blocks (measured on config 4: 2 / 3 / 5 per SM ->
        // 1.573 / 1.553 / 1.547 ms per step; no cap by default)
        static const int slots = getenv("D2D_PLAN_OVERLAP_SLOTS") ? atoi(getenv("D2D_PLAN_OVERLAP_SLOTS")) : 0;
        if (small_grid > 0 && slots > 0 && slots * h->plan_sms < small_grid) small_grid = slots * h->plan_sms;
        if (h->plan_small <= 0) large_grid = h->plan_sms;         // one 104 KB search per SM beside the Oxford blocks
    }
    d2d_step_prim_warp_kernel<WPB><<<(h->B + WPB - 1) / WPB, WPB * 32, smem, st>>>(h->P, actions);
    if (ox_next) {
        CUDA_TRY(h, cudaEventRecord(h->side_ev[0], st));
        CUDA_TRY(h, cudaStreamWaitEvent(ps, h->side_ev[0], 0));
    }
    const int pgrid = h->B < large_grid ? h->B : large_grid;
    if (h->plan_small > 0) {
        const int sgrid = h->B < small_grid ? h->B : small_grid;
        d2d_plan_small_kernel<<<sgrid, D2D_PS_THREADS, d2d_plan_small_smem_bytes(h->NP), ps>>>(h->P);
        d2d_plan_kernel<<<pgrid, D2D_PLAN_THREADS2, h->smem_plan, ps>>>(h->P, 1);      // the abandoned searches, if any
        h->launches++;
    } else {
        d2d_plan_kernel<<<pgrid, D2D_PLAN_THREADS2, h->smem_plan, ps>>>(h->P, 0);
    }
    const int lgrid = (h->B + 3) / 4 < 148 * 4 ? (h->B + 3) / 4 : 148 * 4;
    d2d_step_post_list_kernel<4><<<lgrid, 128, 4 * d2d_warp_slice_bytes(1, 1, 0), ps>>>(h->P, actions);
    h->launches += 3;
    if (ox_next) {
        // the planning envs are scored on the side stream too, as soon as they are complete: the main stream's Oxford kernel
        // outlasts the searches (0.8 ms against 0.25 ms on BASELINE config 4), so nothing of them is left on the critical path
        const int ogrid = h->B < h->plan_sms * D2D_OX_MINB ? h->B : h->plan_sms * D2D_OX_MINB;
        d2d_oxford_list_kernel<<<ogrid, D2D_OX_THREADS, 0, ps>>>(h->P, h->ox_prog, ox_next);
        CUDA_TRY(h, cudaEventRecord(h->side_ev[1], ps));
        d2d_oxford_kernel<<<h->B, D2D_OX_THREADS, 0, st>>>(h->P, h->ox_prog, ox_next, 1);       // beside the searches
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->side_ev[1], 0));
        h->launches += 2;
    }
    return D2D_OK;
}
static int launch_prim_warp_entry(d2d_handle *h, const double *actions, cudaStream_t st, double *ox_next) {
    return launch_prim_warp<4>(h, actions, st, ox_next);
}

static int step_primitive(d2d_handle *h, const double *actions, cudaStream_t st) {
    if (h->cfg.envs_per_block <= 0) return launch_prim_warp<4>(h, actions, st);   // default: one warp per env
    h->P.use_parity = 0;
    int rc;
    switch (h->E) {
        case 4: rc = launch_pre<4>(h, st); break;
        case 16: rc = launch_pre<16>(h, st); break;
        default: rc = launch_pre<8>(h, st); break;
    }
    if (rc != D2D_OK) return rc;
    {
        const int rca = ensure_smem_attr(h, (const void *)d2d_plan_kernel, "plan");
        if (rca != D2D_OK) return rca;
        const int grid = h->B < D2D_PLAN_SLOTS ? h->B : D2D_PLAN_SLOTS;
        d2d_plan_kernel<<<grid, D2D_PLAN_THREADS2, h->smem_plan, st>>>(h->P, 0);
        h->launches++;
    }
    {
        const int E = 8;
        d2d_step_post_kernel<E><<<(h->B + E - 1) / E, 256, d2d_step_smem_bytes(E, 1, 1), st>>>(h->P, actions);
        h->launches++;
    }
    return D2D_OK;
}

extern "C" int d2d_plan_oxford(d2d_handle *h, double *actions_out_dev, void *stream) {
    if (!h || !actions_out_dev) return D2D_ERR_INVALID;
    if (!(h->cfg.oxford & D2D_POLICY_OXFORD)) { h->err = "d2d_plan_oxford: handle was created without the Oxford state (cfg.oxford & 1)"; return D2D_ERR_STATE; }
    if (!h->world_set) { h->err = "d2d_plan_oxford before d2d_set_world"; return D2D_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    d2d_oxford_kernel<<<h->B, D2D_OX_THREADS, 0, (cudaStream_t)stream>>>(h->P, h->ox_prog, actions_out_dev, 0);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return D2D_OK;
}

extern "C" int d2d_plan_gaze(d2d_handle *h, int32_t policy, double *actions_out_dev, void *stream) {
    if (!h || !actions_out_dev) return D2D_ERR_INVALID;
    if (policy < D2D_GAZE_NOCONTROL || policy > D2D_GAZE_OWL) { h->err = "d2d_plan_gaze: unknown policy"; return D2D_ERR_INVALID; }
    if (!h->world_set) { h->err = "d2d_plan_gaze before d2d_set_world"; return D2D_ERR_STATE; }
    if (policy == D2D_GAZE_OWL && !(h->cfg.oxford & D2D_POLICY_OWL)) {
        h->err = "d2d_plan_gaze(Owl): handle was created without the Owl state (cfg.oxford & D2D_POLICY_OWL)"; return D2D_ERR_STATE;
    }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (policy == D2D_GAZE_OWL)
        d2d_owl_kernel<<<(h->B + D2D_OWL_WARPS - 1) / D2D_OWL_WARPS, D2D_OWL_WARPS * 32, 0, (cudaStream_t)stream>>>(h->P, actions_out_dev);
    else
        d2d_gaze_kernel<<<(h->B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(h->P, policy, actions_out_dev);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return D2D_OK;
}
