// d2d_plan_host.inl -- host-side launch code of the Primitive planner path and the Oxford policy (included by drone2d.cu)

template <int E>
static int launch_pre(d2d_handle *h, cudaStream_t st) {
    const int rc = ensure_smem_attr(h, (const void *)d2d_step_pre_kernel<E>, "pre");
    if (rc != D2D_OK) return rc;
    d2d_step_pre_kernel<E><<<(h->B + E - 1) / E, h->T, h->smem_pre, st>>>(h->P);
    h->launches++;
    return D2D_OK;
}

template <int WPB>
static int launch_prim_warp(d2d_handle *h, const double *actions, cudaStream_t st) {
    const size_t smem = (size_t)WPB * d2d_warp_slice_bytes(h->NP, h->HW, d2d_prim_warp_extra(h->NP));
    if (smem > 227 * 1024) { h->err = "warp-per-env Primitive kernel: shared memory per block exceeds 227 KB"; return D2D_ERR_INVALID; }
    int rca = ensure_smem_attr(h, (const void *)d2d_step_prim_warp_kernel<WPB>, "prim warp");
    if (rca == D2D_OK) rca = ensure_smem_attr(h, (const void *)d2d_plan_kernel, "plan");
    if (rca != D2D_OK) return rca;
    h->P.use_parity = 1;
    d2d_step_prim_warp_kernel<WPB><<<(h->B + WPB - 1) / WPB, WPB * 32, smem, st>>>(h->P, actions);
    const int pgrid = h->B < D2D_PLAN_SLOTS ? h->B : D2D_PLAN_SLOTS;
    if (h->plan_small < 0) {
        // the small-footprint A* kernel runs first wherever its packed keys are valid (D2D_PLAN_SMALL=0: A/B switch)
        const char *ev = getenv("D2D_PLAN_SMALL");
        h->plan_small = 0;
        if (d2d_plan_small_ok(h->cfg.n_u, h->cfg.drone_max_speed) && !(ev && ev[0] == '0')) {
            const size_t ssm = d2d_plan_small_smem_bytes(h->NP);
            int per_sm = 0, sms = 0;
            if (ssm <= 227 * 1024 && ensure_smem_attr(h, (const void *)d2d_plan_small_kernel, "plan small") == D2D_OK &&
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, d2d_plan_small_kernel, D2D_PS_THREADS, ssm) == cudaSuccess &&
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device) == cudaSuccess && per_sm > 0)
                h->plan_small = per_sm * sms;
        }
    }
    if (h->plan_small > 0) {
        const int sgrid = h->B < h->plan_small ? h->B : h->plan_small;
        d2d_plan_small_kernel<<<sgrid, D2D_PS_THREADS, d2d_plan_small_smem_bytes(h->NP), st>>>(h->P);
        d2d_plan_kernel<<<pgrid, D2D_PLAN_THREADS2, h->smem_plan, st>>>(h->P, 1);      // the abandoned searches, if any
        h->launches++;
    } else {
        d2d_plan_kernel<<<pgrid, D2D_PLAN_THREADS2, h->smem_plan, st>>>(h->P, 0);
    }
    const int lgrid = (h->B + 3) / 4 < 148 * 4 ? (h->B + 3) / 4 : 148 * 4;
    d2d_step_post_list_kernel<4><<<lgrid, 128, 4 * d2d_warp_slice_bytes(1, 1, 0), st>>>(h->P, actions);
    h->launches += 3;
    return D2D_OK;
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
    d2d_oxford_kernel<<<h->B, D2D_OX_THREADS, 0, (cudaStream_t)stream>>>(h->P, h->ox_prog, actions_out_dev);
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
