// d2d_plan_host.inl -- host-side launch code of the Primitive planner path and the Oxford policy (included by drone2d.cu)

template <int E>
static int launch_pre(d2d_handle *h, cudaStream_t st) {
    static bool attr_done[64] = {false};
    const int dev = h->cfg.device;
    if (!attr_done[dev & 63]) {
        cudaError_t ce = cudaFuncSetAttribute(d2d_step_pre_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (ce != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute(pre): ") + cudaGetErrorString(ce); return D2D_ERR_CUDA; }
        attr_done[dev & 63] = true;
    }
    d2d_step_pre_kernel<E><<<(h->B + E - 1) / E, h->T, h->smem_pre, st>>>(h->P);
    h->launches++;
    return D2D_OK;
}

static int step_primitive(d2d_handle *h, const double *actions, cudaStream_t st) {
    int rc;
    switch (h->E) {
        case 4: rc = launch_pre<4>(h, st); break;
        case 16: rc = launch_pre<16>(h, st); break;
        default: rc = launch_pre<8>(h, st); break;
    }
    if (rc != D2D_OK) return rc;
    {
        static bool attr_done[64] = {false};
        const int dev = h->cfg.device;
        if (!attr_done[dev & 63]) {
            cudaError_t ce = cudaFuncSetAttribute(d2d_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (ce != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute(plan): ") + cudaGetErrorString(ce); return D2D_ERR_CUDA; }
            attr_done[dev & 63] = true;
        }
        const int grid = h->B < D2D_PLAN_SLOTS ? h->B : D2D_PLAN_SLOTS;
        d2d_plan_kernel<<<grid, D2D_PLAN_THREADS2, h->smem_plan, st>>>(h->P);
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
    if (!h->cfg.oxford) { h->err = "d2d_plan_oxford: handle was created with oxford = 0"; return D2D_ERR_STATE; }
    if (!h->world_set) { h->err = "d2d_plan_oxford before d2d_set_world"; return D2D_ERR_STATE; }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    {
        static bool attr_done[64] = {false};
        const int dev = h->cfg.device;
        if (!attr_done[dev & 63]) {
            CUDA_TRY(h, cudaFuncSetAttribute(d2d_oxford_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_done[dev & 63] = true;
        }
    }
    d2d_oxford_kernel<<<h->B, D2D_OX_THREADS, d2d_oxford_smem_bytes(h->cfg.n_yaw), (cudaStream_t)stream>>>(h->P, h->ox_prog, actions_out_dev);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return D2D_OK;
}
