static int step_primitive(d2d_handle *h, const double *actions, cudaStream_t st) {
    (void)actions; (void)st;
    h->err = "Primitive planner path not built yet";
    return D2D_ERR_INVALID;
}
extern "C" int d2d_plan_oxford(d2d_handle *h, double *actions_out_dev, void *stream) {
    (void)actions_out_dev; (void)stream;
    if (!h) return D2D_ERR_INVALID;
    h->err = "Oxford kernel not built yet";
    return D2D_ERR_INVALID;
}
