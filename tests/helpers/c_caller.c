/* A plain C caller of libdrone2d.so (include/drone2d.h): the drop-in boundary exercised without Python, ctypes or torch.
 * tests/test_gpu_c_abi.py writes the d2d_config and the world of a seeded batch to a file; this program creates the
 * handle, uploads the world, steps it through d2d_step_host with ordinary malloc'ed buffers and dumps what it got.
 *   c_caller <in.bin> <out.bin>
 * in.bin : int32 magic, B, N, T | d2d_config | agent_pos | agent_pref | agent_radius | tracker_radius | gt_grid | pose | actions[T][B]
 * out.bin: per step local_map [B][1089] u8, yaw [B] f32, done [B] u8; then int64 stats[D2D_NUM_STATS] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "drone2d.h"

static void *rd(FILE *f, size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p || (n && fread(p, 1, n, f) != n)) { fprintf(stderr, "short read (%zu bytes)\n", n); exit(2); }
    return p;
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: c_caller in.bin out.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 0x44324432) { fprintf(stderr, "bad header\n"); return 2; }
    const size_t B = (size_t)hdr[1], N = (size_t)hdr[2], T = (size_t)hdr[3];
    d2d_config *cfg = rd(f, sizeof(d2d_config));
    if (cfg->struct_size != (int32_t)sizeof(d2d_config)) { fprintf(stderr, "config size mismatch\n"); return 2; }
    double *pos = rd(f, B * N * 16), *pref = rd(f, B * N * 16), *rad = rd(f, B * N * 8), *trad = rd(f, B * N * 8);
    uint8_t *gt = rd(f, B * 2500);
    double *pose = rd(f, B * 24), *actions = rd(f, T * B * 8);
    fclose(f);

    d2d_handle *h = NULL;
    if (d2d_create(cfg, &h) != D2D_OK) { fprintf(stderr, "d2d_create: %s\n", d2d_last_error(NULL)); return 1; }
    if (d2d_set_world(h, 0, (int32_t)B, pos, pref, rad, trad, gt, pose) != D2D_OK) { fprintf(stderr, "d2d_set_world: %s\n", d2d_last_error(h)); return 1; }
    uint8_t *lm = malloc(B * 1089), *done = malloc(B);
    float *yaw = malloc(B * 4);
    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    for (size_t t = 0; t < T; t++) {
        if (d2d_step_host(h, actions + t * B, lm, yaw, done, NULL) != D2D_OK) { fprintf(stderr, "d2d_step_host: %s\n", d2d_last_error(h)); return 1; }
        fwrite(lm, 1, B * 1089, o); fwrite(yaw, 4, B, o); fwrite(done, 1, B, o);
    }
    int64_t st[D2D_NUM_STATS];
    if (d2d_stats(h, st, 0, NULL) != D2D_OK) { fprintf(stderr, "d2d_stats: %s\n", d2d_last_error(h)); return 1; }
    fwrite(st, 8, D2D_NUM_STATS, o);
    fclose(o);
    /* error behaviour across the ABI: a bad call returns a status and leaves a message, nothing aborts */
    if (d2d_step_host(h, NULL, NULL, NULL, NULL, NULL) != D2D_OK) { fprintf(stderr, "unexpected: NULL actions must use the staging buffer\n"); return 1; }
    d2d_buffer_info bi;
    if (d2d_get_buffer(h, "no_such_buffer", &bi) != D2D_ERR_INVALID) { fprintf(stderr, "unknown buffer not rejected\n"); return 1; }
    printf("c_caller ok: version %d, %zu envs x %zu steps, env_steps=%lld launches=%lld\n", d2d_version(), B, T, (long long)st[0],
           (long long)d2d_launch_count(h));
    d2d_destroy(h);
    return 0;
}
