/* main_b200.c — minimal C host on the tier-B API: separates N copies of a synthetic 10 s stereo
 * stream into 4 stems on one B200.  Replaces main.c:774-792 (stft -> processMT -> istft).
 *
 *   gcc -O2 -I include examples/main_b200.c -L spleeterrt_b200 -lspleeterrt_b200 \
 *       -Wl,-rpath,$PWD/spleeterrt_b200 -lm -o main_b200
 *   ./main_b200 spleeterrt_b200/weights/model_fp16.bin 8
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "srt_b200.h"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s model_fp16.bin [n_streams]\n", argv[0]); return 2; }
    const int n_streams = argc > 2 ? atoi(argv[2]) : 4, n_stems = 4;
    const size_t n = 441000;
    /* the reference's 2-net fp16 blob -> fp32 (f32Decompress); stems 0,1 use net 0, stems 2,3 net 1 */
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    uint16_t* h = (uint16_t*)malloc(2 * (size_t)SRT_COEFF_FLOATS * 2);
    if (fread(h, 2, 2 * (size_t)SRT_COEFF_FLOATS, f) != 2 * (size_t)SRT_COEFF_FLOATS) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    float* w = (float*)malloc(2 * (size_t)SRT_COEFF_FLOATS * sizeof(float));
    srt_half_to_float(h, w, 2 * (size_t)SRT_COEFF_FLOATS);
    const float* coeffs[4] = {w, w, w + SRT_COEFF_FLOATS, w + SRT_COEFF_FLOATS};
    const int modes[4] = {1, 1, 1, 1};
    srt_config cfg = {0};
    cfg.n_stems = n_stems; cfg.time_step = 512; cfg.bin_limit = 1024;
    cfg.max_images = n_streams; cfg.max_batch_images = n_streams;
    srt_ctx* ctx = NULL;
    if (srt_create(&cfg, coeffs, modes, &ctx)) { fprintf(stderr, "srt_create: %s\n", srt_last_error()); return 1; }
    float *L = (float*)srt_host_alloc(n * 4), *R = (float*)srt_host_alloc(n * 4);
    for (size_t i = 0; i < n; i++) { L[i] = 0.3f * sinf(0.03f * i); R[i] = 0.3f * sinf(0.05f * i); }
    const float** pl = malloc(sizeof(float*) * n_streams); const float** pr = malloc(sizeof(float*) * n_streams);
    size_t* ns = malloc(sizeof(size_t) * n_streams);
    float** out = malloc(sizeof(float*) * n_streams * n_stems * 2);
    for (int i = 0; i < n_streams; i++) { pl[i] = L; pr[i] = R; ns[i] = n; }
    for (int i = 0; i < n_streams * n_stems * 2; i++) out[i] = (float*)srt_host_alloc(n * 4);
    if (srt_separate_batch(ctx, pl, pr, ns, n_streams, NULL, out)) { fprintf(stderr, "separate: %s\n", srt_last_error()); return 1; }
    double e = 0; for (size_t i = 0; i < n; i++) e += out[0][i] * out[0][i];
    printf("separated %d streams x %d stems; stem0 L rms %.5f; %lld kernel launches\n", n_streams, n_stems, sqrt(e / n), srt_launch_count(ctx));
    srt_destroy(ctx);
    return 0;
}
