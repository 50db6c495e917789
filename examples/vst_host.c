/*
 * vst_host.c — a plain-C stand-in for the JUCE plugin shell of SpleeterRT (VST/Source/PluginProcessor.cpp), driving the
 * streaming C API exactly the way the plugin does:
 *
 *   constructor   (PluginProcessor.cpp:46-87)    4 x malloc(getCoeffSize()), one fread of 39 290 900 bytes per
 *                                                <stem>4stems.dat file
 *   prepareToPlay (PluginProcessor.cpp:114-125)  msr = malloc(sizeof(Spleeter4Stems)); Spleeter4StemsInit(msr, 1536, 256, coeff)
 *   processBlock  (PluginProcessor.cpp:130-182)  the host's block of n frames goes through in slices of at most OVPSIZE
 *                                                (1024) samples: Spleeter4StemsProcessSamples(msr, inL + off, inR + off, m, ptr)
 *   destructor    (PluginProcessor.cpp:88-97)    Spleeter4StemsFree(msr)
 *
 * The same source builds against either header: include/Spleeter4Stems.h + libspleeterrt_b200.so (the GPU drop-in) or the
 * reference's VST/Source/Spleeter4Stems.h + the reference objects (oracle/_ref/vst_host_ref, built by oracle/build_ref.py), so
 * the two can be compared sample by sample (tests/test_vst_host.py) and timed block by block (BASELINE.json config 5).
 *
 *   vst_host drum.dat bass.dat accompaniment.dat vocal.dat in.f32 out.f32 [block=1024] [binLimit=1536] [timeStep=256]
 *
 * in.f32: interleaved stereo float32 frames; out.f32: 8 interleaved float32 channels per frame (components 0..7; the
 * plugin's `outputs[]`).  Untouched outputs stay 0 like a host's cleared buffer.  Prints one JSON line with the per-block
 * wall-clock latency (p50 / p99 / max, microseconds).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "Spleeter4Stems.h"

#ifndef OVPSIZE
#error "Spleeter4Stems.h must define OVPSIZE (PluginProcessor.cpp:178)"
#endif

size_t getCoeffSize(void); /* spleeter.h (both flavours) */

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp_double(const void* a, const void* b)
{
    const double x = *(const double*)a, y = *(const double*)b;
    return x < y ? -1 : x > y;
}

int main(int argc, char** argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s drum.dat bass.dat accompaniment.dat vocal.dat in.f32 out.f32 [block] [binLimit] [timeStep]\n", argv[0]);
        return 2;
    }
    const int block = argc > 7 ? atoi(argv[7]) : 1024;
    const int bin_limit = argc > 8 ? atoi(argv[8]) : 1536, time_step = argc > 9 ? atoi(argv[9]) : 256;
    if (block < 1 || block > 1 << 16) { fprintf(stderr, "bad block size\n"); return 2; }
    /* ---- constructor: the four weight files ---- */
    void* coeff[4];
    for (int i = 0; i < 4; i++) {
        coeff[i] = malloc(getCoeffSize());
        FILE* fp = fopen(argv[1 + i], "rb");
        if (!fp || fread(coeff[i], 1, 39290900, fp) != 39290900) { fprintf(stderr, "cannot read %s\n", argv[1 + i]); return 1; }
        fclose(fp);
    }
    /* ---- input ---- */
    FILE* fi = fopen(argv[5], "rb");
    if (!fi) { fprintf(stderr, "cannot open %s\n", argv[5]); return 1; }
    fseek(fi, 0, SEEK_END);
    const long frames = ftell(fi) / (2 * (long)sizeof(float));
    fseek(fi, 0, SEEK_SET);
    float* inter = (float*)malloc((size_t)frames * 2 * sizeof(float));
    if (fread(inter, sizeof(float), (size_t)frames * 2, fi) != (size_t)frames * 2) { fprintf(stderr, "short read\n"); return 1; }
    fclose(fi);
    float* in[2] = {(float*)malloc((size_t)block * sizeof(float)), (float*)malloc((size_t)block * sizeof(float))};
    float* out[8];
    for (int c = 0; c < 8; c++) out[c] = (float*)malloc((size_t)block * sizeof(float));
    float* result = (float*)calloc((size_t)frames * 8, sizeof(float));
    /* ---- prepareToPlay ---- */
    Spleeter4Stems* msr = (Spleeter4Stems*)malloc(sizeof(Spleeter4Stems));
    Spleeter4StemsInit(msr, bin_limit, time_step, coeff);
    /* ---- processBlock, block after block ---- */
    const long n_blocks = (frames + block - 1) / block;
    double* lat = (double*)malloc((size_t)n_blocks * sizeof(double));
    for (long b = 0; b < n_blocks; b++) {
        const long f0 = b * block;
        const int n = (int)(frames - f0 < block ? frames - f0 : block);
        for (int i = 0; i < n; i++) { in[0][i] = inter[(f0 + i) * 2]; in[1][i] = inter[(f0 + i) * 2 + 1]; }
        for (int c = 0; c < 8; c++) memset(out[c], 0, (size_t)n * sizeof(float));
        const double t0 = now_us();
        int offset = 0;
        while (offset < n) {
            float* ptr[8];
            for (int c = 0; c < 8; c++) ptr[c] = out[c] + offset;
            const int processing = n - offset < OVPSIZE ? n - offset : OVPSIZE;
            Spleeter4StemsProcessSamples(msr, in[0] + offset, in[1] + offset, processing, ptr);
            offset += processing;
        }
        lat[b] = now_us() - t0;
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 8; c++) result[(f0 + i) * 8 + c] = out[c][i];
    }
    /* ---- destructor ---- */
    Spleeter4StemsFree(msr);
    free(msr);
    FILE* fo = fopen(argv[6], "wb");
    if (!fo || fwrite(result, sizeof(float), (size_t)frames * 8, fo) != (size_t)frames * 8) { fprintf(stderr, "cannot write %s\n", argv[6]); return 1; }
    fclose(fo);
    /* steady-state latency: skip the first blocks (context warm-up) */
    const long skip = n_blocks > 40 ? 20 : 0, m = n_blocks - skip;
    double worst = 0;
    for (long b = skip; b < n_blocks; b++) worst = lat[b] > worst ? lat[b] : worst;
    qsort(lat + skip, (size_t)m, sizeof(double), cmp_double);
    printf("{\"blocks\": %ld, \"block\": %d, \"bin_limit\": %d, \"time_step\": %d, \"p50_us\": %.1f, \"p99_us\": %.1f, \"max_us\": %.1f}\n",
           n_blocks, block, bin_limit, time_step, lat[skip + m / 2], lat[skip + (long)(m * 0.99)], worst);
    for (int i = 0; i < 4; i++) free(coeff[i]);
    return 0;
}
