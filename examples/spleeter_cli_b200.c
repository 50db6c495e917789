/* spleeter_cli_b200.c — the reference CLI's job (Executable/main.c:676-970) on the tier-B entry points, in C.
 *
 * Same command line and the same output files as the reference program:
 *     spleeter_cli_b200 spawnNthreads timeStep analyseBinLimit stems audioFile [model_fp16.bin]
 * writes <audioFile basename>_Vocal.wav and _Accompaniment.wav (stems <= 2) or _Drum.wav, _Vocal.wav and
 * _Accompaniment.wav (stems >= 3) into the current directory, as 32-bit float stereo at 44.1 kHz.
 *
 * Where main.c runs  decode -> channel_splitFloat -> stft -> processMT [-> residual -> processMT] -> istft (x n)
 * -> time-domain subtraction -> channel_joinFloat -> WAV writer  with every arrow a host loop over pageable
 * buffers, this host does  decode into pinned memory -> ONE call (srt_separate_batch_interleaved on a context
 * from srt_create_cli) -> WAV writer from pinned memory.  Split, join, cascade and subtraction happen on the
 * device.  spawnNthreads is accepted and ignored (tiles are batched on the GPU instead of threaded).
 *
 * Files that are not at 44.1 kHz are converted like main.c:264-270 does (srt_resample_host = the reference's libsamplerate
 * sinc converter) when the host's coefficient table is supplied in SRT_RESAMPLER_TABLE (raw float32 file, 22438 entries:
 * what decompressResamplerMQ produces, main.c:693-694).  Out of scope on purpose: FLAC/MP3 decoding; the input must be a
 * WAV (16-bit PCM or 32-bit float, 1 or 2 channels).
 *
 *   gcc -O2 -I include examples/spleeter_cli_b200.c -L spleeterrt_b200 -lspleeterrt_b200 \
 *       -Wl,-rpath,$PWD/spleeterrt_b200 -lm -o spleeter_cli_b200
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "srt_b200.h"

static uint32_t rd32(const unsigned char* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }
static uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | p[1] << 8); }

/* Decodes a RIFF/WAVE file into interleaved float frames in pinned memory (srt_host_alloc). */
static float* load_wav(const char* path, unsigned* channels, unsigned* rate, size_t* frames)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return NULL; }
    unsigned char hdr[12];
    if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) { fprintf(stderr, "%s: not a WAV file\n", path); fclose(f); return NULL; }
    unsigned fmt = 0, bits = 0;
    *channels = 0;
    for (;;) {
        unsigned char ck[8];
        if (fread(ck, 1, 8, f) != 8) break;
        const uint32_t sz = rd32(ck + 4);
        if (!memcmp(ck, "fmt ", 4)) {
            unsigned char b[40] = {0};
            const size_t take = sz < sizeof b ? sz : sizeof b;
            if (fread(b, 1, take, f) != take) break;
            fmt = rd16(b); *channels = rd16(b + 2); *rate = rd32(b + 4); bits = rd16(b + 14);
            if (fmt == 0xfffe && sz >= 26) fmt = rd16(b + 24);       /* WAVE_FORMAT_EXTENSIBLE sub-format */
            fseek(f, (long)(sz - take + (sz & 1)), SEEK_CUR);
        } else if (!memcmp(ck, "data", 4)) {
            const int is_f32 = fmt == 3 && bits == 32, is_i16 = fmt == 1 && bits == 16;
            if (!*channels || (!is_f32 && !is_i16)) { fprintf(stderr, "%s: only 16-bit PCM and 32-bit float WAV are decoded here\n", path); break; }
            const size_t n = sz / (bits / 8);
            float* x = (float*)srt_host_alloc((n ? n : 1) * sizeof(float));
            if (!x) break;
            if (is_f32) {
                if (fread(x, 4, n, f) != n) { fprintf(stderr, "%s: truncated\n", path); srt_host_free(x); break; }
            } else {
                int16_t* t = (int16_t*)malloc(n * 2);
                if (fread(t, 2, n, f) != n) { fprintf(stderr, "%s: truncated\n", path); free(t); srt_host_free(x); break; }
                for (size_t i = 0; i < n; i++) x[i] = (float)t[i] * (1.0f / 32768.0f);
                free(t);
            }
            fclose(f);
            *frames = n / *channels;
            return x;
        } else {
            fseek(f, (long)(sz + (sz & 1)), SEEK_CUR);
        }
    }
    fclose(f);
    return NULL;
}

/* 32-bit float stereo WAV, the format main.c:815-824 asks dr_wav for. */
static int save_wav_f32_stereo(const char* path, const float* frames, size_t n)
{
    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); return -1; }
    const uint32_t bytes = (uint32_t)(n * 8), rate = 44100;
    unsigned char h[58];
    memcpy(h, "RIFF", 4);
    const uint32_t riff = 4 + (8 + 16) + (8 + bytes);
    memcpy(h + 4, &riff, 4); memcpy(h + 8, "WAVEfmt ", 8);
    const uint32_t fsz = 16, brate = rate * 8; const uint16_t tag = 3, ch = 2, align = 8, bits = 32;
    memcpy(h + 16, &fsz, 4); memcpy(h + 20, &tag, 2); memcpy(h + 22, &ch, 2); memcpy(h + 24, &rate, 4);
    memcpy(h + 28, &brate, 4); memcpy(h + 32, &align, 2); memcpy(h + 34, &bits, 2);
    memcpy(h + 36, "data", 4); memcpy(h + 40, &bytes, 4);
    const int ok = fwrite(h, 1, 44, f) == 44 && fwrite(frames, 8, n, f) == n;
    return fclose(f) == 0 && ok ? 0 : -1;
}

static const char* base_name(const char* p)
{
    const char* s = strrchr(p, '/');
    return s ? s + 1 : p;
}

int main(int argc, char** argv)
{
    if (argc < 6) {
        printf("Invalid program arguments.\nExample:\n%s spawnNthreads timeStep analyseBinLimit stems audioFile [model_fp16.bin]\n", argv[0]);
        return -2;
    }
    /* argument clamps of main.c:720-748 */
    int T = atoi(argv[2]), F = atoi(argv[3]);
    const int n_out = atoi(argv[4]) <= 2 ? 2 : 3;
    if (T < 64) T = 64;
    if (F < 512) F = 512;
    if (F > 2048) F = 2048;
    const char* model = argc > 6 ? argv[6] : getenv("SRT_MODEL");
    if (!model) { fprintf(stderr, "model file missing: pass model_fp16.bin as the last argument or in SRT_MODEL\n"); return 1; }

    unsigned channels = 0, rate = 0;
    size_t n = 0;
    float* pcm = load_wav(argv[5], &channels, &rate, &n);
    if (!pcm) return -1;
    if (channels < 1 || channels > 2 || n == 0 || rate == 0) {
        fprintf(stderr, "%s: need 1 or 2 channels (got %u ch, %u Hz)\n", argv[5], channels, rate);
        return -1;
    }
    float* table = NULL;
    if (rate != 44100) {      /* main.c:262-270: ratio = 44100 / fs, ceil(n * ratio) frames */
        const char* tpath = getenv("SRT_RESAMPLER_TABLE");
        FILE* tf = tpath ? fopen(tpath, "rb") : NULL;
        table = (float*)malloc(22438 * sizeof(float));
        if (!tf || fread(table, 4, 22438, tf) != 22438) {
            fprintf(stderr, "%s is at %u Hz: set SRT_RESAMPLER_TABLE to the sinc table file (22438 float32) to convert it\n", argv[5], rate);
            return -1;
        }
        fclose(tf);
    }
    /* net 0 of the blob = drum net, net 1 = vocal net (main.c:759-760) */
    float* w = (float*)malloc(2 * (size_t)SRT_COEFF_FLOATS * sizeof(float));
    if (srt_load_model_fp16(model, 0, w) || srt_load_model_fp16(model, 1, w + SRT_COEFF_FLOATS)) { fprintf(stderr, "%s\n", srt_last_error()); return 1; }
    const float* coeffs[2] = {w, w + SRT_COEFF_FLOATS};

    /* tiles for the length at 44.1 kHz */
    const double ratio = 44100.0 / (double)rate;
    const size_t n_src = n;
    if (table) n = srt_resample_frames(n_src, ratio);
    const size_t padded = (size_t)SRT_FFTSIZE * ((n + SRT_FFTSIZE - 1) / SRT_FFTSIZE) + 2 * SRT_FFTSIZE;   /* main.c:762-763 */
    const int tiles = (int)((padded / SRT_HOPSIZE + (size_t)T - 1) / (size_t)T);
    srt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.time_step = T; cfg.bin_limit = F;
    cfg.max_batch_images = tiles;
    cfg.max_images = tiles < 16 ? tiles : 16;           /* U-Net tiles per pass; a longer file runs in several passes */
    srt_ctx* ctx = NULL;
    if (srt_create_cli(&cfg, n_out, n_out == 3 ? coeffs : coeffs + 1, &ctx)) { fprintf(stderr, "srt_create_cli: %s\n", srt_last_error()); return 1; }

    if (table) {
        float* conv = (float*)srt_host_alloc(n * channels * sizeof(float));
        size_t gen = 0;
        if (srt_resample_host(ctx, pcm, n_src, (int)channels, ratio, table, 22438, 491, conv, n, &gen)) { fprintf(stderr, "resample: %s\n", srt_last_error()); return 1; }
        srt_host_free(pcm);
        pcm = conv;
    }
    float* out[3];
    for (int q = 0; q < n_out; q++) out[q] = (float*)srt_host_alloc(n * 2 * sizeof(float));
    const float* in[1] = {pcm};
    const int ch[1] = {(int)channels};
    const size_t ns[1] = {n};
    if (srt_separate_batch_interleaved(ctx, in, ch, ns, 1, NULL, out)) { fprintf(stderr, "separate: %s\n", srt_last_error()); return 1; }

    static const char* names2[2] = {"Vocal", "Accompaniment"};
    static const char* names3[3] = {"Drum", "Vocal", "Accompaniment"};
    for (int q = 0; q < n_out; q++) {
        char path[4096];
        snprintf(path, sizeof path, "%s_%s.wav", base_name(argv[5]), n_out == 2 ? names2[q] : names3[q]);
        if (save_wav_f32_stereo(path, out[q], n)) return 1;
        printf("Saving file -> %s\n", path);
    }
    printf("%lld kernel launches, %d tile(s) of %d x %d\n", srt_launch_count(ctx), tiles, T, F);
    srt_destroy(ctx);
    for (int q = 0; q < n_out; q++) srt_host_free(out[q]);
    srt_host_free(pcm);
    free(w);
    return 0;
}
