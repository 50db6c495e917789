// srt_tier_a.cu — the reference's own C API (Executable/spleeter.h:64-69, stftFix.h:32-35)
// implemented on top of the tier-B context, so Executable/main.c links unchanged.
// These entry points have no error channel in the reference; failures abort loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/spleeter.h"
#include "../../include/srt_b200.h"
#include "../../include/stftFix.h"

struct _spleeter {
    srt_ctx* ctx;
    float* host_mask;   // getMaskPtr hands this out (spleeter.c:306-309 returns catLayer)
    size_t P;
};

static void die(const char* where)
{
    fprintf(stderr, "[spleeterrt_b200] %s failed: %s\n", where, srt_last_error());
    abort();
}

extern "C" size_t getCoeffSize(void) { return sizeof(spleeterCoeff); }

extern "C" void* allocateSpleeterStr(void) { return calloc(1, sizeof(struct _spleeter)); }

extern "C" void initSpleeter(struct _spleeter* nn, size_t width, size_t height, int stemMode, void* coeff)
{
    // One symbol for both of the reference's ABIs: Executable/spleeter.h:66 declares the sizes as size_t, VST/Source/spleeter.h:4 as
    // int.  An int caller leaves the upper halves of the two 64-bit argument registers undefined (SysV x86-64), so only the low 32
    // bits are looked at; no valid size needs more.
    width &= 0xffffffffu;
    height &= 0xffffffffu;
    srt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    const char* dev = getenv("SRT_DEVICE");
    cfg.device = dev ? atoi(dev) : 0;
    cfg.n_stems = 1;
    cfg.time_step = (int)height;
    cfg.bin_limit = (int)width;
    cfg.max_images = 1;
    cfg.flavour = 0;   // Executable flavour: LUT sigmoid, ELU clamp
    cfg.share_weights = 1;   // processMT hands ONE coefficient pointer to all its instances (main.c:557): pack / upload once
    const char* impl = getenv("SRT_CONV_IMPL");
    cfg.conv_impl = (impl && !strcmp(impl, "simt")) ? 1 : 0;
    const float* cp = (const float*)coeff;
    if (width % 64 || height % 64 || width < 64 || width > 2048 || height < 64) {
        // The reference CLI only warns about sizes that are not powers of two (main.c:739-742) and carries on; its six
        // halvings then truncate and the decoder's concatenations no longer line up.  This library needs multiples of 64
        // (include/spleeter.h) and says so instead of aborting or computing garbage.
        fprintf(stderr, "[spleeterrt_b200] initSpleeter: timeStep (%zu) and analyseBinLimit (%zu) must be multiples of 64 with "
                        "64 <= analyseBinLimit <= 2048; pick e.g. 512 1024 (the reference's defaults)\n", height, width);
        exit(2);
    }
    if (srt_create(&cfg, &cp, &stemMode, &nn->ctx)) die("initSpleeter");
    nn->P = width * height;
    nn->host_mask = (float*)malloc(sizeof(float) * 2 * nn->P);
}

extern "C" void getMaskPtr(struct _spleeter* nn, float** mask) { *mask = nn->host_mask; }

extern "C" void processSpleeter(struct _spleeter* nn, float* x, float* y)
{
    if (srt_unet_host(nn->ctx, x, 1, y)) die("processSpleeter");
}

extern "C" void freeSpleeter(struct _spleeter* nn)
{
    srt_destroy(nn->ctx);
    free(nn->host_mask);
    nn->ctx = nullptr;
    nn->host_mask = nullptr;
}

// ---- transforms -------------------------------------------------------------------------
struct stft_impl {
    srt_ctx* ctx;
    size_t cap_rows;
};

static srt_ctx* xform_ctx(OfflineSTFT* st, size_t rows)
{
    stft_impl* im = (stft_impl*)st->impl;
    if (im->ctx && im->cap_rows >= rows) return im->ctx;
    if (im->ctx) srt_destroy(im->ctx);
    srt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    const char* dev = getenv("SRT_DEVICE");
    cfg.device = dev ? atoi(dev) : 0;
    cfg.n_stems = 0;
    cfg.time_step = 64;
    cfg.max_images = (int)((rows + 63) / 64) + 1;
    if (srt_create(&cfg, nullptr, nullptr, &im->ctx)) die("InitSTFT");
    im->cap_rows = (size_t)cfg.max_images * 64;
    return im->ctx;
}

extern "C" void InitSTFT(OfflineSTFT* st, size_t targetCore)
{
    st->impl = calloc(1, sizeof(stft_impl));
    st->targetCore = targetCore;
}

extern "C" void FreeSTFT(OfflineSTFT* st)
{
    stft_impl* im = (stft_impl*)st->impl;
    if (im) {
        if (im->ctx) srt_destroy(im->ctx);
        free(im);
    }
    st->impl = nullptr;
}

extern "C" size_t stft(OfflineSTFT* st, const float* dataL, const float* dataR, size_t data_size, float** resultLRe, float** resultLIm,
                       float** resultRRe, float** resultRIm)
{
    const size_t rows = srt_stft_rows(data_size);
    srt_ctx* c = xform_ctx(st, rows);
    // caller frees these with free() (main.c:786-789): plain calloc, as stftFix.c:368-371
    *resultLRe = (float*)calloc(rows * FFTSIZE, sizeof(float));
    *resultLIm = (float*)calloc(rows * FFTSIZE, sizeof(float));
    *resultRRe = (float*)calloc(rows * FFTSIZE, sizeof(float));
    *resultRIm = (float*)calloc(rows * FFTSIZE, sizeof(float));
    if (srt_stft_host(c, dataL, dataR, data_size, *resultLRe, *resultLIm, *resultRRe, *resultRIm)) die("stft");
    return rows;
}

extern "C" size_t istft(OfflineSTFT* st, float* dataLRe, float* dataLIm, float* dataRRe, float* dataRIm, size_t data_size, float** resultL,
                        float** resultR)
{
    srt_ctx* c = xform_ctx(st, data_size);
    const size_t n = data_size * HOPSIZE + (FFTSIZE - HOPSIZE);
    *resultL = (float*)calloc(n, sizeof(float));
    *resultR = (float*)calloc(n, sizeof(float));
    if (srt_istft_host(c, dataLRe, dataLIm, dataRRe, dataRIm, data_size, *resultL, *resultR)) die("istft");
    return n;
}

// main.c calls this on Linux (main.c:675,689); nothing to configure here.
extern "C" void openblas_set_num_threads(int) {}
