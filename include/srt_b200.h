/*
 * srt_b200.h — tier-B C ABI of libspleeterrt_b200.so: the batched, device-resident form of
 * SpleeterRT's spectrogram-to-mask hot path on one B200 (sm_100a).
 *
 *   PCM -> STFT framer -> magnitude -> U-Net soft masks (n_stems nets) -> mask * spectrum
 *       -> inverse STFT -> overlap-add -> stems
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature.  Every entry
 * point returns 0 on success or a negative srt_status; srt_last_error() gives the text.
 * There is no CPU fallback: without a working CUDA device every call fails loudly.
 *
 * What each entry point replaces in the reference (james34602/SpleeterRT):
 *   srt_create            initSpleeter()  Executable/spleeter.c:111-172 (x n_stems, weights
 *                         repacked for the tensor cores) + InitSTFT() stftFix.c:302-341 tables
 *   srt_unet_host         processSpleeter()  Executable/spleeter.c:177-301, batched
 *   srt_separate_batch    main.c:762-806 = channel_splitFloat framing + stft() + processMT()
 *                         (main.c:444-541) + istft() + channel_joinFloat, for many streams and
 *                         n_stems nets (the VST's "one net per stem", Spleeter4Stems.c:135)
 *   srt_stft_host / srt_istft_host   stft()/istft()  Executable/stftFix.c:363-579
 * The reference-compatible tier-A symbols (spleeter.h, stftFix.h) are thin shims over these.
 */
#ifndef SRT_B200_H
#define SRT_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SRT_FFTSIZE 4096
#define SRT_HOPSIZE 1024
#define SRT_BINS 2049
#define SRT_COEFF_FLOATS 9822725 /* sizeof(spleeterCoeff)/4, Executable/spleeter.h:5-31 */
#define SRT_MAX_STEMS 8

typedef enum {
    SRT_OK = 0,
    SRT_ERR_ARG = -1,      /* invalid argument (T/F not multiples of 64, too many stems, ...) */
    SRT_ERR_CUDA = -2,     /* CUDA runtime / driver error, or no sm_100 device */
    SRT_ERR_CAPACITY = -3, /* batch larger than the context was created for */
    SRT_ERR_STATE = -4
} srt_status;

typedef struct srt_ctx srt_ctx;

typedef struct {
    int device;            /* CUDA device ordinal */
    int n_stems;           /* nets evaluated per image, 0..SRT_MAX_STEMS (0 = transforms only) */
    int time_step;         /* T: spectrogram frames per U-Net image (height), multiple of 64 */
    int bin_limit;         /* F: frequency bins fed to the U-Net (width), multiple of 64, <= 2048 */
    int max_images;        /* U-Net batch capacity: T-frame tiles evaluated per pass */
    int max_batch_images;  /* tiles held per srt_separate_* call (>= max_images; 0 = max_images) */
    int flavour;           /* 0 = Executable (LUT sigmoid, ELU clamp -15), 1 = VST (exact sigmoid) */
    int conv_impl;         /* 0 = tcgen05 tensor-core kernels (the product), 1 = SIMT verification kernels */
    void* cuda_stream;     /* optional cudaStream_t to enqueue on (NULL = context-owned stream) */
    int precision;         /* arithmetic of the ten tensor-core layers (accumulation is fp32 in all):
                            *   0 = SRT_PRECISION_COMPENSATED (default): activations feed the MMAs as tf32(a) PLUS the residual
                            *       a - tf32(a) (a second, smaller contraction into the same accumulator).  The residual travels in
                            *       8 bits (e5m2, scaled by 4): operand error 2^-12 -> ~2^-16, about what the tensor core's own
                            *       accumulation adds; stems agree with the fp32 reference to a few 1e-6 RMS at any input level;
                            *   2 = SRT_PRECISION_COMPENSATED_BF16: all residuals in bf16 (operand error ~2^-19; ~6 % slower);
                            *   1 = SRT_PRECISION_TF32: single-pass TF32 operands (2^-12): ~1.3x faster U-Net, stem error
                            *       ~1e-4 of the stem's level (3e-5 RMS on the -12 dBFS test signal, 1e-4 at full scale). */
    int share_weights;     /* 1: contexts created on the same device with the same coeffs POINTERS, stem modes and configuration
                            * share one device copy of the packed weights and tables (created once, reference-counted) instead of
                            * packing and uploading per context.  The caller promises the blobs stay unchanged while any such
                            * context lives - exactly the contract of the reference's initSpleeter, which keeps the pointer
                            * (spleeter.c:129) and is handed ONE pointer for all instances of a run (main.c:557).  The tier-A
                            * initSpleeter sets it; 0 (default) copies as before. */
} srt_config;
#define SRT_PRECISION_COMPENSATED 0
#define SRT_PRECISION_TF32 1
#define SRT_PRECISION_COMPENSATED_BF16 2

/* coeffs[s]: one spleeterCoeff blob (SRT_COEFF_FLOATS floats, host memory) per stem;
 * stem_modes[s]: 0 = LeakyReLU(0.2)/ReLU, !=0 = ELU/ELU (spleeter.c:130-139).
 * The blobs are copied; they need not outlive the call. */
int srt_create(const srt_config* cfg, const float* const* coeffs, const int* stem_modes, srt_ctx** out);
/* The CLI's output modes as one device-resident call (Executable/main.c:776-970; the reference runs them as
 * stft -> processMT -> [residual -> processMT] -> istft x n -> time-domain subtraction on the host).
 *   n_outputs = 2: coeffs = {vocal net}.        Output pairs per stream: [vocal, accompaniment = input - vocal]
 *                  (main.c:782-794; vocal net = coeffProvPtr1, stemMode 0).
 *   n_outputs = 3: coeffs = {drum net, vocal net}.  Output pairs: [drum, vocal, accompaniment]: the drum net
 *                  (coeffProvPtr2, stemMode 1) masks the spectrum, the vocal net runs on the magnitudes of
 *                  residual = spectrum - drum spectrum, accompaniment = istft(residual) - vocal (main.c:845-927).
 * cfg->n_stems is ignored.  The returned context is used with srt_separate_batch / _batch_async / _device exactly
 * like one from srt_create: stems_out holds n_outputs (L, R) pairs per stream and unaffected[0] (NULL = 0.1) is
 * the single unaffectedWeight of main.c:773. */
int srt_create_cli(const srt_config* cfg, int n_outputs, const float* const* coeffs, srt_ctx** out);
/* (L, R) output pairs per stream of srt_separate_*: n_stems, or n_outputs for a context from srt_create_cli */
int srt_output_pairs(const srt_ctx* ctx);
void srt_destroy(srt_ctx* ctx);
const char* srt_last_error(void);

/* fp16 model blob (spleeterQuantized halves, Executable/spleeter.h:32-62) -> fp32 with
 * denormals flushed to zero (f32Decompress, main.c:423-434).  Host side, exact. */
void srt_half_to_float(const uint16_t* in, float* out, size_t n);

/* ---- weight files (host side, exact) ----------------------------------------------------------
 * srt_load_coeff_dat: one fp32 spleeterCoeff dump of exactly 39 290 900 bytes, the files the VST reads with one
 *   fread (VST/Source/PluginProcessor.cpp:48-80: drum4stems.dat ...).  Unlike the reference a missing or short
 *   file is an error (SRT_ERR_ARG), not uninitialised weights.  srt_save_coeff_dat writes the same format.
 * srt_model_fp16_nets / srt_load_model_fp16: the fp16 model blob (spleeterQuantized, Executable/spleeter.h:32-62:
 *   consecutive nets of 9 822 725 halves); net `net` is expanded like f32Decompress (main.c:423-443).  In the
 *   reference's model net 0 is the drum net (stemMode 1) and net 1 the vocal net (stemMode 0), main.c:759-760.
 * srt_pack_layer: one net's weights of a tensor-core layer (0..4 = down2..down6, 5..9 = up1..up5) in the K-major,
 *   128-byte-swizzled B-operand layout the tcgen05 kernels stream (form 0: generic kernel, form 1: row-patch kernel;
 *   weights that are not TF32-exact get the two-term split).  out == NULL returns the size in floats. */
int srt_load_coeff_dat(const char* path, float* coeff_out /* SRT_COEFF_FLOATS */);
int srt_save_coeff_dat(const char* path, const float* coeff);
int srt_model_fp16_nets(const char* path);
int srt_load_model_fp16(const char* path, int net, float* coeff_out /* SRT_COEFF_FLOATS */);
long long srt_pack_layer(int layer, int form, int time_step, int bin_limit, const float* coeff, float* out, size_t cap_floats);

/* ---- U-Net only -------------------------------------------------------------------------
 * x: n_img images, each planar [2][T][F] like processSpleeter's input (host memory).
 * y: [n_stems][n_img][2][T][F] masks (host memory).  n_img <= max_images. */
int srt_unet_host(srt_ctx* ctx, const float* x, int n_img, float* y);
/* device-resident variant: d_mag [n_img][T][F][2] (channel-interleaved), d_mask
 * [n_stems][n_img][T][F][2]; pointers are device addresses on ctx's device. */
int srt_unet_device(srt_ctx* ctx, const float* d_mag, int n_img, float* d_mask);

/* ---- full path --------------------------------------------------------------------------
 * n_streams stereo streams; stream i has n_samples[i] samples per channel (planar).
 * stems_out[i * n_stems * 2 + s * 2 + c] receives stem s, channel c of stream i
 * (n_samples[i] floats; for a srt_create_cli context read n_outputs for n_stems).  unaffected[s] scales the bins >= F (main.c:486-493; the VST uses
 * 0.25 / 0.0, Spleeter4Stems.c:73,281); NULL = 0.1 for every stem (main.c:773).
 * *_batch takes host pointers (copies inside), *_device takes device pointers. */
int srt_separate_batch(srt_ctx* ctx, const float* const* pcmL, const float* const* pcmR,
                       const size_t* n_samples, int n_streams, const float* unaffected,
                       float* const* stems_out);
/* Asynchronous flavour of srt_separate_batch for a server that keeps the GPU and both PCIe
 * directions busy: enqueues the upload, the kernels and the download of one batch and returns a
 * ticket without waiting.  Three batches may be in flight per context (triple-buffered staging):
 * batch k+1 uploads while batch k computes and batch k-1 is still being copied back.  A fourth
 * call blocks until the oldest batch has drained.  pcm and stems_out buffers must stay valid (and
 * should be pinned, srt_host_alloc) until srt_batch_wait(ticket) returns; results are complete
 * only then.  Copies of channels that sit back to back in host memory are merged into one DMA.
 * (The reference has no counterpart: main.c processes one file per process invocation.) */
int srt_separate_batch_async(srt_ctx* ctx, const float* const* pcmL, const float* const* pcmR,
                             const size_t* n_samples, int n_streams, const float* unaffected,
                             float* const* stems_out, int* ticket_out);
int srt_batch_wait(srt_ctx* ctx, int ticket);
int srt_separate_device(srt_ctx* ctx, const float* const* d_pcmL, const float* const* d_pcmR,
                        const size_t* n_samples, int n_streams, const float* unaffected,
                        float* const* d_stems_out);

/* ---- interleaved frames either side of the path ------------------------------------------------
 * The reference CLI decodes to interleaved frames and splits them on the host (channel_splitFloat, main.c:53-76,
 * 767), duplicates a mono channel (main.c:768-769), and joins each result back into interleaved frames for the
 * float32 WAV writer (channel_joinFloat, main.c:806, 815-824).  These entry points take and return those buffers
 * directly: the split is the stride of the STFT framer's loads, the join the stride of the overlap-add stores.
 *   pcm[i]       n_samples[i] frames of channels[i] (1 or 2) interleaved floats
 *   out[i * pairs + q]   n_samples[i] interleaved stereo frames (2 * n_samples[i] floats) of output pair q,
 *                pairs = srt_output_pairs(ctx)
 * Results are bit-identical to the planar entry points.  Device pointers must be 4-byte aligned. */
int srt_separate_batch_interleaved(srt_ctx* ctx, const float* const* pcm, const int* channels, const size_t* n_samples,
                                   int n_streams, const float* unaffected, float* const* out);
int srt_separate_batch_interleaved_async(srt_ctx* ctx, const float* const* pcm, const int* channels,
                                         const size_t* n_samples, int n_streams, const float* unaffected,
                                         float* const* out, int* ticket_out);
int srt_separate_device_interleaved(srt_ctx* ctx, const float* const* d_pcm, const int* channels,
                                    const size_t* n_samples, int n_streams, const float* unaffected,
                                    float* const* d_out);

/* ---- sample-rate conversion in front of the path ----------------------------------------------
 * JamesDSPOfflineResampling (main.c:209-224, called at main.c:264-270 when the file is not at 44.1 kHz) = libsamplerate
 * src_simple() on the sinc interpolator of Executable/libsamplerate/src_sinc.c.  The coefficient table is the host's
 * (`decompressedCoefficients`, main.c:277, 693-694): coeff_count = 22438 floats and index_inc = 491 in the reference
 * (src_sinc.c:141-143).  in: n_in frames of `channels` (1 or 2) interleaved floats; out: n_out frames, normally
 * srt_resample_frames(n_in, ratio) = ceil(n_in * ratio) (main.c:265); *n_generated = frames the converter produced
 * (it may stop one frame short, src_sinc.c:323-326), the rest of `out` is zero like the reference's buffer.
 * Bit-identical to the reference build.  ratio = 44100 / file rate. */
size_t srt_resample_frames(size_t n_in, double ratio);
int srt_resample_host(srt_ctx* ctx, const float* in, size_t n_in, int channels, double ratio, const float* coeffs,
                      int coeff_count, int index_inc, float* out, size_t n_out, size_t* n_generated);
int srt_resample_device(srt_ctx* ctx, const float* d_in, size_t n_in, int channels, double ratio, const float* d_coeffs,
                        int coeff_count, int index_inc, float* d_out, size_t n_out, size_t* n_generated);
/* host-only: the converter's bookkeeping (input frame and 12-bit fixed-point table offset per output frame);
 * returns the number of frames generated, frames / starts may be NULL */
long long srt_resample_plan(size_t n_in, int channels, double ratio, int coeff_count, int index_inc, size_t n_out,
                            int32_t* frames, int32_t* starts);

/* ---- transforms (Executable/stftFix.c) --------------------------------------------------
 * srt_stft_host: rows = ceil(n/1024); planes are [rows][4096] host buffers supplied by the
 * caller, zero-filled on return outside bins 0..2048 of the computed rows (stftFix.c:367-371).
 * srt_istft_host: planes [frames][4096] -> outL/outR of frames*1024+3072 samples. */
size_t srt_stft_rows(size_t n);
int srt_stft_host(srt_ctx* ctx, const float* L, const float* R, size_t n,
                  float* reL, float* imL, float* reR, float* imR);
int srt_istft_host(srt_ctx* ctx, const float* reL, const float* imL, const float* reR, const float* imR,
                   size_t frames, float* outL, float* outR);

/* ---- real-time streaming flavour (VST/Source/Spleeter4Stems.c) ------------------------------
 * One stereo stream, n_stems nets (the VST uses 4, all ELU).  Every 1024 new samples ("hop") the
 * newest 4096-sample frame is analysed (asymmetric window, Spleeter4Stems.c:383-416) and the frame
 * recorded two T-frame tiles earlier is synthesised with that tile's masks (:257-320); every T hops
 * the nets run on the finished tile in the background (:351-371).  Latency 2*T*1024 + 1024 samples.
 * srt_stream_process mirrors Spleeter4StemsProcessSamples (:512-582): `components` = 2*n_stems
 * planar outputs (stem-major, L then R), written only when output is available.
 * cfg: n_stems, time_step, bin_limit, device are used (flavour is forced to 1 = VST: exact sigmoid,
 * unclamped ELU); unaffected[s] = weight of the bins >= F (NULL = 0.25, 0.0, 0.25, 0.25 as
 * Spleeter4Stems.c:73,281). */
typedef struct srt_stream srt_stream;
int srt_stream_create(const srt_config* cfg, const float* const* coeffs, const float* unaffected, srt_stream** out);
int srt_stream_process(srt_stream* st, const float* inL, const float* inR, int n, float* const* components);
void srt_stream_destroy(srt_stream* st);
long long srt_stream_launch_count(const srt_stream* st);

/* ---- introspection ------------------------------------------------------------------------ */
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long long srt_launch_count(const srt_ctx* ctx);
/* srt_set_timing(ctx, 1) makes every kernel launch of the following calls be bracketed by CUDA
 * events on the context's stream (spans accumulate over calls; srt_set_timing clears them).
 * srt_last_timing sums the spans of one category, in ms: 0..9 = tensor-core layers down2..down6,
 * up1..up5; 10 down1, 11 up6, 12 up7, 13 STFT, 14 mask+iSTFT, 15 OLA, 16 H2D, 17 D2H. */
int srt_last_timing(const srt_ctx* ctx, int which, float* ms_out);
int srt_set_timing(srt_ctx* ctx, int enable);
/* copy an internal activation tensor of the last U-Net pass to the host, converted to the
 * reference's planar [stem][img][C][H][W] order.  name: "skip1".."skip6", "up1".."up6".
 * Returns the number of floats written, or a negative status. */
long long srt_debug_tensor(srt_ctx* ctx, const char* name, float* dst, size_t max_floats);
/* ---- measured roofline denominators (no context needed) ------------------------------------------------------------
 * srt_probe_tensor_peak: every SM issues M = 128, N = 256 tcgen05 MMAs back to back from shared memory for `seconds` of device
 *   time; kind 0 = kind::tf32 (the main term of the tensor-core layers), 1 = kind::f16 on bf16 operands (the compensation term).
 *   The result is the MMA pipe's issue rate x the clock the board holds under that load: the tensor roofline bench.py reports
 *   against (a short probe gives the burst figure, a multi-second one the power-capped sustained figure).
 * srt_probe_copy_bandwidth: best of 5 device-to-device copies of `bytes` bytes (read + write bytes per second): the HBM roofline
 *   measured by this library's own kernel, beside MEASURED_PEAKS.json. */
int srt_probe_tensor_peak(int device, int kind, double seconds, double* tflops_out);
int srt_probe_copy_bandwidth(int device, size_t bytes, double* gbs_out);
/* the cudaStream_t the context enqueues on (its own, or srt_config.cuda_stream): for callers that order their own device work
 * or CUDA events against the library's kernels (bench.py, the NCCL dispatcher of srt_dispatch.h) */
void* srt_cuda_stream(srt_ctx* ctx);
/* pinned host memory helpers (so callers can hand page-locked buffers to *_batch) */
void* srt_host_alloc(size_t bytes);
void srt_host_free(void* p);
int srt_synchronize(srt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
