// srt_weights.cpp — the weight file formats either side of srt_create.  Host only.
//
//   * fp32 `.dat` dumps: one spleeterCoeff per file, 39 290 900 bytes, read with a single fread by the VST
//     (VST/Source/PluginProcessor.cpp:48-80: drum4stems.dat, bass4stems.dat, accompaniment4stems.dat, vocal4stems.dat);
//   * the fp16 model blob: spleeterQuantized = consecutive nets of 9 822 725 halves in spleeterCoeff member order
//     (Executable/spleeter.h:32-62), expanded by f32Decompress with denormals flushed (main.c:423-443);
//   * the packer that turns one spleeterCoeff into the k-block-major, 128B-swizzled B-operand blobs the tcgen05
//     kernels stream with bulk copies (srt_plan.cpp pack_layer / pack_row_layer), exposed for offline tooling.
// No CUDA here; failures are error codes with srt_last_error() text, never a silent partial read (the reference
// ignores fread's result, PluginProcessor.cpp:60).
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "srt_internal.h"
#include "srt_plan.h"

using namespace srt;

namespace {
constexpr size_t kDatBytes = (size_t)kCoeffFloats * 4;   // 39 290 900

int failf(int code, const char* fmt, const char* path, const char* detail)
{
    char buf[768];
    snprintf(buf, sizeof buf, fmt, path ? path : "(null)", detail ? detail : "");
    return internal::set_error(code, buf);
}

int file_size(FILE* f, size_t* out)
{
    if (fseek(f, 0, SEEK_END) != 0) return -1;
    const long n = ftell(f);
    if (n < 0 || fseek(f, 0, SEEK_SET) != 0) return -1;
    *out = (size_t)n;
    return 0;
}
}  // namespace

extern "C" int srt_load_coeff_dat(const char* path, float* coeff_out)
{
    if (!path || !coeff_out) return internal::set_error(SRT_ERR_ARG, "srt_load_coeff_dat: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return failf(SRT_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    size_t n = 0;
    if (file_size(f, &n) != 0 || n != kDatBytes) {
        fclose(f);
        return failf(SRT_ERR_ARG, "%s is not a spleeterCoeff dump (need exactly 39290900 bytes)%s", path, "");
    }
    const size_t got = fread(coeff_out, 1, kDatBytes, f);
    fclose(f);
    if (got != kDatBytes) return failf(SRT_ERR_ARG, "short read from %s%s", path, "");
    return 0;
}

extern "C" int srt_save_coeff_dat(const char* path, const float* coeff)
{
    if (!path || !coeff) return internal::set_error(SRT_ERR_ARG, "srt_save_coeff_dat: null argument");
    FILE* f = fopen(path, "wb");
    if (!f) return failf(SRT_ERR_ARG, "cannot create %s: %s", path, strerror(errno));
    const size_t put = fwrite(coeff, 1, kDatBytes, f);
    const int rc = fclose(f);
    if (put != kDatBytes || rc != 0) return failf(SRT_ERR_ARG, "short write to %s%s", path, "");
    return 0;
}

extern "C" int srt_model_fp16_nets(const char* path)
{
    if (!path) return internal::set_error(SRT_ERR_ARG, "srt_model_fp16_nets: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return failf(SRT_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    size_t n = 0;
    const int rc = file_size(f, &n);
    fclose(f);
    if (rc != 0 || n == 0 || n % ((size_t)kCoeffFloats * 2) != 0)
        return failf(SRT_ERR_ARG, "%s is not a whole number of fp16 nets (9822725 halves each)%s", path, "");
    return (int)(n / ((size_t)kCoeffFloats * 2));
}

extern "C" int srt_load_model_fp16(const char* path, int net, float* coeff_out)
{
    if (!coeff_out) return internal::set_error(SRT_ERR_ARG, "srt_load_model_fp16: null argument");
    const int nets = srt_model_fp16_nets(path);
    if (nets < 0) return nets;
    if (net < 0 || net >= nets) return failf(SRT_ERR_ARG, "%s: net index out of range%s", path, "");
    FILE* f = fopen(path, "rb");
    if (!f) return failf(SRT_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    std::vector<uint16_t> h(kCoeffFloats);
    const bool ok = fseek(f, (long)((size_t)net * kCoeffFloats * 2), SEEK_SET) == 0 && fread(h.data(), 2, kCoeffFloats, f) == (size_t)kCoeffFloats;
    fclose(f);
    if (!ok) return failf(SRT_ERR_ARG, "short read from %s%s", path, "");
    srt_half_to_float(h.data(), coeff_out, kCoeffFloats);
    return 0;
}

// layer: 0..4 = down2..down6, 5..9 = up1..up5 (the ten tensor-core layers).  form 0 = generic kernel blob
// [phase][n-tile][k-block][n_tile][32], form 1 = row-patch blob [k-block][N][32] (down2, down3, up4, up5 only).
// Returns the number of floats of the blob (written to `out` when it is not NULL and cap suffices), or a negative status.
extern "C" long long srt_pack_layer(int layer, int form, int time_step, int bin_limit, const float* coeff, float* out, size_t cap_floats)
{
    if (!coeff) return internal::set_error(SRT_ERR_ARG, "srt_pack_layer: null coefficients");
    if (layer < 0 || layer > 9 || (form != 0 && form != 1)) return internal::set_error(SRT_ERR_ARG, "srt_pack_layer: layer 0..9, form 0 or 1");
    if (time_step < 64 || time_step % 64 || bin_limit < 64 || bin_limit % 64 || bin_limit > 2048)
        return internal::set_error(SRT_ERR_ARG, "srt_pack_layer: time_step and bin_limit must be multiples of 64, bin_limit <= 2048");
    const bool split = !weights_tf32_exact(coeff);
    const NetGeom g{time_step, bin_limit};
    if (form == 1) {
        if (!row_plan_supported(layer)) return internal::set_error(SRT_ERR_ARG, "srt_pack_layer: no row-patch form for this layer");
        const RowPlan rp = build_row_plan(g, layer, split);
        if (out) {
            if (cap_floats < rp.w_floats_per_stem) return internal::set_error(SRT_ERR_CAPACITY, "srt_pack_layer: output buffer too small");
            pack_row_layer(rp, coeff, out);
        }
        return (long long)rp.w_floats_per_stem;
    }
    const std::vector<LayerPlan> plans = build_plans(g, 1, split);
    const LayerPlan& L = plans[layer];
    if (out) {
        if (cap_floats < L.w_floats_per_stem) return internal::set_error(SRT_ERR_CAPACITY, "srt_pack_layer: output buffer too small");
        pack_layer(L, coeff, out);
    }
    return (long long)L.w_floats_per_stem;
}
