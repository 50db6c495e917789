// srt_internal.h — the few context internals the streaming flavour (srt_stream.cu) needs.
#pragma once
#include <cuda_runtime.h>

#include "../../include/srt_b200.h"

namespace srt {
namespace internal {
// Every entry point works on its context's device and leaves the caller's current device as it found it (a host process
// may drive torch, NCCL or another context on a different GPU from the same thread).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;          // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// U-Net on Bv images starting at image mag_img0 of the context's magnitude buffer, masks to mask_base[s][mask_img0 + b]; enqueued on ctx's stream
int ctx_run_unet(srt_ctx* ctx, int mag_img0, int Bv, float* mask_base, int mask_stride, int mask_img0);
float* ctx_mag(srt_ctx* ctx);   // the context's space-to-depth magnitude buffer: [max_batch_images] hi images, then as many lo images
cudaStream_t ctx_stream(srt_ctx* ctx);
int ctx_device(srt_ctx* ctx);
const float2* ctx_twiddle(srt_ctx* ctx);
void ctx_count_launch(srt_ctx* ctx, int n);
int set_error(int code, const char* msg);
}  // namespace internal
}  // namespace srt
