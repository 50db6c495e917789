/* Test-infrastructure shim (oracle/ only): gives ctypes access to the reference's resampling front end, compiled from
 * the reference's own sources where they lie (oracle/build_ref.py adds -I /root/reference/Executable):
 *   - decompressResamplerMQ + its static knot table (main.c:132-208) -> the 22438-entry sinc coefficient table
 *   - JamesDSPOfflineResampling (main.c:209-224) -> libsamplerate src_simple
 * main.c is included as a whole (its main() renamed); nothing of it is copied into this repository. */
#define main ref_cli_main
#include "main.c"
#undef main

void ref_resampler_table(float *out22438) { decompressResamplerMQ(compressedCoeffMQ, out22438); }

void ref_resample(const float *in, float *out, size_t len_in, size_t len_out, int channels, double ratio, float *table22438)
{
    decompressedCoefficients = table22438;
    JamesDSPOfflineResampling(in, out, len_in, len_out, channels, ratio);
}
