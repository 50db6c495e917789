/*
 * Spleeter4Stems.h — tier-A drop-in for SpleeterRT's VST/Source/Spleeter4Stems.h: the real-time
 * 4-stem streamer the JUCE plugin drives (PluginProcessor.cpp:123-181).
 *
 * Same three entry points and the same argument meaning (Spleeter4Stems.h:67-69):
 *   Spleeter4StemsInit(msr, spectralBinLimit, timeStep, coeffProvider[4])   4 x spleeterCoeff (fp32)
 *   Spleeter4StemsProcessSamples(msr, inL, inR, n, components[8])           planar stem outputs
 *   Spleeter4StemsFree(msr)
 * `Spleeter4Stems` stays a complete type because the host mallocs it (PluginProcessor.cpp:123);
 * it is far smaller than the reference's, so objects built against either header work.
 * All transforms and the four U-Nets run on the B200; there is no error channel, failures abort.
 */
#ifndef SRT_TIERA_SPLEETER4STEMS_H
#define SRT_TIERA_SPLEETER4STEMS_H
#ifdef __cplusplus
extern "C" {
#endif

#define FFTSIZE 4096
#define ANALYSIS_OVERLAP 4
#define OVPSIZE (FFTSIZE / ANALYSIS_OVERLAP)
#define OUTPUTSEG ((OVPSIZE >> 1) << 1)
#define SAMPLESHIFT (FFTSIZE - (OVPSIZE << 1))
#define HALFWNDLEN ((FFTSIZE >> 1) + 1)
#define LATENCY ((OVPSIZE << 1) - OUTPUTSEG)
#define COMPONENTS 8

typedef struct {
    void* impl;                  /* device-side streamer, owned by Init/Free */
    int analyseBinLimit, timeStep;
    unsigned char reserved[48];
} Spleeter4Stems;

void Spleeter4StemsInit(Spleeter4Stems* msr, int initSpectralBinLimit, int initTimeStep, void* coeffProvider[4]);
void Spleeter4StemsFree(Spleeter4Stems* msr);
void Spleeter4StemsProcessSamples(Spleeter4Stems* msr, const float* inLeft, const float* inRight, int inSampleCount, float** components);

#ifdef __cplusplus
}
#endif
#endif
