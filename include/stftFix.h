/*
 * stftFix.h — tier-A drop-in for SpleeterRT's Executable/stftFix.h.
 *
 * InitSTFT / FreeSTFT / stft / istft keep the reference's contract (stftFix.h:32-35,
 * stftFix.c:302-579): 4096-point frames, hop 1024, symmetric Hann analysis window scaled by
 * 1/4096, spectra returned as four calloc'd planes of [frames][4096] floats (bins 0..2048
 * used, imaginary parts conjugated) that the CALLER frees with free(); istft returns two
 * calloc'd planes of frames*1024+3072 samples.  The transforms run on the B200.
 * OfflineSTFT stays a complete type because hosts malloc(sizeof(OfflineSTFT)) (main.c:775);
 * it is no larger than the reference's, so objects compiled against either header work.
 */
#ifndef SRT_TIERA_STFTFIX_H
#define SRT_TIERA_STFTFIX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum pt_state { SETUP, IDLE, WORKING, GET_OFF_FROM_WORK };

#define FFTSIZE 4096
#define LAP 4
#define HOPSIZE (FFTSIZE / LAP)
#define HALFWNDLEN ((FFTSIZE >> 1) + 1)
#define NPTDIV2 (FFTSIZE >> 2)

typedef struct {
    void* impl;          /* device context, owned by InitSTFT/FreeSTFT */
    size_t targetCore;   /* kept for source compatibility; the GPU path ignores it */
    unsigned char reserved[48];
} OfflineSTFT;

void InitSTFT(OfflineSTFT* st, size_t targetCore);
void FreeSTFT(OfflineSTFT* st);
size_t stft(OfflineSTFT* st, const float* dataL, const float* dataR, size_t data_size,
            float** resultLRe, float** resultLIm, float** resultRRe, float** resultRIm);
size_t istft(OfflineSTFT* st, float* dataLRe, float* dataLIm, float* dataRRe, float* dataRIm,
             size_t data_size, float** resultL, float** resultR);

#ifdef __cplusplus
}
#endif
#endif
