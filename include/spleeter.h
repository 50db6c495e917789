/*
 * spleeter.h — tier-A drop-in for SpleeterRT's Executable/spleeter.h (james34602/SpleeterRT).
 *
 * Same symbols, argument meaning and ownership rules as the reference so that
 * Executable/main.c compiles and links against libspleeterrt_b200.so unchanged:
 *   getCoeffSize / allocateSpleeterStr / initSpleeter / getMaskPtr / processSpleeter /
 *   freeSpleeter                                       (Executable/spleeter.h:64-69)
 * The U-Net itself runs on the B200 (tcgen05 implicit-GEMM kernels); these functions have no
 * error channel in the reference, so on any CUDA failure they print to stderr and abort().
 *
 * The weight containers keep the reference's memory layout (it is the file / blob format):
 * conv weights [O][I][5][5], transposed-conv weights [I][O][5][5], batchNorm = C offsets
 * followed by C scales (Executable/spleeter.h:5-31, spleeter.c:188).
 */
#ifndef SRT_TIERA_SPLEETER_H
#define SRT_TIERA_SPLEETER_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TBL_SIZE (1025)
#define TBL_SIZE_MINUS1 (TBL_SIZE - 1)

/* One net = 13 layers; the member list is shared by the fp32 and the IEEE-half containers. */
#define SRT_NET_MEMBERS(T)                                                                           \
    T down1_convWeight[5 * 5 * 2 * 16], down1_convBias[16], down1_batchNorm[16 * 2];                 \
    T down2_convWeight[5 * 5 * 16 * 32], down2_convBias[32], down2_batchNorm[32 * 2];                \
    T down3_convWeight[5 * 5 * 32 * 64], down3_convBias[64], down3_batchNorm[64 * 2];                \
    T down4_convWeight[5 * 5 * 64 * 128], down4_convBias[128], down4_batchNorm[128 * 2];             \
    T down5_convWeight[5 * 5 * 128 * 256], down5_convBias[256], down5_batchNorm[256 * 2];            \
    T down6_convWeight[5 * 5 * 256 * 512], down6_convBias[512];                                      \
    T up1_transp_convWeight[5 * 5 * 256 * 512], up1_transp_convBias[256], up1_batchNorm[256 * 2];    \
    T up2_transp_convWeight[5 * 5 * 128 * 512], up2_transp_convBias[128], up2_batchNorm[128 * 2];    \
    T up3_transp_convWeight[5 * 5 * 64 * 256], up3_transp_convBias[64], up3_batchNorm[64 * 2];       \
    T up4_transp_convWeight[5 * 5 * 32 * 128], up4_transp_convBias[32], up4_batchNorm[32 * 2];       \
    T up5_transp_convWeight[5 * 5 * 16 * 64], up5_transp_convBias[16], up5_batchNorm[16 * 2];        \
    T up6_transp_convWeight[5 * 5 * 1 * 32], up6_transp_convBias[1], up6_batchNorm[1 * 2];           \
    T up7_convWeight[4 * 4 * 1 * 2], up7_convBias[2];

typedef struct { SRT_NET_MEMBERS(float) } spleeterCoeff;                 /* 9 822 725 floats */
typedef struct { SRT_NET_MEMBERS(uint16_t) } spleeterQuantizedSubNet;    /* IEEE half, Executable/spleeter.h:32-58 */
typedef struct {
    spleeterQuantizedSubNet model1; /* ELU "drum" net   (main.c:759: coeffProvPtr2) */
    spleeterQuantizedSubNet model2; /* LeakyReLU/ReLU "vocal" net (main.c:760: coeffProvPtr1) */
} spleeterQuantized;

typedef struct _spleeter* spleeter;

size_t getCoeffSize(void);
void* allocateSpleeterStr(void);
/* width = analyseBinLimit (F), height = timeStep (T); stemMode 0: LeakyReLU/ReLU, else ELU.
 * Callers compiled against the VST flavour's prototype (int width, int height: VST/Source/spleeter.h:4) bind to the same symbol:
 * only the low 32 bits of the two sizes are used.
 * Both must be multiples of 64 (six halvings), 64 <= width <= 2048: other sizes - which the reference CLI merely warns
 * about, main.c:739-742 - end the process with exit status 2 and a message naming the restriction.  CUDA failures
 * (there is no error channel in this API) print the reason and abort().
 * `coeff` (one spleeterCoeff, host memory) must stay valid until freeSpleeter, as in the reference. */
void initSpleeter(struct _spleeter* nn, size_t width, size_t height, int stemMode, void* coeff);
/* hands out an internal HOST buffer of 2*T*F floats the caller may pass back as `y` */
void getMaskPtr(struct _spleeter* nn, float** mask);
/* x: magnitude [2][T][F] (host); y: soft mask [2][T][F] (host) */
void processSpleeter(struct _spleeter* nn, float* x, float* y);
void freeSpleeter(struct _spleeter* nn);

#ifdef __cplusplus
}
#endif
#endif
