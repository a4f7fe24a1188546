/* diffspectra_b200 — C-ABI of the B200-native DiffSpectra sampling hot path.  (work in progress header;
 * the full list of entry points and the reference interface each replaces is below) */
#ifndef DIFFSPECTRA_B200_H
#define DIFFSPECTRA_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ds_ctx ds_ctx;
const char* ds_last_error(void);
int ds_version(void);
int ds_create(ds_ctx** out, int device, int mode, int spectra_version);
int ds_destroy(ds_ctx* ctx);
long long ds_launch_count(ds_ctx* ctx);
int ds_gemm(ds_ctx* ctx, int use_tensor_cores, const void* A, int lda, const void* W, int ldw, const float* bias,
            const float* addmat, int ldadd, void* out, int ldo, int M, int N, int K, int in_dtype, int out_dtype,
            int act, void* stream);
#ifdef __cplusplus
}
#endif
#endif
