/* Shim for <fftw3.h> (FFTW 3 is not installed in this image and is not part of the reference
 * tree).  Only the names the reference's hot path uses (fft.h:3, fft.cpp:8-10,14-16,23;
 * process.cpp:19,25,103,105,113) are declared; the implementation is oracle/shim/shim_impl.cpp.
 * TEST INFRASTRUCTURE: lets the reference's own .cpp files be compiled, unmodified, where they
 * lie, into oracle/_ref/. */
#ifndef SCN_SHIM_FFTW3_H_
#define SCN_SHIM_FFTW3_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct scn_shim_fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void* fftwf_malloc(size_t n);
void fftwf_free(void* p);
fftwf_complex* fftwf_alloc_complex(size_t n);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan plan);
void fftwf_destroy_plan(fftwf_plan plan);
#ifdef __cplusplus
}
#endif
#endif
