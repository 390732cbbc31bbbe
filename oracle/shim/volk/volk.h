/* Shim for <volk/volk.h> (VOLK is not installed / not part of the reference tree).
 * volk_32fc_32f_multiply_32fc_a (process.cpp:30-33): c[i] = a[i] * b[i], complex times real,
 * i.e. one IEEE fp32 multiply per component -- what every VOLK protokernel computes. */
#ifndef SCN_SHIM_VOLK_H_
#define SCN_SHIM_VOLK_H_
#include <complex>
typedef std::complex<float> lv_32fc_t;
static inline void volk_32fc_32f_multiply_32fc_a(lv_32fc_t* c, const lv_32fc_t* a, const float* b,
                                                 unsigned int num_points) {
  const float* af = reinterpret_cast<const float*>(a);
  float* cf = reinterpret_cast<float*>(c);
  for (unsigned int i = 0; i < num_points; i++) {
    const float w = b[i];
    const float re = af[2 * i] * w, im = af[2 * i + 1] * w;
    cf[2 * i] = re;
    cf[2 * i + 1] = im;
  }
}
#endif
