// Implementation of the fftw3.h shim: forward/backward unnormalised DFT of size n (power of two or
// not), evaluated in DOUBLE and rounded once to float -- i.e. the correctly rounded result that
// any single-precision FFTW plan approximates to within its own rounding noise (FFTW_MEASURE makes
// the reference's fp32 rounding pattern machine dependent, SURVEY.md section 0 defect 6).
// Deliberately independent of oracle/scanner_oracle.cpp's Stockham transform: recursive
// radix-2 decimation in time, O(n^2) fallback for odd sizes.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>
#include "fftw3.h"

struct scn_shim_fftwf_plan_s {
  int n;
  int sign;
  fftwf_complex* in;
  fftwf_complex* out;
};

namespace {
typedef std::complex<double> cd;
const double kPi = 3.14159265358979323846264338327950288;

void dft_rec(const cd* x, cd* y, int n, int stride, int sign) {
  if (n == 1) { y[0] = x[0]; return; }
  if (n % 2) {
    for (int k = 0; k < n; k++) {
      cd acc(0, 0);
      for (int j = 0; j < n; j++) {
        const double a = sign * 2.0 * kPi * double((long long)j * k % n) / n;
        acc += x[j * stride] * cd(std::cos(a), std::sin(a));
      }
      y[k] = acc;
    }
    return;
  }
  const int h = n / 2;
  std::vector<cd> e(h), o(h);
  dft_rec(x, e.data(), h, stride * 2, sign);
  dft_rec(x + stride, o.data(), h, stride * 2, sign);
  for (int k = 0; k < h; k++) {
    const double a = sign * 2.0 * kPi * k / n;
    const cd w = cd(std::cos(a), std::sin(a)) * o[k];
    y[k] = e[k] + w;
    y[k + h] = e[k] - w;
  }
}
}  // namespace

extern "C" {
void* fftwf_malloc(size_t n) { void* p = nullptr; return posix_memalign(&p, 64, n ? n : 64) ? nullptr : p; }
void fftwf_free(void* p) { free(p); }
fftwf_complex* fftwf_alloc_complex(size_t n) { return static_cast<fftwf_complex*>(fftwf_malloc(sizeof(fftwf_complex) * n)); }
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned) {
  return new scn_shim_fftwf_plan_s{n, sign, in, out};
}
void fftwf_execute(const fftwf_plan p) {
  std::vector<cd> x(p->n), y(p->n);
  for (int i = 0; i < p->n; i++) x[i] = cd(p->in[i][0], p->in[i][1]);
  dft_rec(x.data(), y.data(), p->n, 1, p->sign);
  for (int i = 0; i < p->n; i++) { p->out[i][0] = float(y[i].real()); p->out[i][1] = float(y[i].imag()); }
}
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
}
