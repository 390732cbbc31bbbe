/* Shim for <gnuradio/fft/window.h> (GNU Radio is not installed / not part of the reference tree).
 * gr::fft::window::build(type, ntaps, beta) as called at process.cpp:18: the published gr-fft
 * definitions -- symmetric windows over M = ntaps - 1, evaluated in double, stored as float;
 * blackman_harris defaults to the 92 dB 4-term set.  (The window crosses the product's C ABI as a
 * table, so this shim only matters for running the reference's own ProcessSamples.) */
#ifndef SCN_SHIM_GR_WINDOW_H_
#define SCN_SHIM_GR_WINDOW_H_
#include <cmath>
#include <vector>
namespace gr { namespace fft {
class window {
 public:
  enum win_type {
    WIN_NONE = -1, WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2, WIN_RECTANGULAR = 3,
    WIN_KAISER = 4, WIN_BLACKMAN_hARRIS = 5, WIN_BLACKMAN_HARRIS = 5, WIN_BARTLETT = 6, WIN_FLATTOP = 7
  };
  static std::vector<float> build(win_type type, int ntaps, double /*beta*/) {
    const double pi = 3.14159265358979323846264338327950288;
    std::vector<float> taps(ntaps);
    const double M = double(ntaps - 1);
    for (int n = 0; n < ntaps; n++) {
      const double x = ntaps > 1 ? double(n) / M : 0.0;
      double w = 1.0;
      switch (type) {
        case WIN_HAMMING: w = 0.54 - 0.46 * std::cos(2 * pi * x); break;
        case WIN_HANN: w = 0.5 - 0.5 * std::cos(2 * pi * x); break;
        case WIN_BLACKMAN: w = 0.42 - 0.5 * std::cos(2 * pi * x) + 0.08 * std::cos(4 * pi * x); break;
        case WIN_BLACKMAN_hARRIS:
          w = 0.35875 - 0.48829 * std::cos(2 * pi * x) + 0.14128 * std::cos(4 * pi * x) -
              0.01168 * std::cos(6 * pi * x);
          break;
        default: w = 1.0; break;
      }
      taps[n] = float(w);
    }
    return taps;
  }
};
} }
#endif
