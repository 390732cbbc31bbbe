// Stubs behind the libhackrf shim header + a deterministic clock.  TEST INFRASTRUCTURE.
#include <atomic>
#include <cstring>
#include <ctime>
#include "libhackrf/hackrf.h"

struct hackrf_device { int unused; };

namespace {
hackrf_device g_device;
hackrf_sample_block_cb_fn g_callback = nullptr;
void* g_rx_ctx = nullptr;
std::atomic<int> g_stopped(0);
std::atomic<long> g_clock_calls(0);
}

extern "C" {
int hackrf_init(void) { return HACKRF_SUCCESS; }
int hackrf_open(hackrf_device** device) { *device = &g_device; return HACKRF_SUCCESS; }
int hackrf_close(hackrf_device*) { return HACKRF_SUCCESS; }
int hackrf_board_id_read(hackrf_device*, uint8_t* value) { *value = 2; return HACKRF_SUCCESS; }
int hackrf_version_string_read(hackrf_device*, char* version, uint8_t length) {
  strncpy(version, "shim", length);
  return HACKRF_SUCCESS;
}
int hackrf_set_sample_rate(hackrf_device*, const double) { return HACKRF_SUCCESS; }
uint32_t hackrf_compute_baseband_filter_bw(const uint32_t bandwidth_hz) { return bandwidth_hz; }
int hackrf_set_baseband_filter_bandwidth(hackrf_device*, const uint32_t) { return HACKRF_SUCCESS; }
int hackrf_set_lna_gain(hackrf_device*, uint32_t) { return HACKRF_SUCCESS; }
int hackrf_set_vga_gain(hackrf_device*, uint32_t) { return HACKRF_SUCCESS; }
int hackrf_set_amp_enable(hackrf_device*, const uint8_t) { return HACKRF_SUCCESS; }
int hackrf_set_antenna_enable(hackrf_device*, const uint8_t) { return HACKRF_SUCCESS; }
int hackrf_set_freq(hackrf_device*, const uint64_t) { return HACKRF_SUCCESS; }
int hackrf_set_scan_parameters(hackrf_device*, uint64_t, uint64_t, uint32_t) { return HACKRF_SUCCESS; }
int hackrf_start_rx(hackrf_device*, hackrf_sample_block_cb_fn callback, void* rx_ctx) {
  g_callback = callback;
  g_rx_ctx = rx_ctx;
  return HACKRF_SUCCESS;
}
int hackrf_stop_rx(hackrf_device*) { g_stopped = 1; return HACKRF_SUCCESS; }
int hackrf_init_sweep(hackrf_device*, const uint16_t*, const int, const uint32_t, const uint32_t, const uint32_t,
                      const enum sweep_style) { return HACKRF_SUCCESS; }
const char* hackrf_error_name(enum hackrf_error) { return "shim"; }

int shim_hackrf_deliver(uint8_t* buffer, int valid_length) {
  if (!g_callback) return -1000;
  hackrf_transfer t;
  t.device = &g_device;
  t.buffer = buffer;
  t.buffer_length = valid_length;
  t.valid_length = valid_length;
  t.rx_ctx = g_rx_ctx;
  t.tx_ctx = nullptr;
  return g_callback(&t);
}
int shim_hackrf_rx_stopped(void) { return g_stopped.load(); }

// The reference stamps scan starts with time(NULL) (hackRFSource.cpp:247-249) and prints them
// (process.cpp:280-287).  Interposed here so the golden stdout is reproducible: the k-th call
// returns 1500000000 + 1000 k.
time_t time(time_t* out) {
  const time_t t = time_t(1500000000L + 1000L * g_clock_calls.fetch_add(1));
  if (out) *out = t;
  return t;
}
}
