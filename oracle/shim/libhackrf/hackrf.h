// Shim for <libhackrf/hackrf.h> (libhackrf is not installed here).  TEST INFRASTRUCTURE: lets the
// reference's hackRFSource.cpp compile unmodified so ref_tool can drive its rx callback
// (hackRFSource.cpp:180-264) with captured / synthetic sweep transfers.  Declares exactly the API
// that file uses; the stubs in shim_hackrf.cpp succeed without hardware and remember the callback.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum hackrf_error {
  HACKRF_SUCCESS = 0,
  HACKRF_TRUE = 1,
  HACKRF_ERROR_INVALID_PARAM = -2,
  HACKRF_ERROR_NOT_FOUND = -5,
  HACKRF_ERROR_OTHER = -9999
};

enum sweep_style { LINEAR = 0, INTERLEAVED = 1 };

typedef struct hackrf_device hackrf_device;

typedef struct {
  hackrf_device* device;
  uint8_t* buffer;
  int buffer_length;
  int valid_length;
  void* rx_ctx;
  void* tx_ctx;
} hackrf_transfer;

typedef int (*hackrf_sample_block_cb_fn)(hackrf_transfer* transfer);

int hackrf_init(void);
int hackrf_open(hackrf_device** device);
int hackrf_close(hackrf_device* device);
int hackrf_board_id_read(hackrf_device* device, uint8_t* value);
int hackrf_version_string_read(hackrf_device* device, char* version, uint8_t length);
int hackrf_set_sample_rate(hackrf_device* device, const double freq_hz);
uint32_t hackrf_compute_baseband_filter_bw(const uint32_t bandwidth_hz);
int hackrf_set_baseband_filter_bandwidth(hackrf_device* device, const uint32_t bandwidth_hz);
int hackrf_set_lna_gain(hackrf_device* device, uint32_t value);
int hackrf_set_vga_gain(hackrf_device* device, uint32_t value);
int hackrf_set_amp_enable(hackrf_device* device, const uint8_t value);
int hackrf_set_antenna_enable(hackrf_device* device, const uint8_t value);
int hackrf_set_freq(hackrf_device* device, const uint64_t freq_hz);
int hackrf_set_scan_parameters(hackrf_device* device, uint64_t start, uint64_t stop, uint32_t step);
int hackrf_start_rx(hackrf_device* device, hackrf_sample_block_cb_fn callback, void* rx_ctx);
int hackrf_stop_rx(hackrf_device* device);
int hackrf_init_sweep(hackrf_device* device, const uint16_t* frequency_list, const int num_ranges,
                      const uint32_t num_bytes, const uint32_t step_width, const uint32_t offset,
                      const enum sweep_style style);
const char* hackrf_error_name(enum hackrf_error errcode);

// ---- shim-only controls (not part of libhackrf)
// Delivers one transfer to the callback registered by hackrf_start_rx; returns -1000 if none.
int shim_hackrf_deliver(uint8_t* buffer, int valid_length);
// 1 once hackrf_stop_rx has been called (the reference's ThreadWorker calls it when it leaves its loop).
int shim_hackrf_rx_stopped(void);

#ifdef __cplusplus
}
#endif
