// CPU harness around csrc/nws_fft_mixed.cuh (the header the reverb's column kernels use) for
// tests/test_math_cpu.py: the same stage code, run serially (tid 0 of 1 thread).
#include "../../neural_waveshaping_synthesis_b200/csrc/nws_fft_mixed.cuh"
#include <stdlib.h>
#include <string.h>

extern "C" {

// W = 1 << log_w transforms of length n1 (125 or 250), interleaved layout buf[(e << log_w) + f], complex64.
// tw = exp(-2*pi*i*m/n1), m < n1.  Returns 0, or -1 for an unsupported length.
int h_fft_mixed(int n1, int inverse, int log_w, const float* in, const float* tw, float* out) {
  const size_t n = (size_t)n1 << log_w;
  float2* a = (float2*)malloc(n * sizeof(float2));
  float2* b = (float2*)malloc(n * sizeof(float2));
  memcpy(a, in, n * sizeof(float2));
  const float2* t = (const float2*)tw;
  float2* z = nullptr;
  if (n1 == 125) z = inverse ? nws_fft_mixed<125, true>(a, b, t, log_w, 0, 1) : nws_fft_mixed<125, false>(a, b, t, log_w, 0, 1);
  if (n1 == 250) z = inverse ? nws_fft_mixed<250, true>(a, b, t, log_w, 0, 1) : nws_fft_mixed<250, false>(a, b, t, log_w, 0, 1);
  if (z) memcpy(out, z, n * sizeof(float2));
  free(a);
  free(b);
  return z ? 0 : -1;
}
}

#include "../../neural_waveshaping_synthesis_b200/csrc/nws_fft_f64.cuh"

extern "C" {
// One forward transform of length 1 << log_n, complex128, natural order.  tw = exp(-2*pi*i*m/N), m < N/2.
void h_fft_f64(int log_n, const double* in, const double* tw, double* out) {
  const size_t n = (size_t)1 << log_n;
  double2* a = (double2*)malloc(n * sizeof(double2));
  double2* b = (double2*)malloc(n * sizeof(double2));
  memcpy(a, in, n * sizeof(double2));
  double2* z = nws_fft_f64(a, b, (const double2*)tw, log_n, 0, 1);
  memcpy(out, z, n * sizeof(double2));
  free(a);
  free(b);
}
}
