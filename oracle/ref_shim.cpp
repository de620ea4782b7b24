// C-ABI shim over the UNMODIFIED reference C++ sources, compiled from where
// they lie under /root/reference (never copied into this repo).
// TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_vs_reference.py to pin
// the float oracle's accumulation order and the I/Q semantics.
//
//   Fir1D::fir_1d_idel   fir_1d/model/cpp/fir_1d_ref.cpp:37-63  ("full"
//       convolution, length N+L-1, y[n] = sum_k h[k]*x[n-k], k ascending) —
//       the Python model's same-mode output is its slice [L/2, L/2+N).
//   Fir1D::process_sample fir_1d_ref.cpp:16-33 (causal streaming form).
//   MyComplex operator*(double), operator+  lib/mycomplex.cpp:12-22 — a real
//       tap times a complex sample scales Re and Im independently.
#include <cstddef>
#include <vector>

#include "fir_1d_ref.h"
#include "mycomplex.h"

extern "C" {

// y must hold n + l - 1 doubles.  Returns the number written.
long ref_fir1d_idel_full(const double* x, long n, const double* h, long l, double* y) {
  std::vector<double> xv(x, x + n), hv(h, h + l);
  std::vector<double> out = Fir1D::fir_1d_idel(xv, hv);
  for (std::size_t i = 0; i < out.size(); ++i) y[i] = out[i];
  return static_cast<long>(out.size());
}

// Streaming form: y[i] = process_sample(x[i]) for i < n, state starts at 0.
void ref_fir1d_stream(const double* x, long n, const double* h, long l, double* y) {
  Fir1D f(std::vector<double>(h, h + l));
  for (long i = 0; i < n; ++i) y[i] = f.process_sample(x[i]);
}

// acc = sum_k (z[k] * h[k]) using only the operators MyComplex defines.
// z is interleaved (re, im); the magnitude of the result goes to out_mag and
// the components are recovered through operator- against pure re / im parts.
void ref_complex_real_tap_mac(const double* z_interleaved, const double* h, long l,
                              double* out_re, double* out_im) {
  MyComplex acc(0.0, 0.0);
  MyComplex acc_re_only(0.0, 0.0), acc_im_only(0.0, 0.0);
  for (long k = 0; k < l; ++k) {
    MyComplex zk(z_interleaved[2 * k], z_interleaved[2 * k + 1]);
    acc = acc + zk * h[k];
    acc_re_only = acc_re_only + MyComplex(z_interleaved[2 * k], 0.0) * h[k];
    acc_im_only = acc_im_only + MyComplex(0.0, z_interleaved[2 * k + 1]) * h[k];
  }
  // MyComplex keeps x_/y_ private; |acc - acc_im_only| == |Re|, etc.
  double re_mag = (acc - acc_im_only).get_magnitude();
  double im_mag = (acc - acc_re_only).get_magnitude();
  // Sign: compare against the component-only accumulators' magnitudes after
  // adding a large positive offset (operator+(double, complex) shifts Re).
  double re_shift = (1.0e6 + (acc - acc_im_only)).get_magnitude();
  *out_re = re_shift >= 1.0e6 ? re_mag : -re_mag;
  MyComplex im_as_re = MyComplex(0.0, 0.0) + (acc - acc_re_only);
  (void)im_as_re;
  // Im sign via division by i: (0 + i*b) / (0 + i*1) = b + i*0.
  MyComplex q = (acc - acc_re_only) / MyComplex(0.0, 1.0);
  double im_shift = (1.0e6 + q).get_magnitude();
  *out_im = im_shift >= 1.0e6 ? im_mag : -im_mag;
}

}  // extern "C"
