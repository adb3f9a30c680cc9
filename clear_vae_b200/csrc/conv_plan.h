// Host-side geometry of the convolution GEMMs ("gather plans"), shared by the
// tensor-core (conv_tc.cu) and exact-fp32 (conv_ffma.cu) kernels.
//
// Every conv-shaped op on the path is one gathered GEMM
//     dst[m, n] = sum_{t in taps} sum_{c < Cs}  pre(src[pixel(m) + tap t, c]) * W[n, c, t]
// in one of two gather modes:
//   * strided gather  (Conv2d forward, ConvTranspose2d data-gradient):
//       source row  hs = hd * stride - pad + kh,  all k*k taps, one class;
//   * class gather    (ConvTranspose2d forward, Conv2d data-gradient):
//       the big grid (stride x larger) is split by output parity (a, b); class (a, b)
//       only touches the taps with (a + pad - kh) % stride == 0, so no MAC is spent on
//       the zeros a zero-insertion formulation would multiply:
//       hb = hd * stride + a,  hs = hd + (a + pad - kh) / stride.
// Reference semantics: nn.Conv2d / nn.ConvTranspose2d as used at vae.py:15-46,113-156
// (SURVEY.md §8a' "Layer semantics").
#pragma once
#include <stdint.h>

#include "clearvae_b200.h"

namespace cvplan {

constexpr int kMaxTaps = 16;
constexpr int kMaxClasses = 4;

struct Cls {
  int Hd, Wd;      // class-local dst grid
  int oa, ob;      // dst offset inside the big grid
  int ntaps;
  int w_off;       // element offset of this class inside the packed weight buffer
  int Kp;          // padded K of this class in the packed buffer
  int8_t dh[kMaxTaps], dw[kMaxTaps];
  int16_t wtap[kMaxTaps];  // kh * k + kw
};

struct Plan {
  int n_classes;
  int sh;        // source step per dst step
  int os;        // dst step inside the big grid
  int Hs, Ws, Cs;  // source dims / GEMM-K channels
  int Hb, Wb, Nn;  // dst big-grid dims / GEMM-N channels
  long long ws_n, ws_c;  // strides of (n, c) in the reference weight layout
  int kk;        // k * k
  Cls cls[kMaxClasses];
};

inline int out_size(const clearvae_conv_geom& g, int in) {
  return g.transposed ? (in - 1) * g.stride - 2 * g.pad + g.k + g.out_pad : (in + 2 * g.pad - g.k) / g.stride + 1;
}

enum Role { kFprop = 0, kDgrad = 1 };

// returns false when the geometry is not supported
inline bool make_plan(const clearvae_conv_geom& g, int role, int k_align, Plan* p) {
  if (g.k < 1 || g.k > 4 || g.stride < 1 || g.stride > 2 || g.k * g.k > kMaxTaps) return false;
  const int Hout = out_size(g, g.Hin), Wout = out_size(g, g.Win);
  if (Hout <= 0 || Wout <= 0) return false;
  const bool strided = (!g.transposed && role == kFprop) || (g.transposed && role == kDgrad);
  p->kk = g.k * g.k;
  const long long kk = p->kk;
  if (!g.transposed) {
    // weight [Cout, Cin, k, k]
    if (role == kFprop) { p->Cs = g.Cin; p->Nn = g.Cout; p->ws_n = g.Cin * kk; p->ws_c = kk; }
    else                { p->Cs = g.Cout; p->Nn = g.Cin; p->ws_n = kk; p->ws_c = g.Cin * kk; }
  } else {
    // weight [Cin, Cout, k, k]
    if (role == kFprop) { p->Cs = g.Cin; p->Nn = g.Cout; p->ws_n = kk; p->ws_c = g.Cout * kk; }
    else                { p->Cs = g.Cout; p->Nn = g.Cin; p->ws_n = g.Cout * kk; p->ws_c = kk; }
  }
  auto pad_k = [&](int k) { return (k + k_align - 1) / k_align * k_align; };
  int w_off = 0;
  const int n_pad = (p->Nn + 15) / 16 * 16;
  if (strided) {
    // big side is the source
    const int Hbig = g.transposed ? Hout : g.Hin, Wbig = g.transposed ? Wout : g.Win;
    const int Hsm = g.transposed ? g.Hin : Hout, Wsm = g.transposed ? g.Win : Wout;
    p->n_classes = 1; p->sh = g.stride; p->os = 1;
    p->Hs = Hbig; p->Ws = Wbig; p->Hb = Hsm; p->Wb = Wsm;
    Cls& c = p->cls[0];
    c.Hd = Hsm; c.Wd = Wsm; c.oa = c.ob = 0; c.ntaps = 0;
    for (int kh = 0; kh < g.k; ++kh)
      for (int kw = 0; kw < g.k; ++kw) {
        c.dh[c.ntaps] = (int8_t)(kh - g.pad); c.dw[c.ntaps] = (int8_t)(kw - g.pad);
        c.wtap[c.ntaps] = (int16_t)(kh * g.k + kw); ++c.ntaps;
      }
    c.w_off = 0; c.Kp = pad_k(c.ntaps * p->Cs);
    return true;
  }
  const int Hbig = g.transposed ? Hout : g.Hin, Wbig = g.transposed ? Wout : g.Win;
  const int Hsm = g.transposed ? g.Hin : Hout, Wsm = g.transposed ? g.Win : Wout;
  p->sh = 1; p->os = g.stride; p->Hs = Hsm; p->Ws = Wsm; p->Hb = Hbig; p->Wb = Wbig;
  p->n_classes = 0;
  for (int a = 0; a < g.stride; ++a)
    for (int b = 0; b < g.stride; ++b) {
      Cls& c = p->cls[p->n_classes];
      c.oa = a; c.ob = b;
      c.Hd = (Hbig - a + g.stride - 1) / g.stride;
      c.Wd = (Wbig - b + g.stride - 1) / g.stride;
      c.ntaps = 0;
      for (int kh = 0; kh < g.k; ++kh) {
        if ((a + g.pad - kh) % g.stride != 0) continue;
        for (int kw = 0; kw < g.k; ++kw) {
          if ((b + g.pad - kw) % g.stride != 0) continue;
          c.dh[c.ntaps] = (int8_t)((a + g.pad - kh) / g.stride);
          c.dw[c.ntaps] = (int8_t)((b + g.pad - kw) / g.stride);
          c.wtap[c.ntaps] = (int16_t)(kh * g.k + kw);
          ++c.ntaps;
        }
      }
      c.w_off = w_off; c.Kp = pad_k(c.ntaps * p->Cs);
      w_off += n_pad * c.Kp;
      if (c.Hd > 0 && c.Wd > 0) ++p->n_classes;  // empty classes cannot occur for stride <= 2, Hbig >= 2
    }
  return p->n_classes > 0;
}

inline long long packed_weight_elems(const Plan& p) {
  const int n_pad = (p.Nn + 15) / 16 * 16;
  long long tot = 0;
  for (int i = 0; i < p.n_classes; ++i) tot += (long long)n_pad * p.cls[i].Kp;
  return tot;
}

}  // namespace cvplan
